#!/bin/bash
# compare builds in variants/ on INIT-only and full runs
mkdir -p gpurun_out
for v in "$@"; do
for mode in 1 0; do
GIBBS_B200_LIB=$PWD/variants/lib_$v.so GIBBS_B200_INIT_KERNEL=$mode timeout 600 python - $mode > gpurun_out/var_${v}_init$mode.log 2>&1 <<'PY'
import sys
from gibbssampling_b200.engine import GibbsEngine, make_params
from gibbssampling_b200.synthetic import planted_motif_set, background_of
from gibbssampling_b200 import _abi
mode = int(sys.argv[1])
cases = [(1000,500,12,1024,3),(10000,1000,16,64,1),(10000,1000,16,512,1)]
if mode == 1: cases.append((100000,200,20,8,1))
for (n,L,k,chains,reps) in cases:
    ps = planted_motif_set(n, L, k); bg = background_of(ps.ascii, 1e-4, 5)
    eng = GibbsEngine(ps.sequences()); p = make_params(k, 1e-4, 5, bg)
    pi = make_params(k, 1e-4, 5, bg, phase_mask=_abi.PHASE_INIT)
    for rep in range(reps):
        r = eng.run(pi, chains, seed=1+rep, want_sites=False, want_scores=False, want_counts=False); st=r.stats
        print(n,L,k,chains,"INIT only kernel_ms",round(st['kernel_ms'],3),"draws/s %.3e"%(st['site_updates']*(n-1)/(st['kernel_ms']*1e-3)),flush=True)
        r = eng.run(p, chains, seed=1+rep, want_sites=False, want_scores=False, want_counts=False); st=r.stats
        print(n,L,k,chains,"full     kernel_ms",round(st['kernel_ms'],3),"win/s %.3e"%(st['window_scores']/(st['kernel_ms']*1e-3)),"launches",st['kernel_launches'],flush=True)
    eng.close()
PY
echo "== $v init_kernel=$mode"; cat gpurun_out/var_${v}_init$mode.log
done
done
