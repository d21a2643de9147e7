#!/bin/bash
# compare builds in variants/ (plus "base") on the grid-wide random starts: INIT phase only
mkdir -p gpurun_out
for v in base "$@"; do
lib=$PWD/variants/lib_$v.so; [ $v = base ] && lib=$PWD/gibbssampling_b200/libgibbs_b200.so
echo "== $v"
GIBBS_B200_LIB=$lib GIBBS_B200_INIT_KERNEL=1 timeout 600 python - <<'PY'
from gibbssampling_b200.engine import GibbsEngine, make_params
from gibbssampling_b200.synthetic import planted_motif_set, background_of
from gibbssampling_b200 import _abi
for (n,L,k,chains,reps) in [(1000,500,12,1024,3),(10000,1000,16,64,2),(100000,200,20,8,2)]:
    ps = planted_motif_set(n, L, k); bg = background_of(ps.ascii, 1e-4, 5)
    eng = GibbsEngine(ps.sequences())
    pi = make_params(k, 1e-4, 5, bg, phase_mask=_abi.PHASE_INIT)
    for rep in range(reps):
        r = eng.run(pi, chains, seed=1+rep, want_sites=False, want_scores=False, want_counts=False); st=r.stats
        print(n,L,k,chains,"INIT only kernel_ms",round(st['kernel_ms'],3),"draws/s %.3e"%(st['site_updates']*(n-1)/(st['kernel_ms']*1e-3)),flush=True)
    eng.close()
PY
done
