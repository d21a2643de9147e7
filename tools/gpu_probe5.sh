#!/bin/bash
# A/B of the grid-wide random-start kernel (GIBBS_B200_INIT_KERNEL=0 keeps the random starts in the chain kernel)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
for mode in 1 0; do
GIBBS_B200_INIT_KERNEL=$mode timeout 900 python - > gpurun_out/probe_init$mode.log 2>&1 <<'PY'
from gibbssampling_b200.engine import GibbsEngine, make_params
from gibbssampling_b200.synthetic import planted_motif_set, background_of
from gibbssampling_b200 import _abi
for (n,L,k,chains,reps) in [(1000,500,12,1024,4),(1000,500,12,64,2),(10000,1000,16,64,1),(10000,1000,16,512,1),(100000,200,20,8,1),(100000,200,20,148,1)]:
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
done
