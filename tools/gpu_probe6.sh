#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 5 gpurun_out/pytest_gpu.log
timeout 900 python - > gpurun_out/probe6.log 2>&1 <<'PY'
from gibbssampling_b200.engine import GibbsEngine, make_params
from gibbssampling_b200.synthetic import planted_motif_set, background_of
for (n,L,k,chains,reps) in [(1000,500,12,1024,5),(1000,500,12,296,2),(1000,500,12,148,2),(1000,500,12,32,2),(10000,1000,16,64,1),(10000,1000,16,512,1),(100000,200,20,8,1)]:
    ps = planted_motif_set(n, L, k); bg = background_of(ps.ascii, 1e-4, 5)
    eng = GibbsEngine(ps.sequences()); p = make_params(k, 1e-4, 5, bg)
    for rep in range(reps):
        r = eng.run(p, chains, seed=1+rep, want_sites=False, want_scores=False, want_counts=False); st=r.stats
        print(n,L,k,chains,"full kernel_ms",round(st['kernel_ms'],3),"win/s %.3e"%(st['window_scores']/(st['kernel_ms']*1e-3)),"launches",st['kernel_launches'],"team",st['team_warps'],"spec",st['speculative_discards'],flush=True)
    eng.close()
PY
cat gpurun_out/probe6.log
