#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python - > gpurun_out/probe.log 2>&1 <<'PY'
import numpy as np
from gibbssampling_b200.engine import GibbsEngine, make_params
from gibbssampling_b200.synthetic import planted_motif_set, background_of
for (n,L,k,chains) in [(1000,500,12,1024),(1000,500,12,2048),(1000,500,12,400)]:
    ps = planted_motif_set(n, L, k); bg = background_of(ps.ascii, 1e-4, 5)
    eng = GibbsEngine(ps.sequences()); p = make_params(k, 1e-4, 5, bg)
    ref=None
    for team in (4,0):
        eng.set_team_warps(team)
        for rep in range(3):
            r = eng.run(p, chains, seed=1+rep, want_sites=True, want_scores=False, want_counts=False); st=r.stats
            if team==4 and rep==0: ref=r.sites.copy()
            if team==0 and rep==0: print(" same sites as forced team 4:", np.array_equal(ref, r.sites))
            print(n,L,k,chains,"team",team,"->",st['team_warps'],"launches",st['kernel_launches'],"kernel_ms",round(st['kernel_ms'],3),"win/s %.3e"%(st['window_scores']/(st['kernel_ms']*1e-3)),"upd",st['site_updates'],"sweeps",st['sweeps'],flush=True)
    eng.close()
PY
