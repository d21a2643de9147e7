#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python - > gpurun_out/probe.log 2>&1 <<'PY'
import time, numpy as np
from gibbssampling_b200.engine import GibbsEngine, make_params
from gibbssampling_b200.synthetic import planted_motif_set, background_of
for (n,L,k,chains) in [(1000,500,12,1024),(1000,500,12,148),(1000,500,12,4096),(20,100,8,1024)]:
    ps = planted_motif_set(n, L, k)
    bg = background_of(ps.ascii, 1e-4, 5)
    eng = GibbsEngine(ps.sequences())
    p = make_params(k, 1e-4, 5, bg)
    for team in (1,4,0):
        eng.set_team_warps(team)
        for rep in range(2):
            r = eng.run(p, chains, seed=1+rep, want_sites=False, want_scores=False, want_counts=False)
            st=r.stats
        print(n,L,k,chains,"team",team,"->",st['team_warps'],"kernel_ms",round(st['kernel_ms'],3),"updates",st['site_updates'],"sweeps",st['sweeps'],
              "win/s %.3e"%(st['window_scores']/(st['kernel_ms']*1e-3)),"upd/s %.3e"%(st['site_updates']/(st['kernel_ms']*1e-3)),
              "rescans",st['exact_rescans'], flush=True)
    eng.close()
PY
echo "probe rc=$?" >> gpurun_out/probe.log
