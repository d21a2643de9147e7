#!/bin/bash
mkdir -p gpurun_out
timeout 600 python - > gpurun_out/probe.log 2>&1 <<'PY'
from gibbssampling_b200.engine import GibbsEngine, make_params
from gibbssampling_b200.synthetic import planted_motif_set, background_of
for (n,L,k,chains) in [(1000,500,12,1024),(1000,500,12,148),(1000,500,12,1)]:
    ps = planted_motif_set(n, L, k); bg = background_of(ps.ascii, 1e-4, 5)
    eng = GibbsEngine(ps.sequences()); p = make_params(k, 1e-4, 5, bg)
    for team in (4,8,32):
        eng.set_team_warps(team)
        for rep in range(2):
            r = eng.run(p, chains, seed=1+rep, want_sites=False, want_scores=False, want_counts=False); st=r.stats
            print(n,L,k,chains,"team",st['team_warps'],"kernel_ms",round(st['kernel_ms'],3),"win/s %.3e"%(st['window_scores']/(st['kernel_ms']*1e-3)),"spec",st['speculative_discards'],"upd",st['site_updates'],"sweeps",st['sweeps'],flush=True)
    eng.close()
PY
