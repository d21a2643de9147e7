#!/bin/bash
mkdir -p gpurun_out
timeout 120 python - > gpurun_out/pool_first.log 2>&1 <<'PY'
from gibbssampling_b200.engine import GibbsEngine, make_params
from gibbssampling_b200.synthetic import planted_motif_set, background_of
ps = planted_motif_set(30, 120, 10, seed=21); bg = background_of(ps.ascii, 1e-4, 5)
eng = GibbsEngine(ps.sequences()); p = make_params(10, 1e-4, 5, bg)
eng.set_team_warps(4); a = eng.run(p, 16, seed=5)
eng.set_team_warps(32); b = eng.run(p, 16, seed=5)
print("pool == team:", a.sites.tolist() == b.sites.tolist(), a.scores.tobytes() == b.scores.tobytes(), a.stats, b.stats)
PY
echo "first rc=$?" >> gpurun_out/pool_first.log
grep -q "pool == team: True True" gpurun_out/pool_first.log || exit 0
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python - > gpurun_out/probe.log 2>&1 <<'PY'
from gibbssampling_b200.engine import GibbsEngine, make_params
from gibbssampling_b200.synthetic import planted_motif_set, background_of
for (n,L,k,chains) in [(1000,500,12,1024),(1000,500,12,2048),(1000,500,12,148),(1000,500,12,16)]:
    ps = planted_motif_set(n, L, k); bg = background_of(ps.ascii, 1e-4, 5)
    eng = GibbsEngine(ps.sequences()); p = make_params(k, 1e-4, 5, bg)
    for team in (4,32):
        eng.set_team_warps(team)
        for rep in range(3):
            r = eng.run(p, chains, seed=1+rep, want_sites=False, want_scores=False, want_counts=False); st=r.stats
            print(n,L,k,chains,"team",st['team_warps'],"kernel_ms",round(st['kernel_ms'],3),"win/s %.3e"%(st['window_scores']/(st['kernel_ms']*1e-3)),"spec",st['speculative_discards'],"upd",st['site_updates'],flush=True)
    eng.close()
PY
