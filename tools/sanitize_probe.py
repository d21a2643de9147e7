"""Small run of every kernel family for compute-sanitizer (memcheck / racecheck)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gibbssampling_b200 import _abi
from gibbssampling_b200.engine import GibbsEngine, make_params
from gibbssampling_b200.synthetic import planted_motif_set, background_of

for (n, L, Lmin, k) in [(24, 90, 50, 9), (10, 300, None, 16), (6, 40, None, 31)]:
    ps = planted_motif_set(n, L, k, seed=7, min_length=Lmin)
    bg = background_of(ps.ascii, 1e-4, 5)
    eng = GibbsEngine(ps.sequences())
    for team in (1, 4, 8, 0):
        eng.set_team_warps(team)
        r = eng.run(make_params(k, 1e-4, 5, bg), 6, seed=3)
        print("site fixed", n, L, k, "team", r.stats["team_warps"], r.stats["site_updates"], flush=True)
    eng.set_team_warps(0)
    sites = ps.truth.copy()
    eng.loo_counts(sites, 0, k); eng.window_scores(sites, 1, make_params(k, 1e-4, 5, bg)); eng.pick_argmax(sites, 2, make_params(k, 1e-4, 5, bg))
    r = eng.run(make_params(k, 1e-4, 5, bg, background=_abi.GIBBS_BG_DATA), 3, seed=3); print("site data", r.stats["site_updates"], flush=True)
    pm = make_params(k, 1e-4, 5, bg, cutoff=1.0, sampler=_abi.GIBBS_MOTIF_SAMPLER)
    r = eng.run(pm, 3, seed=3, want_counts=False); print("motif fixed", r.stats["site_updates"], flush=True)
    eng.pick_roulette(sites, 0, pm, 0.37)
    pm2 = make_params(k, 1e-4, 5, bg, cutoff=1.0, sampler=_abi.GIBBS_MOTIF_SAMPLER, background=_abi.GIBBS_BG_DATA)
    r = eng.run(pm2, 3, seed=3, want_counts=False); print("motif data", r.stats["site_updates"], flush=True)
    eng.close()
# many chains relative to the pause threshold: exercises the two-pass hand-over on a small problem
ps = planted_motif_set(16, 60, 8, seed=9); bg = background_of(ps.ascii, 1e-4, 5)
eng = GibbsEngine(ps.sequences())
r = eng.run(make_params(8, 1e-4, 5, bg), 400, seed=5); print("hand-over", r.stats, flush=True)
eng.close()
print("sanitize probe done")
