"""How long the first (sequential) greedy sweeps of a C2 step take with one warp per chain (128 registers, no team
barriers) against the four-warp teams (72 registers): is a one-warp stage for sweeps 0-1 worth a hand-over?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gibbssampling_b200.engine import GibbsEngine, make_params
from gibbssampling_b200.synthetic import planted_motif_set, background_of
from gibbssampling_b200 import _abi
n, L, k, chains = 1000, 500, 12, 1024
ps = planted_motif_set(n, L, k); bg = background_of(ps.ascii, 1e-4, 5)
eng = GibbsEngine(ps.sequences())
def run(mask, max_sweeps, team, seed=1):
    eng.set_team_warps(team)
    p = make_params(k, 1e-4, 5, bg, phase_mask=mask, max_sweeps=max_sweeps)
    best = None
    for rep in range(3):
        r = eng.run(p, chains, seed=seed, want_scores=False, want_counts=False)
        if best is None or r.stats["kernel_ms"] < best[0]: best = (r.stats["kernel_ms"], r.stats["site_updates"], int(r.sites.sum()))
    return best
init = run(_abi.PHASE_INIT, 0, 0)
print("init only", init)
for team in (1, 4):
    for ms in (1, 2, 3, 4):
        t = run(_abi.PHASE_INIT | _abi.PHASE_GREEDY, ms, team)
        print(f"team {team}: init + {ms} greedy sweep(s): {t[0]:.3f} ms  -> sweeps alone {t[0] - init[0]:.3f} ms  updates {t[1]} checksum {t[2]}")
eng.close()
