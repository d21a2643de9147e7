"""The same question as seq_regime_probe.py for the other families on C2: first greedy sweeps with one warp per chain against
teams of four (data-derived SiteSampler; MotifSampler after its stochastic sweep)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gibbssampling_b200.engine import GibbsEngine, make_params
from gibbssampling_b200.synthetic import planted_motif_set, background_of
from gibbssampling_b200 import _abi
n, L, k, chains = 1000, 500, 12, 1024
ps = planted_motif_set(n, L, k); bg = background_of(ps.ascii, 1e-4, 5)
eng = GibbsEngine(ps.sequences())
eng.set_option(_abi.GIBBS_OPT_SEQ_SWEEPS, 0)
def run(team, **kw):
    eng.set_team_warps(team)
    p = make_params(k, 1e-4, 5, bg, **kw)
    best = None
    for rep in range(2):
        r = eng.run(p, chains, seed=1, want_scores=False, want_counts=False)
        if best is None or r.stats["kernel_ms"] < best[0]: best = (r.stats["kernel_ms"], r.stats["site_updates"], int(r.sites.sum()))
    return best
print("== data-derived SiteSampler")
init = run(0, background=_abi.GIBBS_BG_DATA, phase_mask=_abi.PHASE_INIT)
print("init only", init)
for team in (1, 4):
    for ms in (1, 2, 3):
        t = run(team, background=_abi.GIBBS_BG_DATA, phase_mask=_abi.PHASE_INIT | _abi.PHASE_GREEDY, max_sweeps=ms)
        print(f"team {team}: init + {ms} greedy sweep(s): {t[0]:.3f} ms -> sweeps alone {t[0] - init[0]:.3f} ms  checksum {t[2]}")
print("== MotifSampler, fixed background")
kw = dict(sampler=_abi.GIBBS_MOTIF_SAMPLER, cutoff=0.0)
for team in (1, 4):
    st = run(team, phase_mask=_abi.PHASE_INIT | _abi.PHASE_STOCHASTIC, **kw)
    print(f"team {team}: init + stochastic sweep {st[0]:.3f} ms")
    for ms in (1, 2, 3):
        t = run(team, phase_mask=_abi.PHASE_INIT | _abi.PHASE_STOCHASTIC | _abi.PHASE_MOTIF_GREEDY, max_sweeps=ms, **kw)
        print(f"team {team}: init + stochastic + {ms} greedy sweep(s): {t[0]:.3f} ms -> greedy alone {t[0] - st[0]:.3f} ms  checksum {t[2]}")
eng.close()
