import sys, os
sys.path.insert(0, '/root/repo')
from gibbssampling_b200.engine import GibbsEngine, make_params
from gibbssampling_b200.synthetic import planted_motif_set, background_of
from gibbssampling_b200 import _abi
n, L, k = 1000, 500, 12
ps = planted_motif_set(n, L, k); bg = background_of(ps.ascii, 1e-4, 5)
eng = GibbsEngine(ps.sequences())
for data in (0, 1):
    base = dict(cutoff=0.0, sampler=_abi.GIBBS_MOTIF_SAMPLER, background=_abi.GIBBS_BG_DATA if data else _abi.GIBBS_BG_FIXED)
    full = eng.run(make_params(k, 1e-4, 5, bg, **base), 256, seed=1, want_counts=False)
    print("data" if data else "fixed", "full ms %.1f" % full.stats["kernel_ms"], "updates", full.stats["site_updates"], "slow", full.stats["exact_rescans"], "sweeps/chain %.1f" % (full.stats["sweeps"] / 256))
    for mask, name in ((_abi.PHASE_INIT, "init"), (_abi.PHASE_INIT | _abi.PHASE_STOCHASTIC, "init+stoch")):
        r = eng.run(make_params(k, 1e-4, 5, bg, phase_mask=mask, **base), 256, seed=1, want_counts=False)
        print("   ", name, "ms %.1f" % r.stats["kernel_ms"], "updates", r.stats["site_updates"], "slow", r.stats["exact_rescans"])
