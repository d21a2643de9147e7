import sys
sys.path.insert(0, '/root/repo')
from gibbssampling_b200.engine import GibbsEngine, make_params
from gibbssampling_b200.synthetic import planted_motif_set
from gibbssampling_b200 import _abi
for (n, L, k, chains) in [(10000, 1000, 16, 64), (100000, 200, 20, 8)]:
    ps = planted_motif_set(n, L, k)
    eng = GibbsEngine(ps.sequences())
    p = make_params(k, 1e-4, 5, [0.25] * 4, background=_abi.GIBBS_BG_DATA)
    r = eng.run(p, chains, seed=1, want_sites=False, want_scores=False, want_counts=False); st = r.stats
    print(n, L, k, chains, "data-derived full kernel_ms %.1f" % st["kernel_ms"], "win/s %.3e" % (st["window_scores"] / (st["kernel_ms"] * 1e-3)), "launches", st["kernel_launches"], flush=True)
    eng.close()
