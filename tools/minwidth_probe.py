"""Whole C2 restarts (1024 chains): narrowest speculative round of the greedy sweeps (GIBBS_OPT_MIN_WIDTH)."""
from gibbssampling_b200 import _abi
from gibbssampling_b200.engine import GibbsEngine, make_params
from gibbssampling_b200.synthetic import planted_motif_set, background_of
n, L, k, chains = 1000, 500, 12, 1024
ps = planted_motif_set(n, L, k)
bg = background_of(ps.ascii, 1e-4, 5)
eng = GibbsEngine(ps.sequences())
p = make_params(k, 1e-4, 5, bg)
eng.run(p, chains, seed=99, want_sites=False, want_scores=False, want_counts=False)
for mw in (1, 2, 3, 1, 2):
    eng.set_option(_abi.GIBBS_OPT_MIN_WIDTH, mw)
    ms, spec = [], []
    for seed in range(8):
        st = eng.run(p, chains, seed=0xB200 + seed, want_sites=False, want_scores=False, want_counts=False).stats
        ms.append(st["kernel_ms"]); spec.append(st["speculative_discards"] / st["site_updates"])
    print(f"min_width={mw}: mean {sum(ms)/len(ms):.3f} ms  min {min(ms):.3f} max {max(ms):.3f}  discarded/committed {sum(spec)/len(spec):.3f}", flush=True)
eng.close()
