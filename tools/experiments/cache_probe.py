"""Margin cache of the greedy sweeps on / off (GIBBS_OPT_MARGIN_CACHE): whole C2 restarts, mean kernel_ms over seeds; a library
built with -DGIBBS_CACHE_DEBUG_COUNT reports the cached updates in the capped_chains counter."""
import sys
from gibbssampling_b200 import _abi
from gibbssampling_b200.engine import GibbsEngine, make_params
from gibbssampling_b200.synthetic import planted_motif_set, background_of
n, L, k, chains = 1000, 500, 12, 1024
ps = planted_motif_set(n, L, k)
bg = background_of(ps.ascii, 1e-4, 5)
eng = GibbsEngine(ps.sequences())
p = make_params(k, 1e-4, 5, bg)
eng.run(p, chains, seed=99, want_sites=False, want_scores=False, want_counts=False)
for on in (1, 0, 1, 0):
    eng.set_option(_abi.GIBBS_OPT_MARGIN_CACHE, on)
    ms, upd, cached = [], 0, 0
    for seed in range(6):
        st = eng.run(p, chains, seed=0xB200 + seed, want_sites=False, want_scores=False, want_counts=False).stats
        ms.append(st["kernel_ms"]); upd += st["site_updates"]; cached += st["capped_chains"]
    print(f"cache={on}: mean {sum(ms)/len(ms):.3f} ms  min {min(ms):.3f} max {max(ms):.3f}  updates {upd}  cached(debug builds) {cached}", flush=True)
# one greedy phase alone from random starts: where the cached updates are
for on in (1, 0):
    eng.set_option(_abi.GIBBS_OPT_MARGIN_CACHE, on)
    for cap in (1, 2, 3, 4, 6, 0):
        q = make_params(k, 1e-4, 5, bg, phase_mask=_abi.PHASE_INIT | _abi.PHASE_GREEDY, max_sweeps=cap)
        st = eng.run(q, chains, seed=0xB200, want_sites=False, want_scores=False, want_counts=False).stats
        print(f"cache={on} greedy sweeps <= {cap}: {st['kernel_ms']:.3f} ms updates {st['site_updates']} capped/cached {st['capped_chains']}", flush=True)
eng.close()
