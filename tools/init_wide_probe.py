"""Grid-wide random starts with global gathers (init_kernel) on the C3 / C4 shapes: INIT-only kernel time."""
from gibbssampling_b200 import _abi
from gibbssampling_b200.engine import GibbsEngine, make_params
from gibbssampling_b200.synthetic import planted_motif_set, background_of
for (n, L, k, chains, reps) in [(1000, 500, 12, 1024, 2), (10000, 1000, 16, 64, 2), (100000, 200, 20, 8, 2)]:
    ps = planted_motif_set(n, L, k)
    bg = background_of(ps.ascii, 1e-4, 5)
    eng = GibbsEngine(ps.sequences())
    eng.set_option(_abi.GIBBS_OPT_INIT_PATH, _abi.GIBBS_INIT_WIDE)
    pi = make_params(k, 1e-4, 5, bg, phase_mask=_abi.PHASE_INIT)
    for rep in range(reps):
        st = eng.run(pi, chains, seed=1 + rep, want_sites=False, want_scores=False, want_counts=False).stats
        print(n, L, k, chains, "INIT wide kernel_ms", round(st["kernel_ms"], 3), "draws/s %.3e" % (st["site_updates"] * (n - 1) / (st["kernel_ms"] * 1e-3)), flush=True)
    eng.close()
