"""Random starts on every init path (gibbs_set_option GIBBS_OPT_INIT_PATH): INIT-only and whole-run kernel times."""
import sys
from gibbssampling_b200 import _abi
from gibbssampling_b200.engine import GibbsEngine, make_params
from gibbssampling_b200.synthetic import planted_motif_set, background_of

shapes = [(1000, 500, 12, 1024, 3), (10000, 1000, 16, 64, 2), (100000, 200, 20, 8, 1)]
if len(sys.argv) > 1 and sys.argv[1] == "c2":
    shapes = shapes[:1]
names = {1: "chain", 2: "wide", 3: "smem"}
for (n, L, k, chains, reps) in shapes:
    ps = planted_motif_set(n, L, k)
    bg = background_of(ps.ascii, 1e-4, 5)
    eng = GibbsEngine(ps.sequences())
    pi = make_params(k, 1e-4, 5, bg, phase_mask=_abi.PHASE_INIT)
    pf = make_params(k, 1e-4, 5, bg)
    for path in (1, 2, 3):
        if path == 1 and n >= 10000:
            continue   # minutes on the chain's own team
        eng.set_option(_abi.GIBBS_OPT_INIT_PATH, path)
        for rep in range(reps):
            r = eng.run(pi, chains, seed=1 + rep, want_sites=False, want_scores=False, want_counts=False)
            st = r.stats
            print(n, L, k, chains, "INIT", names[path], "->", names.get(st["init_path"]), "kernel_ms", round(st["kernel_ms"], 3),
                  "draws/s %.3e" % (st["site_updates"] * (n - 1) / (st["kernel_ms"] * 1e-3)), flush=True)
        if n <= 1000:
            for rep in range(reps):
                r = eng.run(pf, chains, seed=1 + rep, want_sites=False, want_scores=False, want_counts=False)
                st = r.stats
                print(n, L, k, chains, "FULL", names[path], "->", names.get(st["init_path"]), "kernel_ms", round(st["kernel_ms"], 3),
                      "win/s %.3e" % (st["window_scores"] / (st["kernel_ms"] * 1e-3)), flush=True)
    eng.close()
