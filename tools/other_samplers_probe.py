"""Throughput of the kernels beside the benchmarked one (C2 shape): data-derived SiteSampler, MotifSampler."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gibbssampling_b200.engine import GibbsEngine, make_params
from gibbssampling_b200.synthetic import planted_motif_set, background_of
from gibbssampling_b200 import _abi
n, L, k = 1000, 500, 12
chains = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
ps = planted_motif_set(n, L, k); bg = background_of(ps.ascii, 1e-4, 5)
eng = GibbsEngine(ps.sequences())
cases = [("site/fixed", make_params(k, 1e-4, 5, bg)),
         ("site/data ", make_params(k, 1e-4, 5, bg, background=_abi.GIBBS_BG_DATA)),
         ("motif/fixed", make_params(k, 1e-4, 5, bg, sampler=_abi.GIBBS_MOTIF_SAMPLER, cutoff=0.0)),
         ("motif/data ", make_params(k, 1e-4, 5, bg, sampler=_abi.GIBBS_MOTIF_SAMPLER, cutoff=0.0, background=_abi.GIBBS_BG_DATA))]
for name, p in cases:
    for rep in range(2):
        r = eng.run(p, chains, seed=1 + rep, want_sites=False, want_scores=False, want_counts=False); st = r.stats
        print(name, "chains", chains, "kernel_ms %.2f" % st["kernel_ms"], "win/s %.3e" % (st["window_scores"] / (st["kernel_ms"] * 1e-3)),
              "upd/s %.3e" % (st["site_updates"] / (st["kernel_ms"] * 1e-3)), "sweeps/chain %.1f" % (st["sweeps"] / chains), flush=True)
