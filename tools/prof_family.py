"""ncu workload for the kernels beside the benchmarked one (C2 shape). usage: prof_family.py <data|motif|motif-data> <chains>"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gibbssampling_b200.engine import GibbsEngine, make_params
from gibbssampling_b200.synthetic import planted_motif_set, background_of
from gibbssampling_b200 import _abi
fam, chains = sys.argv[1], int(sys.argv[2])
n, L, k = 1000, 500, 12
ps = planted_motif_set(n, L, k); bg = background_of(ps.ascii, 1e-4, 5)
eng = GibbsEngine(ps.sequences())
p = make_params(k, 1e-4, 5, bg, cutoff=0.0,
                background=_abi.GIBBS_BG_DATA if fam in ("data", "motif-data") else _abi.GIBBS_BG_FIXED,
                sampler=_abi.GIBBS_MOTIF_SAMPLER if fam.startswith("motif") else _abi.GIBBS_SITE_SAMPLER)
r = eng.run(p, chains, seed=1, want_sites=False, want_scores=False, want_counts=False)
print(r.stats)
eng.close()
