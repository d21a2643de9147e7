"""ncu workload for the random-start kernel: INIT phase only.
usage: prof_init.py <n> <L> <k> <chains>"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gibbssampling_b200.engine import GibbsEngine, make_params
from gibbssampling_b200.synthetic import planted_motif_set, background_of
from gibbssampling_b200 import _abi

n, L, k, chains = (int(x) for x in sys.argv[1:5])
ps = planted_motif_set(n, L, k)
bg = background_of(ps.ascii, 1e-4, 5)
eng = GibbsEngine(ps.sequences())
p = make_params(k, 1e-4, 5, bg, phase_mask=_abi.PHASE_INIT)
r = eng.run(p, chains, seed=1, want_sites=False, want_scores=False, want_counts=False)
print(r.stats)
eng.close()
