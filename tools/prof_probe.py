"""Small fixed workload for ncu: C2-shaped input, a few chains, one launch of the chain kernel.
usage: prof_probe.py <chains> <team 0|1|4> [chain_id_base]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gibbssampling_b200.engine import GibbsEngine, make_params
from gibbssampling_b200.synthetic import planted_motif_set, background_of

n, L, k = 1000, 500, 12
chains = int(sys.argv[1]) if len(sys.argv) > 1 else 148
team = int(sys.argv[2]) if len(sys.argv) > 2 else 0
base = int(sys.argv[3]) if len(sys.argv) > 3 else 0
ps = planted_motif_set(n, L, k)
bg = background_of(ps.ascii, 1e-4, 5)
eng = GibbsEngine(ps.sequences())
eng.set_team_warps(team)
p = make_params(k, 1e-4, 5, bg)
r = eng.run(p, chains, seed=1, chain_id_base=base)
print(r.stats)
eng.close()
