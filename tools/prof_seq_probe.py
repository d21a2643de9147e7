"""The latency-bound regime of the chain kernel for ncu: C2-shaped input, random starts, then ONE greedy sweep (almost every
update of the first sweep moves a site, so its rounds commit one update each) -- per-update latency, not throughput.
usage: prof_seq_probe.py <chains> <team 0|4|8|16> [max_sweeps]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gibbssampling_b200 import _abi
from gibbssampling_b200.engine import GibbsEngine, make_params
from gibbssampling_b200.synthetic import planted_motif_set, background_of

n, L, k = 1000, 500, 12
chains = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
team = int(sys.argv[2]) if len(sys.argv) > 2 else 4
max_sweeps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
ps = planted_motif_set(n, L, k)
bg = background_of(ps.ascii, 1e-4, 5)
eng = GibbsEngine(ps.sequences())
eng.set_team_warps(team)
p = make_params(k, 1e-4, 5, bg, phase_mask=_abi.PHASE_INIT | _abi.PHASE_GREEDY, max_sweeps=max_sweeps)
for it in range(3):
    r = eng.run(p, chains, seed=1)
    print({f: r.stats[f] for f in ("site_updates", "sweeps", "speculative_discards", "kernel_ms", "team_warps", "kernel_launches")})
p0 = make_params(k, 1e-4, 5, bg, phase_mask=_abi.PHASE_INIT)
r0 = eng.run(p0, chains, seed=1)
print("random starts alone:", r0.stats["kernel_ms"])
eng.close()
