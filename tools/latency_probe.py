"""Per-round latency of one chain: kernel time / rounds for each team size (C2 shape)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gibbssampling_b200.engine import GibbsEngine, make_params
from gibbssampling_b200.synthetic import planted_motif_set, background_of
from gibbssampling_b200 import _abi
n, L, k = 1000, 500, 12
ps = planted_motif_set(n, L, k); bg = background_of(ps.ascii, 1e-4, 5)
eng = GibbsEngine(ps.sequences())
full = make_params(k, 1e-4, 5, bg)
r0 = eng.run(full, 1, seed=3)
for team in (1, 4, 8, 16):
    eng.set_team_warps(team)
    for mask, name in ((_abi.PHASE_GREEDY, "greedy"), (_abi.PHASE_LEFT, "left"), (_abi.PHASE_GREEDY | _abi.PHASE_LEFT | _abi.PHASE_RIGHT, "g+l+r")):
        eng.set_start_state(r0.sites, r0.scores)           # converged state: every sweep is quiet, rounds are T wide
        p = make_params(k, 1e-4, 5, bg, phase_mask=mask)
        for rep in range(2):
            eng.set_start_state(r0.sites, r0.scores)
            r = eng.run(p, 1, seed=3, want_counts=False)
        st = r.stats
        rounds = st["site_updates"] / team
        print(f"T={team:2d} {name:6s} kernel_ms {st['kernel_ms']:.3f} sweeps {st['sweeps']} updates {st['site_updates']} "
              f"us/round {1e3 * st['kernel_ms'] / rounds:.2f} us/sweep {1e3 * st['kernel_ms'] / st['sweeps']:.1f}", flush=True)
