"""Small fixed workload for ncu: C2-shaped input, MotifSampler restarts. usage: prof_motif.py <chains> <data 0|1>"""
import sys
from gibbssampling_b200 import _abi
from gibbssampling_b200.engine import GibbsEngine, make_params
from gibbssampling_b200.synthetic import planted_motif_set, background_of
n, L, k = 1000, 500, 12
chains = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
data = int(sys.argv[2]) if len(sys.argv) > 2 else 0
ps = planted_motif_set(n, L, k)
bg = background_of(ps.ascii, 1e-4, 5)
eng = GibbsEngine(ps.sequences())
p = make_params(k, 1e-4, 5, bg, sampler=_abi.GIBBS_MOTIF_SAMPLER, cutoff=0.0,
                background=_abi.GIBBS_BG_DATA if data else _abi.GIBBS_BG_FIXED)
for ph in (_abi.PHASE_INIT, _abi.PHASE_INIT | _abi.PHASE_STOCHASTIC, 0):
    p.phase_mask = ph
    r = eng.run(p, chains, seed=1, want_sites=False, want_scores=False, want_counts=False)
    print("phase_mask", ph, {k_: r.stats[k_] for k_ in ("site_updates", "sweeps", "exact_rescans", "speculative_discards", "kernel_ms", "team_warps")}, flush=True)
eng.close()
