"""Where the MotifSampler step goes (C2 shape, 1024 restarts, fixed background): random starts alone, + the stochastic sweep,
+ greedy sweeps capped at 1, 2, 4 and uncapped; the run statistics of each.
usage: motif_phase_probe.py [chains]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gibbssampling_b200 import _abi
from gibbssampling_b200.engine import GibbsEngine, make_params
from gibbssampling_b200.synthetic import planted_motif_set, background_of

n, L, k = 1000, 500, 12
chains = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
ps = planted_motif_set(n, L, k, seed=0xB200)
bg = background_of(ps.ascii, 1e-4, 5)
eng = GibbsEngine(ps.sequences())
I, S, G = _abi.PHASE_INIT, _abi.PHASE_STOCHASTIC, _abi.PHASE_MOTIF_GREEDY
for name, mask, cap in (("init", I, 0), ("init+stoch", I | S, 0), ("init+stoch+greedy1", I | S | G, 1), ("greedy2", I | S | G, 2),
                        ("greedy4", I | S | G, 4), ("greedy8", I | S | G, 8), ("full", 0, 0)):
    p = make_params(k, 1e-4, 5, bg, sampler=_abi.GIBBS_MOTIF_SAMPLER, cutoff=0.0, phase_mask=mask, max_sweeps=cap)
    for it in range(2):
        r = eng.run(p, chains, seed=1, want_counts=False)
    st = r.stats
    print(f"{name:22s} kernel_ms {st['kernel_ms']:8.3f} updates {st['site_updates']:10d} sweeps/chain {st['sweeps']/chains:6.2f} "
          f"discards {st['speculative_discards']:9d} exact {st['exact_rescans']:9d} capped {st['capped_chains']}")
eng.close()
