"""Is the slower rank of a multi-GPU run slower because of its GPU or because of its chains? Runs the chain sets of
ranks 0..3 (chain_id_base = rank * 1024, seeds of bench.py) on ONE GPU and prints the mean step time of each."""
from gibbssampling_b200.engine import GibbsEngine, make_params
from gibbssampling_b200.synthetic import planted_motif_set, background_of
n, L, k, chains, SEED = 1000, 500, 12, 1024, 0xB200
ps = planted_motif_set(n, L, k, seed=SEED)
bg = background_of(ps.ascii, 1e-4, 5)
eng = GibbsEngine(ps.sequences())
p = make_params(k, 1e-4, 5, bg)
eng.run(p, chains, seed=1, want_sites=False, want_scores=False, want_counts=False)
for rank in range(4):
    ms = [eng.run(p, chains, chain_id_base=rank * chains, seed=SEED + s, want_sites=False, want_scores=False, want_counts=False).stats["kernel_ms"] for s in range(10)]
    print(f"chain set of rank {rank}: mean {sum(ms)/len(ms):.3f} ms  median {sorted(ms)[5]:.3f}  min {min(ms):.3f}  max {max(ms):.3f}", flush=True)
eng.close()
