"""One process, every GPU of the box (gibbs_multi_*): C2 with 1024 restarts per GPU, wall-clock per run against the same
restarts on ONE GPU, and the result of both (must be the same restart, same sites)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import bench
from gibbssampling_b200 import engine, _abi

from gibbssampling_b200.synthetic import background_of, planted_motif_set

n_seqs, length, k, chains, shifts = bench.CONFIGS["C2"]
ps = planted_motif_set(n_seqs, length, k, seed=bench.SEED)
seqs = ps.sequences()
bg = background_of(ps.ascii, bench.PSEUDOCOUNT, bench.ALPHABET_SIZE)
params = engine.make_params(k, bench.PSEUDOCOUNT, bench.ALPHABET_SIZE, bg, phase_shifts=shifts)
ndev = int(_abi.load().gibbs_device_count())
res = {}
for nd in sorted({1, min(2, ndev), ndev}):
    total = chains * nd
    with engine.MultiEngine(seqs, n_devices=nd) as m:
        for it in range(4):
            t0 = time.perf_counter()
            m.run_device(params, total, seed=bench.SEED)
            r = m.fetch_best(total - 1)
            dt = time.perf_counter() - t0
        print(f"devices {nd}: {total} restarts, wall {dt*1e3:.2f} ms per run (incl. fetch of the winner), kernel_ms {r.stats['kernel_ms']:.2f}, "
              f"window-scores {r.stats['window_scores']:.4g} -> {r.stats['window_scores']/dt:.4g}/s, restart {r.restart}, sum {r.total:.6f}")
        res[nd] = r
    if nd > 1:  # the same restarts on one device
        with engine.MultiEngine(seqs, n_devices=1) as m1:
            m1.run_device(params, total, seed=bench.SEED)
            r1 = m1.fetch_best(total - 1)
        same = r1.restart == r.restart and np.array_equal(r1.sites, r.sites) and np.array_equal(r1.scores, r.scores)
        print(f"  same result as one device running all {total} restarts: {same}")
        assert same
