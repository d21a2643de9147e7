"""Where the end-to-end step of bench.py spends host time (C2)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gibbssampling_b200.engine import GibbsEngine, make_params
from gibbssampling_b200.synthetic import planted_motif_set, background_of
n, L, k, chains = 1000, 500, 12, 1024
ps = planted_motif_set(n, L, k); bg = background_of(ps.ascii, 1e-4, 5)
p = make_params(k, 1e-4, 5, bg)
eng = GibbsEngine(ps.sequences())
host_ascii = torch.empty(ps.ascii.size, dtype=torch.uint8, pin_memory=True); host_ascii.numpy()[:] = ps.ascii
host_off = torch.empty(ps.offsets.size, dtype=torch.int64, pin_memory=True); host_off.numpy()[:] = ps.offsets
reps = chains - 1
for s in range(8):
    t = [time.perf_counter()]
    eng.upload_flat(host_ascii.numpy(), host_off.numpy()); t.append(time.perf_counter())      # H2D + pack + symbol check (one sync)
    eng.run_device(p, chains, seed=s); t.append(time.perf_counter())                          # launches only
    best = eng.fetch_best(reps, pinned=True); t.append(time.perf_counter())                   # sync + the winner's rows
    d = [1e3 * (b - a) for a, b in zip(t, t[1:])]
    print("upload %.3f  launch %.3f  wait+fetch_best %.3f  total %.3f  kernel_ms %.3f  host overhead %.3f" %
          (*d, sum(d), best.stats["kernel_ms"], sum(d) - best.stats["kernel_ms"]))
