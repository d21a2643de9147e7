"""Where the end-to-end step of bench.py spends host time (C2)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gibbssampling_b200.engine import GibbsEngine, make_params
from gibbssampling_b200.synthetic import planted_motif_set, background_of
from gibbssampling_b200 import SiteSampler
n, L, k, chains = 1000, 500, 12, 1024
ps = planted_motif_set(n, L, k); bg = background_of(ps.ascii, 1e-4, 5)
p = make_params(k, 1e-4, 5, bg)
eng = GibbsEngine(ps.sequences())
host_ascii = torch.empty(ps.ascii.size, dtype=torch.uint8, pin_memory=True); host_ascii.numpy()[:] = ps.ascii
host_off = torch.empty(ps.offsets.size, dtype=torch.int64, pin_memory=True); host_off.numpy()[:] = ps.offsets
for s in range(6):
    t = [time.perf_counter()]
    eng.upload_flat(host_ascii.numpy(), host_off.numpy()); eng.synchronize(); t.append(time.perf_counter())
    eng.run_device(p, chains, seed=s); t.append(time.perf_counter())
    eng.synchronize(); t.append(time.perf_counter())
    r = eng.fetch(want_counts=False); t.append(time.perf_counter())
    best = SiteSampler.replay_restart_loop(chains - 1, r.scores, r.sites, r.sums); t.append(time.perf_counter())
    d = [1e3 * (b - a) for a, b in zip(t, t[1:])]
    print("upload %.2f  launch %.2f  wait %.2f  fetch %.2f  replay %.2f  total %.2f  kernel_ms %.2f" % (*d, sum(d), r.stats["kernel_ms"]))
