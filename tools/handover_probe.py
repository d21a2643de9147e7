"""Whole C2 restarts (1024 chains) under different straggler hand-over thresholds: mean kernel_ms over seeds."""
import sys
from gibbssampling_b200 import _abi
from gibbssampling_b200.engine import GibbsEngine, make_params
from gibbssampling_b200.synthetic import planted_motif_set, background_of
n, L, k, chains = 1000, 500, 12, 1024
ps = planted_motif_set(n, L, k)
bg = background_of(ps.ascii, 1e-4, 5)
eng = GibbsEngine(ps.sequences())
p = make_params(k, 1e-4, 5, bg)
configs = [tuple(int(x) for x in a.split(",")) for a in sys.argv[1:]] or [(2, 1)]
eng.run(p, chains, seed=99, want_sites=False, want_scores=False, want_counts=False)
for (s2, s3) in configs:
    eng.set_option(_abi.GIBBS_OPT_STAGE2_AT, s2)
    eng.set_option(_abi.GIBBS_OPT_STAGE3_AT, s3)
    ms = []
    for seed in range(8):
        st = eng.run(p, chains, seed=0xB200 + seed, want_sites=False, want_scores=False, want_counts=False).stats
        ms.append(st["kernel_ms"])
    print(f"stage2_at={s2} stage3_at={s3}: mean {sum(ms)/len(ms):.3f} ms  min {min(ms):.3f} max {max(ms):.3f}", flush=True)
eng.close()
