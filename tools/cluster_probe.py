"""Whole C2 restarts (1024 chains) with the cluster stages off / 4 / 8: mean kernel_ms over seeds."""
from gibbssampling_b200 import _abi
from gibbssampling_b200.engine import GibbsEngine, make_params
from gibbssampling_b200.synthetic import planted_motif_set, background_of
n, L, k, chains = 1000, 500, 12, 1024
ps = planted_motif_set(n, L, k)
bg = background_of(ps.ascii, 1e-4, 5)
eng = GibbsEngine(ps.sequences())
p = make_params(k, 1e-4, 5, bg)
eng.run(p, chains, seed=99, want_sites=False, want_scores=False, want_counts=False)
ref = None
for cluster in (0, 4, 8, 0, 8):
    eng.set_option(_abi.GIBBS_OPT_CLUSTER, cluster)
    ms = []
    for seed in range(8):
        r = eng.run(p, chains, seed=0xB200 + seed, want_counts=False)
        ms.append(r.stats["kernel_ms"])
        if seed == 0:
            if ref is None:
                ref = (r.sites.tobytes(), r.scores.tobytes())
            else:
                assert ref == (r.sites.tobytes(), r.scores.tobytes()), "cluster stage changed a result"
    print(f"cluster={cluster}: mean {sum(ms)/len(ms):.3f} ms  min {min(ms):.3f} max {max(ms):.3f} launches {r.stats['kernel_launches']}", flush=True)
eng.close()
