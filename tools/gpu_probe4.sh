#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/probe.log
for f in 1 2 3; do
GIBBS_PAUSE_FACTOR=$f timeout 300 python - >> gpurun_out/probe.log 2>&1 <<'PY'
import os
from gibbssampling_b200.engine import GibbsEngine, make_params
from gibbssampling_b200.synthetic import planted_motif_set, background_of
n,L,k,chains=1000,500,12,1024
ps = planted_motif_set(n, L, k); bg = background_of(ps.ascii, 1e-4, 5)
eng = GibbsEngine(ps.sequences()); p = make_params(k, 1e-4, 5, bg)
for rep in range(4):
    r = eng.run(p, chains, seed=1+rep, want_sites=False, want_scores=False, want_counts=False); st=r.stats
    print("factor",os.environ["GIBBS_PAUSE_FACTOR"],"kernel_ms",round(st['kernel_ms'],3),"win/s %.3e"%(st['window_scores']/(st['kernel_ms']*1e-3)),flush=True)
PY
done
