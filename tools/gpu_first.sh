#!/bin/bash
# first GPU contact: smoke, parity tests, a timing probe
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python - > gpurun_out/probe.log 2>&1 <<'PY'
import time, numpy as np
from gibbssampling_b200.engine import GibbsEngine, make_params, measure_smem_bandwidth
from gibbssampling_b200.synthetic import planted_motif_set, background_of
print("smem GB/s, ms:", measure_smem_bandwidth(0, 20000))
for (n,L,k,chains) in [(20,100,8,1),(20,100,8,1024),(1000,500,12,64),(1000,500,12,1024)]:
    ps = planted_motif_set(n, L, k)
    bg = background_of(ps.ascii, 1e-4, 5)
    eng = GibbsEngine(ps.sequences())
    p = make_params(k, 1e-4, 5, bg)
    for rep in range(2):
        t=time.time(); r = eng.run(p, chains, seed=1+rep, want_sites=True); dt=time.time()-t
        st=r.stats
        hits=int((r.sites==ps.truth[None,:]).all(axis=1).sum())
        print(n,L,k,chains,"wall",round(dt,4),"kernel_ms",round(st['kernel_ms'],3),"updates",st['site_updates'],"sweeps",st['sweeps'],
              "win/s %.3e"%(st['window_scores']/(st['kernel_ms']*1e-3)),"upd/s %.3e"%(st['site_updates']/(st['kernel_ms']*1e-3)),
              "rescans",st['exact_rescans'],"fast",st['fast_path'],"hits",hits,"best",r.best_chain, flush=True)
    eng.close()
PY
echo "probe rc=$?" >> gpurun_out/probe.log
