#!/bin/bash
# tiled random starts after the shared hi word (16 < k <= 20) and the plain-block path: parity, init timings, C4 / C3 / C2 bench
export PYTHONPATH=$PWD
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_drift.py tests/test_gpu_fuzz.py -m gpu -x -q 2>&1 | tail -6
timeout 600 python - <<'PY'
from gibbssampling_b200.engine import GibbsEngine, make_params
from gibbssampling_b200.synthetic import planted_motif_set, background_of
from gibbssampling_b200 import _abi
for (n,L,k,chains,reps) in [(100000,200,20,8,2),(10000,1000,16,64,2),(1000,500,12,1024,2),(20000,300,18,32,2)]:
    ps = planted_motif_set(n, L, k); bg = background_of(ps.ascii, 1e-4, 5)
    eng = GibbsEngine(ps.sequences())
    pi = make_params(k, 1e-4, 5, bg, phase_mask=_abi.PHASE_INIT)
    for path in (_abi.GIBBS_INIT_TILED, _abi.GIBBS_INIT_WIDE, _abi.GIBBS_INIT_SMEM):
        eng.set_option(_abi.GIBBS_OPT_INIT_PATH, path)
        for rep in range(reps):
            r = eng.run(pi, chains, seed=1+rep, want_sites=False, want_scores=False, want_counts=False); st=r.stats
            print(n,L,k,chains,"asked",path,"path",st['init_path'],"kernel_ms",round(st['kernel_ms'],3),"draws/s %.3e"%(st['site_updates']*(n-1)/(st['kernel_ms']*1e-3)),flush=True)
    eng.close()
PY
show() { python - "$1" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print(sys.argv[1], d['value'], d['ms_per_step'], d['roofline']['frac'], d['measurement']['init_path'], d.get('e2e',{}).get('value'))
PY
}
timeout 600 python bench.py --config C4 --steps 2 --warmup 1 --no-cpu --no-families > gpurun_out/z_bench_C4.json 2> gpurun_out/z_bench_C4.err; show gpurun_out/z_bench_C4.json
timeout 600 python bench.py --config C3 --steps 2 --warmup 1 --no-cpu --no-families > gpurun_out/z_bench_C3.json 2> gpurun_out/z_bench_C3.err; show gpurun_out/z_bench_C3.json
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-families > gpurun_out/z_bench_C2.json 2> gpurun_out/z_bench_C2.err; show gpurun_out/z_bench_C2.json
