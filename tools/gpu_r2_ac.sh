#!/bin/bash
# plain-block path in random_draw_loop (init_smem / init_kernel / in-chain random starts): parity of every init path,
# init-only timings, the C2 bench line, the MotifSampler on tiled random starts, where the e2e step spends host time
export PYTHONPATH=$PWD
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_drift.py tests/test_gpu_symbols.py tests/test_gpu_motif.py tests/test_gpu_fuzz.py -m gpu -x -q 2>&1 | tail -4
timeout 600 python - <<'PY'
from gibbssampling_b200.engine import GibbsEngine, make_params
from gibbssampling_b200.synthetic import planted_motif_set, background_of
from gibbssampling_b200 import _abi
for (n,L,k,chains,reps) in [(1000,500,12,1024,3),(10000,1000,16,64,2)]:
    ps = planted_motif_set(n, L, k); bg = background_of(ps.ascii, 1e-4, 5)
    eng = GibbsEngine(ps.sequences())
    pi = make_params(k, 1e-4, 5, bg, phase_mask=_abi.PHASE_INIT)
    for path in (_abi.GIBBS_INIT_SMEM, _abi.GIBBS_INIT_WIDE, _abi.GIBBS_INIT_CHAIN):
        eng.set_option(_abi.GIBBS_OPT_INIT_PATH, path)
        for rep in range(reps):
            r = eng.run(pi, chains, seed=1+rep, want_sites=False, want_scores=False, want_counts=False); st=r.stats
            print(n,L,k,chains,"asked",path,"path",st['init_path'],"kernel_ms",round(st['kernel_ms'],3),"draws/s %.3e"%(st['site_updates']*(n-1)/(st['kernel_ms']*1e-3)),flush=True)
    eng.close()
PY
for rep in 1 2; do timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-families > gpurun_out/ac_bench_$rep.json 2> gpurun_out/ac_bench_$rep.err; python -c "
import json; d=json.load(open('gpurun_out/ac_bench_$rep.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['ms_per_step'])"; done
timeout 200 python tools/e2e_probe.py
