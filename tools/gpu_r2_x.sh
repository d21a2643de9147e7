#!/bin/bash
# tiled random starts (init_tiled_kernel): parity, then C4 / C3 bench A/B against the wide kernel
export PYTHONPATH=$PWD
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "random_starts" 2>&1 | tail -8 > gpurun_out/x_pytest_init.log; cat gpurun_out/x_pytest_init.log
timeout 900 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -8 > gpurun_out/x_pytest_full.log; cat gpurun_out/x_pytest_full.log
show() { python - "$1" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print(sys.argv[1], d['value'], d['ms_per_step'], d['roofline']['frac'], d['measurement']['init_path'], d.get('e2e',{}).get('value'))
PY
}
timeout 600 python bench.py --config C4 --steps 2 --warmup 1 --no-cpu --no-families > gpurun_out/x_bench_C4_tiled.json 2> gpurun_out/x_bench_C4_tiled.err; show gpurun_out/x_bench_C4_tiled.json
timeout 600 python bench.py --config C4 --steps 2 --warmup 1 --no-cpu --no-families --opt init_path=2 > gpurun_out/x_bench_C4_wide.json 2> gpurun_out/x_bench_C4_wide.err; show gpurun_out/x_bench_C4_wide.json
timeout 600 python bench.py --config C3 --steps 2 --warmup 1 --no-cpu --no-families > gpurun_out/x_bench_C3.json 2> gpurun_out/x_bench_C3.err; show gpurun_out/x_bench_C3.json
timeout 600 python bench.py --config C3 --steps 2 --warmup 1 --no-cpu --no-families --opt init_path=4 > gpurun_out/x_bench_C3_tiled.json 2> gpurun_out/x_bench_C3_tiled.err; show gpurun_out/x_bench_C3_tiled.json
tail -3 gpurun_out/x_bench_C4_tiled.err
