#!/bin/bash
export PYTHONPATH=$PWD
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_symbols.py tests/test_gpu_fuzz.py tests/test_gpu_motif.py -m gpu -q -x 2>&1 | tail -30 > gpurun_out/r_pytest.log; cat gpurun_out/r_pytest.log
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r_pytest_all.log; cat gpurun_out/r_pytest_all.log
timeout 300 python bench.py --family motif --steps 5 --warmup 3 --no-cpu --no-families > gpurun_out/r_bench_motif.json 2> gpurun_out/r_bench_motif.err; tail -c 600 gpurun_out/r_bench_motif.json
