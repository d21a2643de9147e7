#!/bin/bash
export PYTHONPATH=$PWD
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_symbols.py tests/test_gpu_fuzz.py tests/test_gpu_motif.py -m gpu -q 2>&1 | tail -40 > gpurun_out/r_pytest.log; cat gpurun_out/r_pytest.log
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r_pytest_all.log; cat gpurun_out/r_pytest_all.log
timeout 300 python bench.py --family motif --steps 5 --warmup 3 --no-cpu --no-families > gpurun_out/r_bench_motif.json 2> gpurun_out/r_bench_motif.err; tail -c 300 gpurun_out/r_bench_motif.json
for t in 4 16; do timeout 120 python tools/prof_seq_probe.py $([ $t = 4 ] && echo 1024 || echo 148) $t 1 2>&1 | tail -5; done
timeout 120 python tools/prof_seq_probe.py 1024 4 2 2>&1 | tail -3
bash tools/gpu_prof_kernel.sh seq4 chain_kernel 0 tools/prof_seq_probe.py 1024 4 1
bash tools/gpu_prof_kernel.sh seq16 chain_kernel 0 tools/prof_seq_probe.py 148 16 1
