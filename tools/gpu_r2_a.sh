#!/bin/bash
# round 2, first call: the whole GPU suite, the init paths, the default bench, a launch list
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/a_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/a_pytest.log
timeout 600 python tools/init_paths_probe.py > gpurun_out/a_init_paths.log 2>&1; echo "rc=$?" >> gpurun_out/a_init_paths.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err; echo "bench rc=$?" >> gpurun_out/a_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/a_bench_ref.json 2> gpurun_out/a_bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/a_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-families > gpurun_out/a_ncu.log 2>&1
tail -n 5 gpurun_out/a_pytest.log; cat gpurun_out/a_init_paths.log; cat gpurun_out/a_bench.json; tail -n 3 gpurun_out/a_bench.err
