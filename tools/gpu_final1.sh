#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "ref rc=$?" >> gpurun_out/bench_ref.err
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/bench_short.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?" >> gpurun_out/ncu_launches.log
timeout 300 python tools/prof_probe.py 1024 0 > gpurun_out/probe_plain.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:chain_kernel -c 2 -o gpurun_out/prof_final -f python tools/prof_probe.py 1024 0 > gpurun_out/ncu_final.log 2>&1
echo "ncu rc=$?" >> gpurun_out/ncu_final.log
