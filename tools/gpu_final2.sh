#!/bin/bash
# Round-1 evidence on the final build: bench lines, reference arm, launch list, ncu summaries (csv pages only).
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "ref rc=$?" >> gpurun_out/bench_ref.err
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
for c in C1 C3 C4; do
  timeout 1200 python bench.py --config $c --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_$c.log 2> gpurun_out/bench_$c.err; echo "rc=$?" >> gpurun_out/bench_$c.err
done
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/bench_short.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?" >> gpurun_out/ncu_launches.log
timeout 300 python tools/prof_probe.py 1024 0 > gpurun_out/probe_plain.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none -k regex:chain_kernel -c 3 -o /tmp/prof_all -f python tools/prof_probe.py 1024 0 > gpurun_out/ncu_all.log 2>&1
echo "ncu rc=$?" >> gpurun_out/ncu_all.log
ncu -i /tmp/prof_all.ncu-rep --page raw --csv > gpurun_out/final_raw.csv 2>/dev/null
timeout 1200 ncu --set full --clock-control none -k regex:chain_kernel -c 1 -o /tmp/prof_one -f python tools/prof_probe.py 1024 0 > gpurun_out/ncu_one.log 2>&1
ncu -i /tmp/prof_one.ncu-rep --page source --csv --print-source sass > gpurun_out/final_sass.csv 2>/dev/null
ls -la gpurun_out
