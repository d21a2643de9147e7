#!/bin/bash
mkdir -p gpurun_out
for c in C1 C3 C4; do
  timeout 1200 python bench.py --config $c --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_$c.log 2> gpurun_out/bench_$c.err; echo "rc=$?" >> gpurun_out/bench_$c.err
done
