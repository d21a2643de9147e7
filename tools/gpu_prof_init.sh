#!/bin/bash
# usage: gpu_prof_init.sh <tag> <n> <L> <k> <chains>; the report itself stays on the box (> 64 MiB), its csv pages come back
mkdir -p gpurun_out
tag=$1; shift
timeout 300 python tools/prof_init.py "$@" > gpurun_out/init_plain_$tag.log 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none -k regex:init_kernel -c 1 -o /tmp/prof_init_$tag -f python tools/prof_init.py "$@" > gpurun_out/ncu_init_$tag.log 2>&1
echo "rc=$?" >> gpurun_out/ncu_init_$tag.log
ncu -i /tmp/prof_init_$tag.ncu-rep --page raw --csv > gpurun_out/init_${tag}_raw.csv 2>/dev/null
ncu -i /tmp/prof_init_$tag.ncu-rep --page source --csv --print-source sass > gpurun_out/init_${tag}_sass.csv 2>/dev/null
ls -la /tmp/prof_init_$tag.ncu-rep gpurun_out
