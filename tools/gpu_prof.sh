#!/bin/bash
# usage: gpu_prof.sh <chains> <team> <tag>
mkdir -p gpurun_out
timeout 300 python tools/prof_probe.py $1 $2 > gpurun_out/probe_plain.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:chain_kernel -c 1 -o gpurun_out/prof_$3 -f python tools/prof_probe.py $1 $2 > gpurun_out/ncu_$3.log 2>&1
echo "ncu rc=$?" >> gpurun_out/ncu_$3.log
