#!/bin/bash
# usage: gpu_prof_chain.sh <tag> <chains> <team> [launch-skip]; csv pages of the chain kernel come back (the report is > 64 MiB)
mkdir -p gpurun_out
tag=$1; chains=$2; team=$3; skip=${4:-0}
timeout 300 python tools/prof_probe.py $chains $team > gpurun_out/probe_plain_$tag.log 2>&1 || exit 1
timeout 1200 ncu --set full --clock-control none -k regex:chain_kernel -s $skip -c 1 -o /tmp/prof_$tag -f python tools/prof_probe.py $chains $team > gpurun_out/ncu_$tag.log 2>&1
echo "rc=$?" >> gpurun_out/ncu_$tag.log
ncu -i /tmp/prof_$tag.ncu-rep --page raw --csv > gpurun_out/${tag}_raw.csv 2>/dev/null
ncu -i /tmp/prof_$tag.ncu-rep --page source --csv --print-source sass > gpurun_out/${tag}_sass.csv 2>/dev/null
ls -la gpurun_out | tail -n 6
