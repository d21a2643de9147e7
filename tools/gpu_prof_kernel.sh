#!/bin/bash
# usage: gpu_prof_kernel.sh <tag> <kernel regex> <launch-skip> <python script + args...>
# one `ncu --set full` capture of the chosen kernel; raw and SASS pages come back as CSV (the report is > 64 MiB)
mkdir -p gpurun_out
tag=$1; kre=$2; skip=$3; shift 3
export PYTHONPATH=$PWD
timeout 300 python "$@" > gpurun_out/${tag}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${tag}_plain.log; exit 1; }
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:$kre -s $skip -c 1 -o /tmp/prof_$tag -f python "$@" > gpurun_out/${tag}_ncu.log 2>&1
echo "rc=$?" >> gpurun_out/${tag}_ncu.log
ncu -i /tmp/prof_$tag.ncu-rep --page raw --csv > gpurun_out/${tag}_raw.csv 2>/dev/null
ncu -i /tmp/prof_$tag.ncu-rep --page source --csv --print-source sass > gpurun_out/${tag}_sass.csv 2>/dev/null
ls -la gpurun_out | grep $tag
