#!/bin/bash
# usage: gpu_prof_family.sh <family> <kernel regex> <chains>; csv pages of the first matching launch come back
mkdir -p gpurun_out
fam=$1; kern=$2; chains=$3
timeout 600 python tools/prof_family.py $fam $chains > gpurun_out/fam_plain_$fam.log 2>&1 || exit 1
timeout 1500 ncu --set full --clock-control none -k regex:$kern -c 1 -o /tmp/prof_$fam -f python tools/prof_family.py $fam $chains > gpurun_out/ncu_fam_$fam.log 2>&1
echo "rc=$?" >> gpurun_out/ncu_fam_$fam.log
ncu -i /tmp/prof_$fam.ncu-rep --page raw --csv > gpurun_out/fam_${fam}_raw.csv 2>/dev/null
ncu -i /tmp/prof_$fam.ncu-rep --page source --csv --print-source sass > gpurun_out/fam_${fam}_sass.csv 2>/dev/null
