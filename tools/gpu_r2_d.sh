#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=$PWD
for v in base iw_mb3 iw_mb4 iw_mb3nb4 iw_mb4nb1; do
  lib=$PWD/variants/lib_$v.so; [ $v = base ] && lib=$PWD/gibbssampling_b200/libgibbs_b200.so
  echo "== $v"; GIBBS_B200_LIB=$lib timeout 300 python tools/init_wide_probe.py 2>&1 | awk 'NR%2==0'
done
