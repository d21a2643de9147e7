#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=$PWD
for v in base iw_mb3 iw_mb4 iw_nb4 iw_nb1; do
  lib=$PWD/variants/lib_$v.so; [ $v = base ] && lib=$PWD/gibbssampling_b200/libgibbs_b200.so
  echo "== $v"; GIBBS_B200_LIB=$lib timeout 300 python tools/init_wide_probe.py 2>&1 | tail -n 6
done > gpurun_out/d_variants.log 2>&1
cat gpurun_out/d_variants.log
bash tools/gpu_prof_kernel.sh d_t4 chain_kernel 0 tools/prof_probe.py 1024 0
