#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c_pytest.log
tail -n 4 gpurun_out/c_pytest.log
for v in base ism32 ism16 ism24nb2; do
  lib=$PWD/variants/lib_$v.so; [ $v = base ] && lib=$PWD/gibbssampling_b200/libgibbs_b200.so
  echo "== $v"; GIBBS_B200_LIB=$lib timeout 300 python tools/init_paths_probe.py c2 2>&1 | grep -E "smem|rc="
done > gpurun_out/c_variants.log 2>&1
cat gpurun_out/c_variants.log
timeout 600 python tools/init_paths_probe.py > gpurun_out/c_init_paths.log 2>&1; grep -v "FULL" gpurun_out/c_init_paths.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/c_bench.json 2> gpurun_out/c_bench.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/c_bench.json')); print(d['ms_per_step'], d['roofline']['frac'], d['e2e']['ms_per_step']); print({k:(round(v['ms_per_step'],1), round(v['frac'],3)) for k,v in d['families'].items()})"
bash tools/gpu_prof_kernel.sh c_ism init_smem_kernel 0 tools/prof_probe.py 1024 0
bash tools/gpu_prof_kernel.sh c_t4 "chain_kernel<6,.4" 0 tools/prof_probe.py 1024 0
