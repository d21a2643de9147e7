#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/b_pytest.log
tail -n 4 gpurun_out/b_pytest.log
timeout 600 python tools/init_paths_probe.py > gpurun_out/b_init_paths.log 2>&1; echo "rc=$?" >> gpurun_out/b_init_paths.log
cat gpurun_out/b_init_paths.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/b_bench.json 2> gpurun_out/b_bench.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/b_bench.json')); print(d['ms_per_step'], d['roofline']['frac'], d['e2e']['ms_per_step']); print({k:(round(v['ms_per_step'],1), round(v['frac'],3)) for k,v in d['families'].items()})"
bash tools/gpu_prof_kernel.sh b_ism init_smem_kernel 0 tools/prof_probe.py 1024 0
bash tools/gpu_prof_kernel.sh b_t4 'chain_kernel<6, 4' 0 tools/prof_probe.py 1024 0
