#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/g_pytest.log
tail -n 5 gpurun_out/g_pytest.log
timeout 300 python tools/cluster_probe.py
timeout 300 python tools/init_paths_probe.py c2 | grep smem
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/g_bench.json 2> gpurun_out/g_bench.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/g_bench.json')); print(d['ms_per_step'], d['roofline']['frac'], d['e2e']['ms_per_step']); print({k:(round(v['ms_per_step'],1), round(v['frac'],3)) for k,v in d['families'].items()})"
