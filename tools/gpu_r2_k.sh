#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/k_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/k_pytest.log
tail -n 4 gpurun_out/k_pytest.log
timeout 300 python tools/prof_motif.py 1024 0
timeout 300 python tools/prof_motif.py 1024 1
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/k_bench.json 2> gpurun_out/k_bench.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/k_bench.json')); print(d['ms_per_step'], d['roofline']['frac'], d['e2e']['ms_per_step']); print({k:(round(v['ms_per_step'],1), round(v['frac'],3)) for k,v in d['families'].items()})"
