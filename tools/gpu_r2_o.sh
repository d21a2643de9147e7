#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/o_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/o_pytest.log
tail -n 3 gpurun_out/o_pytest.log
timeout 300 python tools/minwidth_probe.py | head -2
timeout 300 python tools/init_paths_probe.py c2 | grep "INIT smem" | tail -1
timeout 300 python tools/prof_motif.py 1024 0 | tail -1
