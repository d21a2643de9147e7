#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 600 python -m pytest tests/test_gpu_cluster.py -x -q 2>&1 | tail -n 30
timeout 300 python tools/cluster_probe.py
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/f_pytest.log
tail -n 5 gpurun_out/f_pytest.log
