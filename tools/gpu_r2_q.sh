#!/bin/bash
export PYTHONPATH=$PWD
timeout 300 python tools/init_wide_probe.py
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_drift.py tests/test_gpu_fuzz.py -x -q 2>&1 | tail -n 4
