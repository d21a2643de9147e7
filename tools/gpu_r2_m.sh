#!/bin/bash
export PYTHONPATH=$PWD
timeout 300 python tools/prof_motif.py 1024 0 | tail -1
timeout 300 python tools/prof_motif.py 1024 0 | tail -1
timeout 300 python -m pytest tests/test_gpu_motif.py tests/test_gpu_fuzz.py -x -q 2>&1 | tail -2
