#!/bin/bash
export PYTHONPATH=$PWD
timeout 600 python -m pytest tests/test_gpu_motif2.py -x -q 2>&1 | tail -n 40
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 5
