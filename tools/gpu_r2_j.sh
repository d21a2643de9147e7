#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 300 python tools/prof_motif.py 1024 0
timeout 300 python tools/prof_motif.py 1024 1
bash tools/gpu_prof_kernel.sh j_motif motif_kernel 2 tools/prof_motif.py 1024 0
