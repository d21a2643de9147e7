#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=$PWD
bash tools/gpu_prof_kernel.sh l_motif motif_kernel 2 tools/prof_motif.py 1024 0
