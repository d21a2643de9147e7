#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=$PWD
echo "== base"; timeout 300 python tools/handover_probe.py 2,1 3,1 2,2
echo "== t8b4"; GIBBS_B200_LIB=$PWD/variants/lib_t8b4.so timeout 300 python tools/handover_probe.py 2,1 3,1 4,1 4,2
echo "== t8b4t16b2"; GIBBS_B200_LIB=$PWD/variants/lib_t8b4t16b2.so timeout 300 python tools/handover_probe.py 2,1 4,1 4,2 3,2 2,2
