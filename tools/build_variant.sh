#!/bin/bash
# usage: build_variant.sh <name> <-D flags...>   -> variants/lib_<name>.so
name=$1; shift
nvcc --split-compile 0 -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC -shared -cudart static "$@" -o variants/lib_$name.so gibbssampling_b200/csrc/gibbs_api.cu gibbssampling_b200/csrc/gibbs_drift_launch.cu
