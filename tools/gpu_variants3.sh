#!/bin/bash
# compare builds in variants/ (plus the in-tree build as "base") on the data-derived SiteSampler (C2 shape)
mkdir -p gpurun_out
for v in base "$@"; do
lib=$PWD/variants/lib_$v.so; [ $v = base ] && lib=$PWD/gibbssampling_b200/libgibbs_b200.so
echo "== $v"; GIBBS_B200_LIB=$lib timeout 300 python tools/other_samplers_probe.py 1024 2>&1 | sed -n 3,4p
done
