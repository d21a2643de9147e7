#!/bin/bash
# C5 throughput sweep + profile of the MotifSampler kernel after the coarse-logarithm roulette
export PYTHONPATH=$PWD
mkdir -p gpurun_out
timeout 900 python tools/sweep_c5.py --out gpurun_out/sweep_c5.json > gpurun_out/w_sweep.log 2>&1; tail -16 gpurun_out/w_sweep.log
for fam in motif motif-data data; do timeout 300 python bench.py --family $fam --steps 5 --warmup 3 --no-cpu --no-families > gpurun_out/w_bench_$fam.json 2> gpurun_out/w_bench_$fam.err; python - gpurun_out/w_bench_$fam.json <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print(d['config']['family'], d['value'], d['ms_per_step'], d['roofline']['frac'], d['gpu_launches'])
PY
done
bash tools/gpu_prof_family.sh motif 'motif_kernel' 1024; tail -3 gpurun_out/ncu_fam_motif.log
ls -la gpurun_out | tail
