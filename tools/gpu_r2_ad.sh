#!/bin/bash
# one-warp stages for the first greedy sweeps (GIBBS_OPT_SEQ_SWEEPS): parity, then the C2 bench line for 0 / 1 / 2 / 3 such sweeps
export PYTHONPATH=$PWD
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_seq_stages.py tests/test_gpu_fullsize.py tests/test_gpu_parity.py tests/test_gpu_race.py tests/test_gpu_restart_select.py -m gpu -x -q 2>&1 | tail -6
for rep in 1 2; do for s in 2 0 1 3; do
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-families --opt seq_sweeps=$s > gpurun_out/ad_bench_s${s}_$rep.json 2> gpurun_out/ad_bench_s${s}_$rep.err; python -c "
import json; d=json.load(open('gpurun_out/ad_bench_s${s}_$rep.json')); print('seq_sweeps $s', d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['ms_per_step'], d['gpu_launches'])"; done; done
timeout 300 python bench.py --config C3 --steps 2 --warmup 1 --no-cpu --no-families > gpurun_out/ad_bench_C3.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/ad_bench_C3.json')); print('C3', d['value'], d['ms_per_step'], d['roofline']['frac'])"
for s in 0 1 3; do timeout 300 python bench.py --config C3 --steps 2 --warmup 1 --no-cpu --no-families --opt seq_sweeps=$s > gpurun_out/ad_bench_C3_s$s.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/ad_bench_C3_s$s.json')); print('C3 s$s', d['value'], d['ms_per_step'], d['roofline']['frac'])"; done
