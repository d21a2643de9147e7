#!/bin/bash
# N GPUs of one box: the multi-device C ABI test, then bench.py under torch.distributed.run
mkdir -p gpurun_out
export PYTHONPATH=$PWD
N=${1:-2}
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -n 8
timeout 600 python -m pytest tests/test_gpu_restart_select.py -x -q 2>&1 | tail -n 3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/n_bench_n$N.json 2> gpurun_out/n_bench_n$N.err; echo "bench rc=$?"
tail -n 3 gpurun_out/n_bench_n$N.err
python -c "
import json,sys; d=json.load(open('gpurun_out/n_bench_n$N.json')); print('N', d['n_gpus'], 'ms', d['ms_per_step'], 'value %.4g' % d['value'], 'e2e ms', d['e2e']['ms_per_step']); print(d['per_rank']); print(d['clocks'])"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-families > gpurun_out/n_bench_n1.json 2>/dev/null
python -c "
import json,sys; d=json.load(open('gpurun_out/n_bench_n1.json')); print('N', d['n_gpus'], 'ms', d['ms_per_step'], 'value %.4g' % d['value'], 'e2e ms', d['e2e']['ms_per_step']); print(d['per_rank'])"
