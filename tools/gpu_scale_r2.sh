#!/bin/bash
# one 8-GPU box: the bench at N = 2, 4, 8 (weak scaling, 1024 chains per GPU), the reference arm under torchrun, the
# multi-device tests, C3 as BASELINE states it (8192 restarts over 8 GPUs) and the MultiEngine path
export PYTHONPATH=$PWD
mkdir -p gpurun_out
nvidia-smi -L | wc -l > gpurun_out/s_ngpu.txt
timeout 300 python -m pytest tests/test_gpu_restart_select.py -m gpu -x -q > gpurun_out/s_pytest_multi.log 2>&1; echo "rc=$?" >> gpurun_out/s_pytest_multi.log
port=29520
for N in 2 4 8; do
  port=$((port+1))
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --steps 10 --warmup 3 --no-families > gpurun_out/s_bench_n$N.json 2> gpurun_out/s_bench_n$N.err; echo "rc=$?" >> gpurun_out/s_bench_n$N.err
done
port=$((port+1))
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $port bench.py --impl reference --gpus 8 --steps 2 --warmup 1 > gpurun_out/s_bench_ref_n8.json 2> gpurun_out/s_bench_ref_n8.err; echo "rc=$?" >> gpurun_out/s_bench_ref_n8.err
port=$((port+1))
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $port bench.py --gpus 8 --config C3 --steps 2 --warmup 2 --no-cpu --no-families > gpurun_out/s_bench_n8_C3.json 2> gpurun_out/s_bench_n8_C3.err; echo "rc=$?" >> gpurun_out/s_bench_n8_C3.err
port=$((port+1))
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $port bench.py --gpus 8 --family motif --steps 5 --warmup 3 --no-cpu --no-families > gpurun_out/s_bench_n8_motif.json 2> gpurun_out/s_bench_n8_motif.err; echo "rc=$?" >> gpurun_out/s_bench_n8_motif.err
timeout 300 python tools/multi_engine_probe.py > gpurun_out/s_multi_engine.log 2>&1; echo "rc=$?" >> gpurun_out/s_multi_engine.log
tail -2 gpurun_out/s_pytest_multi.log; for f in gpurun_out/s_bench_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print(d.get('n_gpus'), d.get('value'), d.get('ms_per_step'), (d.get('e2e') or {}).get('value'))
PY
done; tail -12 gpurun_out/s_multi_engine.log
