#!/bin/bash
# one 8-GPU box, final build of round 2: the default bench at N = 2, 4, 8 (weak scaling, 1024 chains per GPU) and N = 1 on the same box
export PYTHONPATH=$PWD
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 --no-cpu --no-families > gpurun_out/s3_bench_n1.json 2> gpurun_out/s3_bench_n1.err
port=29720
for N in 2 4 8; do
  port=$((port+1))
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --steps 10 --warmup 3 --no-cpu --no-families > gpurun_out/s3_bench_n$N.json 2> gpurun_out/s3_bench_n$N.err; echo "rc=$?" >> gpurun_out/s3_bench_n$N.err
done
for f in gpurun_out/s3_bench_n*.json; do python - "$f" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print(d.get('n_gpus'), d.get('value'), d.get('ms_per_step'), (d.get('e2e') or {}).get('value'), [round(r[1],2) for r in d['per_rank']['ranks']])
PY
done
