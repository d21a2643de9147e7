#!/bin/bash
# one 8-GPU box, second session of round 2: C4 and C3 as BASELINE states them (8 x B200) and the default C2 line at N = 8
export PYTHONPATH=$PWD
mkdir -p gpurun_out
nvidia-smi -L | wc -l > gpurun_out/s2_ngpu.txt
port=29620
run() { port=$((port+1)); tag=$1; shift; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $port bench.py --gpus 8 "$@" > gpurun_out/s2_bench_n8_$tag.json 2> gpurun_out/s2_bench_n8_$tag.err; echo "rc=$?" >> gpurun_out/s2_bench_n8_$tag.err; }
run C4 --config C4 --steps 2 --warmup 1 --no-cpu --no-families
run C3 --config C3 --steps 2 --warmup 2 --no-cpu --no-families
run c2 --steps 10 --warmup 3 --no-cpu --no-families
for f in gpurun_out/s2_bench_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print(d.get('n_gpus'), d.get('value'), d.get('ms_per_step'), (d.get('e2e') or {}).get('value'), d['measurement']['init_path'])
PY
done
