#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/i_launches.csv python bench.py --steps 6 --warmup 1 --no-cpu --no-families > gpurun_out/i_ncu.log 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(l for l in open('gpurun_out/i_launches.csv') if l.startswith('"')))
for r in rows[1:]:
    if 'gibbs' in r[4]: print(r[4][:60], r[7], r[8], r[-1])
PY
