#!/bin/bash
export PYTHONPATH=$PWD
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_motif.py tests/test_gpu_symbols.py tests/test_gpu_motif2.py -m gpu -q 2>&1 | tail -30 > gpurun_out/t_pytest.log; cat gpurun_out/t_pytest.log
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/t_pytest_all.log; cat gpurun_out/t_pytest_all.log
timeout 200 python tools/motif_phase_probe.py 2>&1 | tail -8
for fam in motif motif-data; do timeout 300 python bench.py --family $fam --steps 5 --warmup 3 --no-cpu --no-families > gpurun_out/t_bench_$fam.json 2> gpurun_out/t_bench_$fam.err; python - gpurun_out/t_bench_$fam.json <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print(d['config']['family'], d['value'], d['ms_per_step'], d['roofline']['frac'], d['gpu_launches'])
PY
done
