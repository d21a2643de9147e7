#!/bin/bash
# round 2, second session: evidence in one call (1 GPU) -- the GPU suite, smoke(), bench.py (default, reference arm, C1 / C3 /
# C4), the ncu launch lists of short C2 and C4 runs, ncu --set full of init_tiled_kernel<10> on a C4-shaped set
mkdir -p gpurun_out
export PYTHONPATH=$PWD
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,driver_version --format=csv > gpurun_out/f_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/f_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/f_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/f_bench_c2.json 2> gpurun_out/f_bench_c2.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/f_bench_reference_arm.json 2> gpurun_out/f_bench_ref.err
timeout 600 python bench.py --config C1 --steps 5 --warmup 3 --no-families > gpurun_out/f_bench_C1.json 2> gpurun_out/f_bench_C1.err
timeout 900 python bench.py --config C3 --steps 3 --warmup 2 --no-cpu --no-families > gpurun_out/f_bench_C3.json 2> gpurun_out/f_bench_C3.err
timeout 900 python bench.py --config C4 --steps 2 --warmup 1 --no-cpu --no-families > gpurun_out/f_bench_C4.json 2> gpurun_out/f_bench_C4.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/f_launches_c2.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-families > gpurun_out/f_ncu_launches.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/f_launches_C4.csv python bench.py --config C4 --steps 1 --warmup 1 --no-cpu --no-families > gpurun_out/f_ncu_launches_C4.log 2>&1
timeout 1500 ncu --set full --clock-control none -k regex:"chain|init_smem" -c 7 -o /tmp/prof_step -f python tools/prof_probe.py 1024 0 > gpurun_out/f_ncu_step.log 2>&1
ncu -i /tmp/prof_step.ncu-rep --page raw --csv > gpurun_out/f_step_raw.csv 2>/dev/null
timeout 900 ncu --set full --clock-control none -k regex:init_tiled -c 1 -o /tmp/prof_tiled -f python tools/prof_init.py 20000 200 20 2 > gpurun_out/f_ncu_tiled.log 2>&1
ncu -i /tmp/prof_tiled.ncu-rep --page raw --csv > gpurun_out/f_tiled_raw.csv 2>/dev/null
ncu -i /tmp/prof_tiled.ncu-rep --page source --csv --print-source sass > gpurun_out/f_tiled_sass.csv 2>/dev/null
tail -n 3 gpurun_out/f_pytest.log; cat gpurun_out/f_smoke.log | tail -n 2
for f in c2 C1 C3 C4; do python -c "
import json; d=json.load(open('gpurun_out/f_bench_$f.json')); print('$f', round(d['ms_per_step'],3), 'frac', round(d['roofline']['frac'],4), 'value %.4g' % d['value'], 'e2e ms', round(d['e2e']['ms_per_step'],3), 'cpu', (d.get('cpu_baseline') or {}).get('value'), d.get('families'))"; done
python -c "
import json; d=json.load(open('gpurun_out/f_bench_reference_arm.json')); print('ref', d['value'], d['ms_per_step'], d['cpu_baseline']['cores'])"
