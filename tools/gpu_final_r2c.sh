#!/bin/bash
# round 2, end of the second session: evidence of the final defaults (first sweep on one warp, never-narrowing team stages,
# hand-over to 8-warp teams at 4 chains per SM) -- the GPU suite, smoke(), bench lines of C2 / C3 / reference arm, the launch
# list of a short C2 run, ncu --set full of the kernels of one C2 step (raw page)
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/k_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/k_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/k_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/k_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/k_bench_c2.json 2> gpurun_out/k_bench_c2.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/k_bench_reference_arm.json 2> gpurun_out/k_bench_ref.err
timeout 900 python bench.py --config C3 --steps 3 --warmup 2 --no-cpu --no-families > gpurun_out/k_bench_C3.json 2> gpurun_out/k_bench_C3.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/k_launches_c2.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-families > gpurun_out/k_ncu_launches.log 2>&1
timeout 1500 ncu --set full --clock-control none -k regex:"chain|init_smem" -c 7 -o /tmp/prof_step -f python tools/prof_probe.py 1024 0 > gpurun_out/k_ncu_step.log 2>&1
ncu -i /tmp/prof_step.ncu-rep --page raw --csv > gpurun_out/k_step_raw.csv 2>/dev/null
tail -n 3 gpurun_out/k_pytest.log; cat gpurun_out/k_smoke.log | tail -n 2
for f in c2 C3; do python -c "
import json; d=json.load(open('gpurun_out/k_bench_$f.json')); print('$f', round(d['ms_per_step'],3), 'frac', round(d['roofline']['frac'],4), 'value %.4g' % d['value'], 'e2e ms', round(d['e2e']['ms_per_step'],3), 'cpu', (d.get('cpu_baseline') or {}).get('value'), {k:(round(v['ms_per_step'],2), round(v['frac'],4)) for k,v in (d.get('families') or {}).items()})"; done
python -c "
import json; d=json.load(open('gpurun_out/k_bench_reference_arm.json')); print('ref', d['value'], d['ms_per_step'], d['cpu_baseline']['cores'])"
