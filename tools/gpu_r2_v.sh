#!/bin/bash
# coop rounds (GIBBS_OPT_COOP): parity tests, then C2 bench A/B (option on / off / the build before the change)
export PYTHONPATH=$PWD
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_coop.py tests/test_gpu_fullsize.py tests/test_gpu_parity.py tests/test_gpu_race.py tests/test_gpu_cluster.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/v_pytest.log; cat gpurun_out/v_pytest.log
show() { python - "$1" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print(sys.argv[1], d['value'], d['ms_per_step'], d['roofline']['frac'], d['gpu_launches'], d.get('e2e',{}).get('value'))
PY
}
for rep in 1 2; do
GIBBS_B200_LIB=$PWD/variants/lib_precoop.so timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-families > gpurun_out/v_bench_pre_$rep.json 2> gpurun_out/v_bench_pre_$rep.err; show gpurun_out/v_bench_pre_$rep.json
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-families --opt coop=0 > gpurun_out/v_bench_off_$rep.json 2> gpurun_out/v_bench_off_$rep.err; show gpurun_out/v_bench_off_$rep.json
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-families > gpurun_out/v_bench_on_$rep.json 2> gpurun_out/v_bench_on_$rep.err; show gpurun_out/v_bench_on_$rep.json
done
tail -3 gpurun_out/v_bench_on_1.err
