#!/bin/bash
# A/B of the barrier-free shift sweeps (GIBBS_INDEP_SWEEPS): whole GPU suite, then C2 bench lines of both builds
export PYTHONPATH=$PWD
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/u_pytest_all.log; cat gpurun_out/u_pytest_all.log
show() { python - "$1" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print(sys.argv[1], d['value'], d['ms_per_step'], d['roofline']['frac'], d['gpu_launches'], d.get('e2e',{}).get('value'))
PY
}
for rep in 1 2; do
for v in base noindep; do
lib=$PWD/variants/lib_$v.so; [ $v = base ] && lib=$PWD/gibbssampling_b200/libgibbs_b200.so
GIBBS_B200_LIB=$lib timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-families > gpurun_out/u_bench_${v}_$rep.json 2> gpurun_out/u_bench_${v}_$rep.err
show gpurun_out/u_bench_${v}_$rep.json
done
done
