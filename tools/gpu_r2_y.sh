#!/bin/bash
# tiled random starts: warps per CTA variants on a C4-shaped INIT phase, then an ncu capture of init_tiled_kernel<10>
export PYTHONPATH=$PWD
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "random_starts" 2>&1 | tail -4
for v in base tw20 tw24; do
lib=$PWD/variants/lib_$v.so; [ $v = base ] && lib=$PWD/gibbssampling_b200/libgibbs_b200.so
echo "== $v"
GIBBS_B200_LIB=$lib timeout 600 python - <<'PY'
from gibbssampling_b200.engine import GibbsEngine, make_params
from gibbssampling_b200.synthetic import planted_motif_set, background_of
from gibbssampling_b200 import _abi
for (n,L,k,chains,reps) in [(100000,200,20,8,2),(10000,1000,16,64,2)]:
    ps = planted_motif_set(n, L, k); bg = background_of(ps.ascii, 1e-4, 5)
    eng = GibbsEngine(ps.sequences())
    pi = make_params(k, 1e-4, 5, bg, phase_mask=_abi.PHASE_INIT)
    for path in (_abi.GIBBS_INIT_TILED, _abi.GIBBS_INIT_WIDE):
        eng.set_option(_abi.GIBBS_OPT_INIT_PATH, path)
        for rep in range(reps):
            r = eng.run(pi, chains, seed=1+rep, want_sites=False, want_scores=False, want_counts=False); st=r.stats
            print(n,L,k,chains,"path",st['init_path'],"kernel_ms",round(st['kernel_ms'],3),"draws/s %.3e"%(st['site_updates']*(n-1)/(st['kernel_ms']*1e-3)),flush=True)
    eng.close()
PY
done
timeout 900 ncu --set full --clock-control none -k regex:init_tiled -c 1 -o /tmp/prof_tiled -f python tools/prof_init.py 20000 200 20 2 > gpurun_out/ncu_tiled.log 2>&1
echo "rc=$?" >> gpurun_out/ncu_tiled.log; tail -3 gpurun_out/ncu_tiled.log
ncu -i /tmp/prof_tiled.ncu-rep --page raw --csv > gpurun_out/tiled_raw.csv 2>/dev/null
ncu -i /tmp/prof_tiled.ncu-rep --page source --csv --print-source sass > gpurun_out/tiled_sass.csv 2>/dev/null
ls -la gpurun_out/tiled_*
