#!/bin/bash
# init_tiled_kernel with the Philox round keys as a kernel parameter (variants/lib_keyed.so) against the in-tree build
export PYTHONPATH=$PWD
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "random_starts" 2>&1 | tail -3
for v in base; do
lib=$PWD/variants/lib_$v.so; [ $v = base ] && lib=$PWD/gibbssampling_b200/libgibbs_b200.so
echo "== $v"
GIBBS_B200_LIB=$lib timeout 600 python - <<'PY'
from gibbssampling_b200.engine import GibbsEngine, make_params
from gibbssampling_b200.synthetic import planted_motif_set, background_of
from gibbssampling_b200 import _abi
for (n,L,k,chains,reps) in [(100000,200,20,8,3),(10000,1000,16,64,2)]:
    ps = planted_motif_set(n, L, k); bg = background_of(ps.ascii, 1e-4, 5)
    eng = GibbsEngine(ps.sequences())
    pi = make_params(k, 1e-4, 5, bg, phase_mask=_abi.PHASE_INIT)
    eng.set_option(_abi.GIBBS_OPT_INIT_PATH, _abi.GIBBS_INIT_TILED)
    for rep in range(reps):
        r = eng.run(pi, chains, seed=1, want_scores=False, want_counts=False); st=r.stats
        print(n,L,k,chains,"path",st['init_path'],"kernel_ms",round(st['kernel_ms'],3),"draws/s %.3e"%(st['site_updates']*(n-1)/(st['kernel_ms']*1e-3)), int(r.sites.sum()),flush=True)
    eng.close()
PY
done
