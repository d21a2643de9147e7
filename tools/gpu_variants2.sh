#!/bin/bash
# compare builds in variants/ (plus the in-tree build as "base") on full C2 runs
mkdir -p gpurun_out
for v in base "$@"; do
lib=$PWD/variants/lib_$v.so; [ $v = base ] && lib=$PWD/gibbssampling_b200/libgibbs_b200.so
GIBBS_B200_LIB=$lib timeout 600 python - > gpurun_out/var2_$v.log 2>&1 <<'PY'
from gibbssampling_b200.engine import GibbsEngine, make_params
from gibbssampling_b200.synthetic import planted_motif_set, background_of
for (n,L,k,chains,reps) in [(1000,500,12,1024,6),(10000,1000,16,512,1)]:
    ps = planted_motif_set(n, L, k); bg = background_of(ps.ascii, 1e-4, 5)
    eng = GibbsEngine(ps.sequences()); p = make_params(k, 1e-4, 5, bg)
    for rep in range(reps):
        r = eng.run(p, chains, seed=1+rep, want_sites=False, want_scores=False, want_counts=False); st=r.stats
        print(n,L,k,chains,"full kernel_ms",round(st['kernel_ms'],3),"win/s %.3e"%(st['window_scores']/(st['kernel_ms']*1e-3)),flush=True)
    eng.close()
PY
echo "== $v"; cat gpurun_out/var2_$v.log
done
