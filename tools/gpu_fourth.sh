#!/bin/bash
mkdir -p gpurun_out
timeout 600 python - > gpurun_out/probe.log 2>&1 <<'PY'
import numpy as np
from gibbssampling_b200.engine import GibbsEngine, make_params
from gibbssampling_b200.synthetic import planted_motif_set, background_of
from gibbssampling_b200 import _abi
n,L,k=1000,500,12
ps = planted_motif_set(n, L, k); bg = background_of(ps.ascii, 1e-4, 5)
eng = GibbsEngine(ps.sequences())
p = make_params(k, 1e-4, 5, bg)
for team in (1,4):
    eng.set_team_warps(team)
    for cid in (0,1,2,3):
        r = eng.run(p, 1, seed=1, chain_id_base=cid, want_counts=False); st=r.stats
        print("single chain",cid,"team",team,"ms",round(st['kernel_ms'],3),"sweeps",st['sweeps'],"us/update",round(1e3*st['kernel_ms']/st['site_updates'],3),flush=True)
    # phases separately for chain 0
    full = eng.run(p, 1, seed=1, chain_id_base=0, want_counts=False)
    r0 = eng.run(make_params(k,1e-4,5,bg,phase_mask=_abi.PHASE_INIT), 1, seed=1, chain_id_base=0, want_counts=False)
    print(" init only: ms",round(r0.stats['kernel_ms'],3),"updates",r0.stats['site_updates'])
    eng.set_start_state(r0.sites, r0.scores)
    r1 = eng.run(make_params(k,1e-4,5,bg,phase_mask=_abi.PHASE_GREEDY), 1, seed=1, chain_id_base=0, want_counts=False)
    print(" greedy only: ms",round(r1.stats['kernel_ms'],3),"updates",r1.stats['site_updates'],"us/update",round(1e3*r1.stats['kernel_ms']/r1.stats['site_updates'],3))
    eng.set_start_state(r1.sites, r1.scores)
    r2 = eng.run(make_params(k,1e-4,5,bg,phase_mask=_abi.PHASE_LEFT), 1, seed=1, chain_id_base=0, want_counts=False)
    print(" left only: ms",round(r2.stats['kernel_ms'],3),"updates",r2.stats['site_updates'],"us/update",round(1e3*r2.stats['kernel_ms']/r2.stats['site_updates'],3))
# sweeps-per-chain distribution
eng.set_team_warps(0)
sw=[]
for cid in range(0,64):
    r = eng.run(p, 1, seed=1, chain_id_base=cid, want_counts=False, want_sites=False, want_scores=False); sw.append(r.stats['sweeps'])
print("sweeps per chain: mean",np.mean(sw),"max",max(sw),"min",min(sw),sorted(sw))
PY
