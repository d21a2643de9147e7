"""BASELINE.json configs[4]: throughput sweep k in {6,12,20,30} x chains in {1..65536, powers of 4} x L in {100 bp, 1 kb, 10 kb}
on ONE GPU (the multi-GPU axis is weak scaling over chains: see profiles/r02_bench_n{2,4,8}.json).
SiteSampler WithBPV restarts with phase shifts (the benchmarked family). Prints one JSON line per point and a table.
usage: sweep_c5.py [--out profiles/r02_sweep_c5.json] [--max-chains 65536]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gibbssampling_b200.engine import GibbsEngine, make_params  # noqa: E402
from gibbssampling_b200.synthetic import background_of, planted_motif_set  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--out", default="gpurun_out/sweep_c5.json")
ap.add_argument("--max-chains", type=int, default=65536)
args = ap.parse_args()
lib_peak = None
rows = []
for L in (100, 1000, 10000):
    n = 1000 if L <= 1000 else 200          # sequences per set (C5 leaves N open): the C2 count, fewer for the 10 kb rows
    for k in (6, 12, 20, 30):
        ps = planted_motif_set(n, L, k, seed=0xC5 + k)
        bg = background_of(ps.ascii, 1e-4, 5)
        params = make_params(k, 1e-4, 5, bg)
        with GibbsEngine(ps.sequences()) as eng:
            if lib_peak is None:
                lib_peak = eng.measure_smem_bandwidth() if hasattr(eng, "measure_smem_bandwidth") else None
            chains = 1
            while chains <= args.max_chains:
                best = None
                for rep in range(2 if chains <= 4096 else 1):
                    r = eng.run(params, chains, seed=11 + rep, want_sites=False, want_scores=False, want_counts=False)
                    st = r.stats
                    if best is None or st["kernel_ms"] < best["kernel_ms"]:
                        best = st
                ws = best["window_scores"] / (best["kernel_ms"] * 1e-3)
                row = {"L": L, "n_seqs": n, "k": k, "chains": chains, "kernel_ms": round(best["kernel_ms"], 3),
                       "window_scores_per_s": ws, "site_updates_per_s": best["site_updates"] / (best["kernel_ms"] * 1e-3),
                       "sweeps_per_chain": best["sweeps"] / chains, "smem_bytes_per_window": 4.0 * k + k / 4.0,
                       "smem_GBps": ws * (4.0 * k + k / 4.0) / 1e9, "init_path": best["init_path"], "team_warps": best["team_warps"],
                       "fast_path": best["fast_path"], "exact_rescans": best["exact_rescans"]}
                rows.append(row)
                print(json.dumps(row), flush=True)
                chains *= 4
os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
with open(args.out, "w") as f:
    json.dump({"what": "BASELINE.json configs[4] on one B200: SiteSampler WithBPV restarts with phase shifts", "points": rows}, f, indent=1)
print("\nL      k   " + "".join(f"{c:>10d}" for c in sorted({r['chains'] for r in rows})))
for L in (100, 1000, 10000):
    for k in (6, 12, 20, 30):
        pts = {r["chains"]: r for r in rows if r["L"] == L and r["k"] == k}
        print(f"{L:<6d} {k:<3d} " + "".join(f"{pts[c]['window_scores_per_s']:>10.2e}" if c in pts else " " * 10 for c in sorted(pts)))
