"""Writes tests/golden/regression_v1.json: results of the oracle (oracle/gibbs_oracle.c) on fixed seeded inputs.

These are REGRESSION vectors, not reference outputs: the reference cannot run here (no .NET) and has no tests of its
own. They freeze what the oracle -- pinned by SURVEY Appendix B, the Random123 Philox vectors
and the independent Python model -- returned when this file was generated, so that a later change to the uniform
stream, the draw order or a quirk shows up even if oracle and GPU code were changed together.
usage: python tests/golden/make_regression_vectors.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle_lib as O  # noqa: E402
from gibbssampling_b200.synthetic import planted_motif_set  # noqa: E402

CASES = [  # (name, n, L, min_len, k, alphabet_size, pc, family, seed, chain)
    ("bpv_c1_shape", 20, 100, None, 8, 5, 1e-4, "bpv", 0xB200, 0),
    ("bpv_ragged_k7", 12, 60, 33, 7, 5, 1e-4, "bpv", 11, 5),
    ("bpv_k20_alphabet4", 6, 150, 90, 20, 4, 1e-2, "bpv", 12, 1),
    ("data_script_call", 4, 21, None, 6, 5, 1e-4, "data", 2024, 3),
    ("data_k12", 9, 130, 70, 12, 5, 1e-4, "data", 13, 2),
    ("motif_fixed", 7, 80, 50, 8, 5, 1e-4, "motif", 14, 9),
    ("motif_data", 6, 60, None, 6, 5, 1e-4, "motif-data", 15, 4),
]
BG = [0.3, 0.2, 0.2, 0.3]


def run(case):
    name, n, L, lmin, k, alen, pc, family, seed, chain = case
    ps = planted_motif_set(n, L, k, seed=1000 + seed, min_length=lmin)
    seqs = ps.sequences()
    S = O.sources(seqs)
    alphabet = b"ATGC" if alen == 4 else b"ATGC-"
    rng, keep = O.make_rng(seed=seed, chain=chain)
    if family == "bpv":
        score, pos, st = O.site_step("do_site_sampling_with_bpv", S, k, pc, pcv=O.pcv_from_acgt(BG), rng=rng, alphabet=alphabet)
        sites = pos.tolist()
    elif family == "data":
        score, pos, st = O.site_step("do_site_sampling", S, k, pc, rng=rng, alphabet=alphabet)
        sites = pos.tolist()
    else:
        want, st = O.motif_step("do_motif_sampling", 1 if family == "motif-data" else 0, S, 1, k, pc, 0.0,
                                pcv=None if family == "motif-data" else O.pcv_from_acgt(BG), rng=rng, alphabet=alphabet)
        score = np.array([s for s, _ in want])
        sites = [p[0] if p else -1 for _, p in want]
    return {"name": name, "n": n, "L": L, "min_len": lmin, "k": k, "alphabet_size": alen, "pc": pc, "family": family,
            "seed": seed, "chain": chain, "bg": BG, "set_seed": 1000 + seed, "sequences": [s.decode() for s in seqs],
            "sites": sites, "scores": [float(x).hex() for x in score]}


if __name__ == "__main__":
    out = {"note": __doc__.strip().split("\n\n")[0], "cases": [run(c) for c in CASES]}
    with open(os.path.join(HERE, "regression_v1.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", len(out["cases"]), "cases")
