"""Frozen regression vectors (tests/golden/regression_v1.json, written by tests/golden/make_regression_vectors.py):
the oracle must still return them (CPU), and the GPU must return them without consulting the oracle (-m gpu)."""
import json
import os

import numpy as np
import pytest

import oracle_lib as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "regression_v1.json")
with open(GOLDEN) as f:
    CASES = json.load(f)["cases"]


def _scores(case):
    return np.array([float.fromhex(x) for x in case["scores"]])


@pytest.mark.parametrize("case", CASES, ids=lambda c: c["name"])
def test_oracle_still_returns_the_frozen_results(case):
    seqs = [s.encode() for s in case["sequences"]]
    S = O.sources(seqs)
    alphabet = b"ATGC" if case["alphabet_size"] == 4 else b"ATGC-"
    rng, keep = O.make_rng(seed=case["seed"], chain=case["chain"])
    fam, k, pc = case["family"], case["k"], case["pc"]
    if fam in ("bpv", "data"):
        score, pos, _ = O.site_step("do_site_sampling_with_bpv" if fam == "bpv" else "do_site_sampling", S, k, pc,
                                    pcv=O.pcv_from_acgt(case["bg"]) if fam == "bpv" else None, rng=rng, alphabet=alphabet)
        sites = pos.tolist()
    else:
        want, _ = O.motif_step("do_motif_sampling", 1 if fam == "motif-data" else 0, S, 1, k, pc, 0.0,
                               pcv=None if fam == "motif-data" else O.pcv_from_acgt(case["bg"]), rng=rng, alphabet=alphabet)
        score = np.array([s for s, _ in want])
        sites = [p[0] if p else -1 for _, p in want]
    assert sites == case["sites"]
    assert score.tobytes() == _scores(case).tobytes()          # the oracle is deterministic to the last bit


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=lambda c: c["name"])
def test_gpu_returns_the_frozen_results(case):
    from gibbssampling_b200 import _abi
    from gibbssampling_b200.engine import GibbsEngine, make_params
    seqs = [s.encode() for s in case["sequences"]]
    fam = case["family"]
    params = make_params(case["k"], case["pc"], case["alphabet_size"], case["bg"], cutoff=0.0,
                         sampler=_abi.GIBBS_MOTIF_SAMPLER if fam.startswith("motif") else _abi.GIBBS_SITE_SAMPLER,
                         background=_abi.GIBBS_BG_DATA if fam in ("data", "motif-data") else _abi.GIBBS_BG_FIXED)
    with GibbsEngine(seqs) as eng:
        res = eng.run(params, 1, chain_id_base=case["chain"], seed=case["seed"], want_counts=False)
    assert res.sites[0].tolist() == case["sites"]
    np.testing.assert_allclose(res.scores[0], _scores(case), rtol=1e-5)
