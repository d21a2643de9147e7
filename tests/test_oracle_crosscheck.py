"""Two independent restatements of the reference (C oracle, pure-Python model) must agree bit for bit,
on random small inputs including non-ACGT symbols, ragged lengths and every pipeline variant."""
import math

import numpy as np
import pytest

import oracle_lib as O
import pymodel as M

ALPHA = "ATGC-"


def _random_seqs(rng, n, lo, hi, symbols="ACGT"):
    return ["".join(rng.choice(list(symbols), size=rng.integers(lo, hi + 1))) for _ in range(n)]


def _pcv_dict(pcv49):
    return {chr(i + 42): float(pcv49[i]) for i in range(49)}


def _same(a, b):
    return a == b or (math.isnan(a) and math.isnan(b))


@pytest.mark.parametrize("seed", range(6))
def test_site_sampler_fixed_background(seed):
    rng = np.random.default_rng(seed)
    symbols = "ACGT" if seed % 2 == 0 else "ACGT-N*"
    n, k = int(rng.integers(2, 7)), int(rng.integers(1, 7))
    seqs = _random_seqs(rng, n, k, k + 25, symbols)
    pc = float(rng.choice([1e-4, 0.5, 1.0]))
    S = O.sources(seqs)
    pcv = O.pcv_of_sources(S, pc)
    u = rng.random(n * (n - 1))
    r, _ = O.make_rng(uniforms=u)
    score, pos, _ = O.site_step("do_site_sampling_with_bpv", S, k, pc, pcv=pcv, rng=r)
    want = M.do_site_sampling(k, pc, ALPHA, seqs, _pcv_dict(pcv), iter(u))
    assert pos.tolist() == [p for _, p in want]
    assert all(_same(a, b) for a, b in zip(score.tolist(), [s for s, _ in want]))


@pytest.mark.parametrize("seed", range(9))
def test_site_sampler_data_background(seed):
    rng = np.random.default_rng(100 + seed)
    symbols = ("ACGT", "ACGT-", "ACGTN*")[seed % 3]     # Gap is in the alphabet; N and Ter are not (dead rows, fs:67-69, fs:953)
    n, k = int(rng.integers(2, 6)), int(rng.integers(2, 6))
    seqs = _random_seqs(rng, n, k, k + 18, symbols)
    pc = 1e-4
    S = O.sources(seqs)
    u = rng.random(n * (n - 1))
    r, _ = O.make_rng(uniforms=u)
    score, pos, _ = O.site_step("do_site_sampling", S, k, pc, rng=r)
    want = M.do_site_sampling(k, pc, ALPHA, seqs, None, iter(u))
    assert pos.tolist() == [p for _, p in want]
    assert all(_same(a, b) for a, b in zip(score.tolist(), [s for s, _ in want]))


@pytest.mark.parametrize("seed", range(6))
@pytest.mark.parametrize("variant", [0, 1])
def test_motif_sampler(seed, variant):
    rng = np.random.default_rng(200 + seed)
    n, k, m = int(rng.integers(2, 5)), int(rng.integers(2, 5)), int(rng.integers(1, 3))
    seqs = _random_seqs(rng, n, k + 2, k + 14)
    pc, cutoff = 1e-4, float(rng.choice([0.0, 1.0, 3.0]))
    S = O.sources(seqs)
    pcv = O.pcv_of_sources(S, pc) if variant == 0 else None
    u = rng.random(n * (n - 1) + n)
    r, _ = O.make_rng(uniforms=u)
    try:
        got, _ = O.motif_step("do_motif_sampling", variant, S, m, k, pc, cutoff, pcv=pcv, rng=r)
        err = None
    except O.OracleError as e:
        got, err = None, e.code
    try:
        want = M.do_motif_sampling(m, k, pc, cutoff, ALPHA, seqs, _pcv_dict(pcv) if pcv is not None else None, iter(u))
    except IndexError:
        want = None
    if want is None:
        assert err == O.ERR_ROULETTE
        return
    assert err is None
    assert [p for _, p in got] == [p for _, p in want]
    assert all(_same(a, b) for a, b in zip([s for s, _ in got], [s for s, _ in want]))


def test_restart_loop_models_agree(golden):
    seqs, k, pc = golden["sequences"], golden["k"], golden["pc"]
    S = O.sources(seqs)
    pcv = O.pcv_of_sources(S, pc)
    n = len(seqs)
    for reps in (0, 1, 2, 4):
        u = np.random.default_rng(reps).random((reps + 1) * n * (n - 1))
        r, _ = O.make_rng(uniforms=u)
        s, p, _ = O.best_information_content(0, reps, S, k, pc, r, pcv=pcv)
        it = iter(u)
        want = M.restart_loop(reps, lambda: M.do_site_sampling(k, pc, ALPHA, seqs, _pcv_dict(pcv), it))
        assert p.tolist() == [q for _, q in want] and s.tolist() == [v for v, _ in want]


def test_symbol_and_length_errors():
    with pytest.raises(O.OracleError) as e:
        O.pcv_of_sources(O.sources(["ACGTa"]), 1e-4)
    assert e.value.code == O.ERR_SYMBOL
    S = O.sources(["ACGTACGT", "ACG"])
    r, _ = O.make_rng(seed=1)
    with pytest.raises(O.OracleError) as e:
        O.site_step("do_site_sampling_with_bpv", S, 4, 1e-4, pcv=O.pcv_from_acgt([.25] * 4), rng=r)
    assert e.value.code == O.ERR_SHORT_SEQ
    with pytest.raises(O.OracleError) as e:
        O.roulette([1.0, 1.0], 1.5)
    assert e.value.code == O.ERR_ROULETTE
