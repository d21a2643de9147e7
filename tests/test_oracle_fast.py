"""The oracle's incremental mode must be bit-identical to its faithful mode.

The faithful functions restate GibbsSampling.fs operation by operation (from-scratch leave-one-out rebuilds, a PWM
per window, Array.skip copies) and are too slow for the BASELINE sizes; the incremental mode computes the same
float64 operations on the same integer counts (oracle/gibbs_oracle.c, "INCREMENTAL MODE"). Full-size golden
fixtures and the full-size GPU parity tests rely on it, so it is proven here, on CPU, against the faithful mode:
every pipeline, every phase function, ragged lengths, symbols outside the alphabet, pc = 0, injected and Philox
streams, 1 and several threads."""
import numpy as np
import pytest

import oracle_lib as O

ALPHABETS = [b"ATGC-", b"ACGT", b"ATGC"]


def _random_set(rng, n, lmin, lmax, symbols="ACGT", extra="", p_extra=0.0):
    seqs = []
    for _ in range(n):
        L = int(rng.integers(lmin, lmax + 1))
        s = rng.choice(list(symbols), size=L)
        if extra and p_extra > 0:
            m = rng.random(L) < p_extra
            s[m] = rng.choice(list(extra), size=int(m.sum()))
        seqs.append("".join(s))
    return seqs


def _pcv(rng):
    q = rng.random(4) + 0.2
    return O.pcv_from_acgt(q / q.sum())


def _same(a, b):
    assert a[1].tolist() == b[1].tolist()
    assert a[0].tobytes() == b[0].tobytes()


@pytest.mark.parametrize("seed", range(12))
def test_whole_pipelines_identical(seed):
    rng = np.random.default_rng(1000 + seed)
    n = int(rng.integers(2, 14))
    k = int(rng.integers(1, 12))
    seqs = _random_set(rng, n, k, k + int(rng.integers(0, 60)))
    alpha = ALPHABETS[seed % len(ALPHABETS)]
    pc = [1e-4, 1.0, 0.5, 1e-9][seed % 4]
    S = O.sources(seqs)
    pcv = _pcv(rng)
    for variant, name in ((0, "do_site_sampling_with_bpv"), (1, "do_site_sampling")):
        r1, _ = O.make_rng(seed=seed, chain=7)
        ref = O.site_step(name, S, k, pc, pcv=pcv if variant == 0 else None, rng=r1, alphabet=alpha)
        for threads in (1, 3):
            r2, _ = O.make_rng(seed=seed, chain=7)
            got = O.fast_site_pipeline(variant, S, k, pc, pcv=pcv if variant == 0 else None, rng=r2, threads=threads,
                                       alphabet=alpha)
            _same(ref, got)
            assert r2.next == r1.next
            assert (got[2].site_updates, got[2].window_scores, got[2].sweeps) == \
                   (ref[2].site_updates, ref[2].window_scores, ref[2].sweeps)


@pytest.mark.parametrize("seed", range(8))
def test_phase_functions_identical_from_given_state(seed):
    rng = np.random.default_rng(2000 + seed)
    n = int(rng.integers(2, 10))
    k = int(rng.integers(2, 9))
    seqs = _random_set(rng, n, k + 1, k + 40)
    S = O.sources(seqs)
    pcv = _pcv(rng)
    pos0 = np.array([rng.integers(0, len(s) - k + 1) for s in seqs], np.int32)
    sc0 = rng.normal(size=n) * 3
    table = [
        (0, O.PH_GREEDY, "find_best_motif_with_start_position"), (0, O.PH_LEFT, "left_shifted_with_bpv"),
        (0, O.PH_RIGHT, "right_shifted_with_bpv"), (1, O.PH_GREEDY, "best_pwms_with_start_positions"),
        (1, O.PH_LEFT, "left_shifted"), (1, O.PH_RIGHT, "right_shifted"),
    ]
    for variant, mask, name in table:
        ref = O.site_step(name, S, k, 1e-4, pcv=pcv if variant == 0 else None, state=(sc0, pos0))
        got = O.fast_site_pipeline(variant, S, k, 1e-4, pcv=pcv if variant == 0 else None, state=(sc0, pos0), phase_mask=mask)
        _same(ref, got)
        assert got[2].sweeps == ref[2].sweeps
    for variant, name in ((0, "random_starts_with_bpv"), (1, "random_starts")):
        u = rng.random(n * (n - 1))
        r1, k1 = O.make_rng(uniforms=u)
        r2, k2 = O.make_rng(uniforms=u)
        ref = O.site_step(name, S, k, 1e-4, pcv=pcv if variant == 0 else None, rng=r1)
        got = O.fast_site_pipeline(variant, S, k, 1e-4, pcv=pcv if variant == 0 else None, rng=r2, phase_mask=O.PH_INIT, threads=2)
        _same(ref, got)
        assert r1.next == r2.next == n * (n - 1)


@pytest.mark.parametrize("seed", range(6))
def test_symbols_outside_the_alphabet_and_zero_pseudocount(seed):
    rng = np.random.default_rng(3000 + seed)
    n, k = int(rng.integers(3, 9)), int(rng.integers(2, 7))
    seqs = _random_set(rng, n, k + 2, k + 30, extra="N-*RY", p_extra=0.08)
    S = O.sources(seqs)
    pcv = _pcv(rng)
    alpha = [b"ATGC-", b"ACGT"][seed % 2]
    pc = [1e-4, 0.0][seed % 2] if seed < 4 else 1e-4
    for variant, name in ((0, "do_site_sampling_with_bpv"), (1, "do_site_sampling")):
        if variant == 1 and pc == 0.0:
            continue   # 0/0 backgrounds: NaN scores compare unequal to themselves in both modes alike; covered by bytes below
        r1, _ = O.make_rng(seed=seed, chain=3)
        r2, _ = O.make_rng(seed=seed, chain=3)
        ref = O.site_step(name, S, k, pc, pcv=pcv if variant == 0 else None, rng=r1, alphabet=alpha)
        got = O.fast_site_pipeline(variant, S, k, pc, pcv=pcv if variant == 0 else None, rng=r2, alphabet=alpha)
        _same(ref, got)


def test_batched_chains_equal_single_chains_and_sums_are_left_to_right():
    rng = np.random.default_rng(5)
    seqs = _random_set(rng, 9, 30, 45)
    S = O.sources(seqs)
    pcv = _pcv(rng)
    for variant in (0, 1):
        scores, pos, sums, st = O.fast_site_chains(variant, S, 6, 1e-4, seed=99, chain_base=40, n_chains=5,
                                                   pcv=pcv if variant == 0 else None, threads=3)
        for c in range(5):
            r, _ = O.make_rng(seed=99, chain=40 + c)
            ref = O.site_step("do_site_sampling_with_bpv" if variant == 0 else "do_site_sampling", S, 6, 1e-4,
                              pcv=pcv if variant == 0 else None, rng=r)
            assert pos[c].tolist() == ref[1].tolist() and scores[c].tobytes() == ref[0].tobytes()
            acc = 0.0
            for v in ref[0]:
                acc = acc + float(v)
            assert sums[c] == acc
        assert st.restarts == 5


def test_sweep_log_counts_movers():
    rng = np.random.default_rng(6)
    seqs = _random_set(rng, 12, 40, 40)
    S = O.sources(seqs)
    pcv = _pcv(rng)
    r, _ = O.make_rng(seed=1, chain=0)
    score, pos, st, log = O.fast_site_pipeline(0, S, 5, 1e-4, pcv=pcv, rng=r)
    assert len(log) == st.sweeps - 1                      # every sweep but the random starts
    assert [m for m, _, _ in log] == sorted(m for m, _, _ in log)   # greedy, then left, then right
    for mode in (0, 1, 2):
        last = [x for x in log if x[0] == mode][-1]
        assert last[1] == 0                               # a phase ends with a sweep that moves nothing
