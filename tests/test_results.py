"""Host-side result post-processing, checked against the oracle's PPM / PWM."""
import math

import numpy as np
import pytest

import oracle_lib as O
from gibbssampling_b200 import Results
from gibbssampling_b200.MotifSampler import createMotifIndex
from gibbssampling_b200.synthetic import background_of, planted_motif_set


def test_count_by_positions_keeps_first_occurrence_order_and_sorts_stably():
    a = [(1.0, 3), (2.0, 5)]
    b = [(0.5, 1), (0.1, 1)]
    c = [(9.0, 7), (9.0, 7)]
    got = Results.countByPositions([a, b, c, b, a, b])
    assert got == [((1, 1), 3), ((3, 5), 2), ((7, 7), 1)]
    # equal counts: first-seen group first (Array.countBy order + stable sortByDescending)
    assert Results.countByPositions([a, b]) == [((3, 5), 1), ((1, 1), 1)]


def test_count_by_positions_of_motif_indices():
    x = [createMotifIndex(1.0, [4]), createMotifIndex(0.2, [])]
    y = [createMotifIndex(3.0, [4]), createMotifIndex(0.1, [])]
    assert Results.countByPositions([x, y]) == [(((4,), ()), 2)]
    assert Results.scoreTable(x) == [(0, 4, 1.0)]


def test_defined_segment_matches_skip_take():
    s = b"ACGTACGTAA"
    assert Results.getDefinedSegment(4, s, 2) == b"GTAC"
    assert Results.getDefinedSegment(0, s, 10) == b""
    with pytest.raises(ValueError):
        Results.getDefinedSegment(4, s, 7)
    with pytest.raises(ValueError):
        Results.getDefinedSegment(1, s, 11)
    assert Results.segmentsOf([b"AAACCC", b"GGGTTT"], [1, 2], 3) == [b"AAC", b"GTT"]


def test_profile_matches_oracle_ppm_and_pwm():
    n, L, k, pc = 9, 40, 7, 1e-4
    ps = planted_motif_set(n, L, k, seed=4)
    seqs = ps.sequences()
    bg = background_of(ps.ascii, pc, 5)
    S = O.sources(seqs)
    sites = np.array(ps.truth, dtype=np.int32)
    pfm = O.loo_pfm(S, sites, 2, k)
    counts = O.acgt_counts(pfm)
    prof = Results.motifProfile(counts, n - 1, pc, 5, bg)
    ppm = O.ppm_of_pfm(pfm, n - 1, pc, b"ATGC-")
    rows = [ord(c) - 42 for c in "ACGT"]
    want = np.array([[ppm[r, j] for r in rows] for j in range(k)])
    assert np.array_equal(prof.ppm, want)                      # same IEEE operations
    assert np.array_equal(prof.pwm, want / np.asarray(bg))
    # the window product of the consensus against this PWM is the largest any k-mer can reach
    best = float(np.prod(prof.pwm.max(axis=1)))
    idx = ["ACGT".index(c) for c in prof.consensus]
    assert math.isclose(float(np.prod([prof.pwm[j, b] for j, b in enumerate(idx)])), best, rel_tol=1e-12)
    assert prof.information.shape == (k,) and prof.total_information > 0
    # a planted motif with 10 % mutations: consensus of the sites agrees with the planted consensus mostly
    full = Results.motifProfile(O.acgt_counts(O.loo_pfm(S, sites, 0, k)), n - 1, pc, 5, bg)
    agree = sum(a == b for a, b in zip(full.consensus, ps.consensus.decode()))
    assert agree >= k - 2


def test_profile_rejects_bad_shapes():
    with pytest.raises(ValueError):
        Results.motifProfile(np.zeros((3, 5)), 4, 1e-4, 5, [0.25] * 4)
    with pytest.raises(ValueError):
        Results.motifProfile(np.zeros((3, 4)), 4, 1e-4, 5, [0.25] * 3)


def test_count_by_positions_matches_a_naive_count():
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=100, deadline=None)
    @given(st.lists(st.lists(st.integers(0, 3), min_size=2, max_size=2), min_size=0, max_size=30))
    def check(vectors):
        results = [[(0.5, p) for p in v] for v in vectors]
        got = Results.countByPositions(results)
        keys = [tuple(v) for v in vectors]
        assert sum(c for _, c in got) == len(vectors)
        assert {k: c for k, c in got} == {k: keys.count(k) for k in set(keys)}
        counts = [c for _, c in got]
        assert counts == sorted(counts, reverse=True)
        for (k1, c1), (k2, c2) in zip(got, got[1:]):                       # stable: equal counts keep first-seen order
            if c1 == c2:
                assert keys.index(k1) < keys.index(k2)

    check()
