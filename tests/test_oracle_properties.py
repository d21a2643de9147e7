"""Property tests of the oracle (hypothesis): the invariants the GPU design relies on (SURVEY.md section 4 iii)."""
import math

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

import oracle_lib as O

BASES = "ACGT"
dna = st.text(alphabet=BASES, min_size=8, max_size=40)


@settings(max_examples=60, deadline=None)
@given(st.lists(dna, min_size=2, max_size=7), st.integers(1, 8), st.data())
def test_incremental_counts_equal_from_scratch_fuse(seqs, k, data):
    """counts(held-out h) = counts(all sites) - one-hot(site of h): the identity behind the in-place count update."""
    k = min(k, min(len(s) for s in seqs))
    sites = [data.draw(st.integers(0, len(s) - k)) for s in seqs]
    S = O.sources(seqs)
    total = np.zeros((k, 4), dtype=np.int64)
    for s, p in zip(seqs, sites):
        for j in range(k):
            total[j, BASES.index(s[p + j])] += 1
    for h in range(len(seqs)):
        loo = O.acgt_counts(O.loo_pfm(S, sites, h, k))
        own = np.zeros((k, 4), dtype=np.int64)
        for j in range(k):
            own[j, BASES.index(seqs[h][sites[h] + j])] += 1
        assert (total - own).tolist() == loo.tolist()
        # moving one site = -old k-mer + new k-mer
        new_p = data.draw(st.integers(0, len(seqs[h]) - k))
        moved = list(sites)
        moved[h] = new_p
        g = (h + 1) % len(seqs)
        neu = np.zeros((k, 4), dtype=np.int64)
        for j in range(k):
            neu[j, BASES.index(seqs[h][new_p + j])] += 1
        own_g = np.zeros((k, 4), dtype=np.int64)
        for j in range(k):
            own_g[j, BASES.index(seqs[g][sites[g] + j])] += 1
        assert (total - own + neu - own_g).tolist() == O.acgt_counts(O.loo_pfm(S, moved, g, k)).tolist()


@settings(max_examples=60, deadline=None)
@given(dna, st.integers(1, 6), st.floats(1e-4, 1.0), st.data())
def test_argmax_is_first_strict_maximum(seq, k, pc, data):
    k = min(k, len(seq))
    others = [data.draw(dna) for _ in range(3)]
    others = [o for o in others if len(o) >= k] or [seq]
    seqs = [seq] + others
    S = O.sources(seqs)
    sites = [0] + [data.draw(st.integers(0, len(o) - k)) for o in others]
    pcv = O.pcv_of_sources(S, pc)
    ppm = O.ppm_of_pfm(O.loo_pfm(S, sites, 0, k), len(seqs) - 1, pc)
    raw = O.window_scores_bpv(seq.encode(), k, pcv, ppm)
    score, pos = O.best_pwms_with_bpv(seq.encode(), k, pcv, ppm)
    m = raw.max()
    assert pos == int(np.argmax(raw)) == int(np.nonzero(raw == m)[0][0])
    assert score == math.log(m) / math.log(2.0)   # libm, like the oracle (numpy's SIMD log can differ by an ulp)
    # identical k-mers tie exactly; the first one wins
    if k <= len(seq) // 2:
        rep = seq[:k] * 3
        raw2 = O.window_scores_bpv(rep.encode(), k, pcv, ppm)
        _, pos2 = O.best_pwms_with_bpv(rep.encode(), k, pcv, ppm)
        assert raw2[0] == raw2[k] == raw2[2 * k]
        assert pos2 == int(np.nonzero(raw2 == raw2.max())[0][0]) and pos2 < k


@settings(max_examples=100, deadline=None)
@given(st.lists(st.floats(1e-6, 50.0), min_size=1, max_size=30), st.floats(0.0, 0.999999))
def test_roulette_walk_matches_definition(weights, pick):
    """fs:746-754: first n with acc <= pick <= acc + w_n, sequential float64 accumulation."""
    total = 0.0
    for w in weights:
        total = total + w
    acc, want = 0.0, None
    for i, w in enumerate(weights):
        wn = w / total
        if acc <= pick <= acc + wn:
            want = i
            break
        acc = acc + wn
    if want is None:
        with pytest.raises(O.OracleError):
            O.roulette(weights, pick)
    else:
        assert O.roulette(weights, pick) == want
    assert O.roulette(weights, 0.0) == 0            # acc = 0 <= 0 <= w_0


@settings(max_examples=40, deadline=None)
@given(dna, st.integers(1, 6), st.data())
def test_drifting_background_closed_form(seq, k, data):
    """quirk A.6-1: counts at window w = F0 + (w+1) * counts(seq) - sum_{m<=w} counts(window m), never clamped."""
    k = min(k, len(seq))
    f0 = {b: data.draw(st.integers(0, 30)) for b in BASES}
    fcv = np.zeros(O.NSLOT, dtype=np.int32)
    for b, v in f0.items():
        fcv[ord(b) - 42] = v
    ppm = np.full((O.NSLOT, k), 0.25)
    _, _, raw, out = O.best_pwms(seq.encode(), k, 0.5, fcv, ppm)
    W = len(seq) - k + 1
    for b in BASES:
        inside = sum(seq[m:m + k].count(b) for m in range(W))
        assert out[ord(b) - 42] == f0[b] + W * seq.count(b) - inside
    assert len(raw) == W and (raw > 0).all()


@settings(max_examples=30, deadline=None)
@given(st.integers(0, 2**63 - 1), st.integers(0, 2**40), st.integers(0, 2**40))
def test_uniform_stream_is_a_pure_function_of_its_index(seed, chain, draw):
    u = O.uniform_at(seed, chain, draw)
    assert 0.0 <= u < 1.0 and u * 4294967296.0 == int(u * 4294967296.0)
    assert u == O.uniform_at(seed, chain, draw)
    rng, _ = O.make_rng(seed=seed, chain=chain)
    rng.next = draw
    assert O.lib().or_next_uniform(__import__("ctypes").byref(rng)) == u
