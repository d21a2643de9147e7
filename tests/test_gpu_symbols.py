"""Symbols outside A,C,G,T: the reference's 49-slot tables accept any symbol in
'*'..'Z'; one that is not in `alphabet` has a PWM row of 0 (fs:283-287), so a window holding it scores 0, and a
site that covers it adds nothing to the A,C,G,T counts of that column (fs:211-215). GPU (mask plane) vs oracle."""
import numpy as np
import pytest

import oracle_lib as O
from gibbssampling_b200 import _abi
from gibbssampling_b200.BioArray import ofNucleotideString
from gibbssampling_b200.engine import GibbsEngine, draws_per_chain, make_params

pytestmark = pytest.mark.gpu

LOG2_RTOL = 1e-5
BG = [0.3, 0.2, 0.2, 0.3]
TEAMS = [1, 4, 8, 16]


def _seqs(seed, n, lo, hi, symbols, p_other):
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        L = int(rng.integers(lo, hi + 1))
        s = rng.choice(list("ACGT"), size=L)
        other = rng.random(L) < p_other
        s[other] = rng.choice(list(symbols), size=int(other.sum()))
        out.append("".join(s).encode())
    return out


def _alphabet(alen):
    return b"ATGC" if alen == 4 else b"ATGC-"


CASES = [  # (seed, n, lo, hi, k, other symbols, fraction, alphabet_size)
    (1, 6, 30, 60, 6, "NRY*", 0.05, 5),
    (2, 9, 40, 90, 12, "N", 0.02, 5),
    (3, 5, 25, 40, 7, "N-W", 0.10, 4),       # Gap is just another non-alphabet symbol when |alphabet| = 4
    (4, 12, 300, 400, 16, "NB", 0.004, 5),   # long rows: most windows valid, 16-window chunks next to masked rows
    (5, 4, 12, 20, 3, "N", 0.30, 5),         # heavy masking: whole rows can score 0 -> (-inf, 0)
    (6, 7, 70, 70, 20, "KM", 0.01, 5),
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"s{c[0]}_n{c[1]}_k{c[4]}")
def test_primitives_with_masked_symbols(case):
    seed, n, lo, hi, k, sym, frac, alen = case
    seqs = _seqs(seed, n, lo, hi, sym, frac)
    S = O.sources(seqs)
    pcv = O.pcv_from_acgt(BG)
    alphabet = _alphabet(alen)
    rng = np.random.default_rng(seed)
    p = make_params(k, 1e-4, alen, BG)
    with GibbsEngine(seqs) as eng:
        for h in range(n):
            sites = np.array([rng.integers(0, len(s) - k + 1) for s in seqs], dtype=np.int32)
            pfm = O.loo_pfm(S, sites, h, k)
            assert eng.loo_counts(sites, h, k).tolist() == O.acgt_counts(pfm).tolist()
            ppm = O.ppm_of_pfm(pfm, n - 1, 1e-4, alphabet)
            want_raw = O.window_scores_bpv(seqs[h], k, pcv, ppm, alphabet)
            raw, lg = eng.window_scores(sites, h, p)
            assert raw.tolist() == want_raw.tolist()
            score, pos = O.best_pwms_with_bpv(seqs[h], k, pcv, ppm, alphabet)
            got_score, got_pos = eng.pick_argmax(sites, h, p)
            assert got_pos == pos
            if np.isinf(score):
                assert got_score == score
            else:
                assert got_score == pytest.approx(score, rel=LOG2_RTOL)


@pytest.mark.parametrize("wide", ["0", "1"])
@pytest.mark.parametrize("team", TEAMS)
@pytest.mark.parametrize("case", CASES, ids=lambda c: f"s{c[0]}_n{c[1]}_k{c[4]}")
def test_chains_with_masked_symbols(case, team, wide):
    seed, n, lo, hi, k, sym, frac, alen = case
    seqs = _seqs(seed, n, lo, hi, sym, frac)
    S = O.sources(seqs)
    pcv = O.pcv_from_acgt(BG)
    alphabet = _alphabet(alen)
    n_chains = 5
    with GibbsEngine(seqs) as eng:
        eng.set_team_warps(team)
        eng.set_option(_abi.GIBBS_OPT_INIT_PATH, _abi.GIBBS_INIT_WIDE if wide == "1" else _abi.GIBBS_INIT_CHAIN)
        res = eng.run(make_params(k, 1e-4, alen, BG), n_chains, chain_id_base=40, seed=77 + seed)
    for c in range(n_chains):
        rng, keep = O.make_rng(seed=77 + seed, chain=40 + c)
        score, pos, st = O.site_step("do_site_sampling_with_bpv", S, k, 1e-4, pcv=pcv, rng=rng, alphabet=alphabet)
        assert res.sites[c].tolist() == pos.tolist(), f"chain {c}"
        finite = np.isfinite(score)
        assert np.array_equal(np.isfinite(res.scores[c]), finite)
        np.testing.assert_allclose(res.scores[c][finite], score[finite], rtol=LOG2_RTOL)
        assert np.array_equal(res.scores[c][~finite], score[~finite])
    best = res.best_chain
    want = np.zeros((k, 4), dtype=np.int64)
    for i, s in enumerate(seqs):
        for j in range(k):
            ch = chr(s[res.sites[best][i] + j])
            if ch in "ACGT":
                want[j, "ACGT".index(ch)] += 1
    assert res.counts.tolist() == want.tolist()


def test_injected_uniforms_and_phase_functions_with_masked_symbols():
    seqs = _seqs(9, 8, 30, 50, "NS", 0.06)
    n, k = len(seqs), 5
    S = O.sources(seqs)
    pcv = O.pcv_from_acgt(BG)
    u = np.random.default_rng(3).random((2, draws_per_chain(n)))
    with GibbsEngine(seqs) as eng:
        res = eng.run(make_params(k, 1e-4, 5, BG), 2, uniforms=u)
        for c in range(2):
            rng, keep = O.make_rng(uniforms=u[c])
            score, pos, _ = O.site_step("do_site_sampling_with_bpv", S, k, 1e-4, pcv=pcv, rng=rng)
            assert res.sites[c].tolist() == pos.tolist()
        # sweeps from a start state whose sites cover masked bases
        start = np.array([max(0, s.find(b"N") - 1) if b"N" in s else 0 for s in seqs], dtype=np.int32)
        start = np.minimum(start, [len(s) - k for s in seqs]).astype(np.int32)
        scores0 = np.full(n, -50.0)
        for mask, name in ((_abi.PHASE_GREEDY, "find_best_motif_with_start_position"),
                           (_abi.PHASE_LEFT, "left_shifted_with_bpv"), (_abi.PHASE_RIGHT, "right_shifted_with_bpv")):
            eng.set_start_state(start, scores0)
            got = eng.run(make_params(k, 1e-4, 5, BG, phase_mask=mask), 1)
            score, pos, _ = O.site_step(name, S, k, 1e-4, pcv=pcv, state=(scores0, start))
            assert got.sites[0].tolist() == pos.tolist(), name
            np.testing.assert_allclose(got.scores[0], score, rtol=LOG2_RTOL)


def test_symbol_errors_and_unsupported_combinations():
    with pytest.raises(_abi.GibbsSymbolError):
        GibbsEngine([b"ACGTacgt", b"ACGTACGT"])                 # lower case is outside '*'..'Z' (the parser upper-cases)
    with pytest.raises(_abi.GibbsSymbolError):
        GibbsEngine([b"ACGT ACGT", b"ACGTACGT"])
    seqs = [ofNucleotideString("acgt-acgtnacg\n tacgt"), ofNucleotideString("ACGTACGTAC GTAC x GT")]
    assert seqs == [b"ACGT-ACGTNACGTACGT", b"ACGTACGTACGTACGT"]
    with GibbsEngine(seqs) as eng:
        with pytest.raises(_abi.GibbsUnsupportedError):           # Gap present and in the alphabet: a fifth PWM row
            eng.run(make_params(4, 1e-4, 5, BG), 1)
        ok = eng.run(make_params(4, 1e-4, 4, BG), 2, seed=5)      # alphabet of 4: Gap scores 0 like N
        assert ok.sites.shape == (2, 2)
        data = eng.run(make_params(4, 1e-4, 4, BG, background=_abi.GIBBS_BG_DATA), 2, seed=5)   # built (see below)
        assert data.sites.shape == (2, 2)
        motif = eng.run(make_params(4, 1e-4, 4, BG, sampler=_abi.GIBBS_MOTIF_SAMPLER, cutoff=0.0), 2, seed=5)   # built (see below)
        assert motif.sites.shape == (2, 2)
        with pytest.raises(_abi.GibbsUnsupportedError):           # two sites per sequence need ACGT-only sequences
            eng.run(make_params(4, 1e-4, 4, BG, sampler=_abi.GIBBS_MOTIF_SAMPLER, cutoff=0.0, motif_amount=2), 1)
    with GibbsEngine([b"ACGTACGT", b"ACGTACGT"]) as eng:          # no masked symbol: everything stays available
        eng.run(make_params(4, 1e-4, 5, BG, background=_abi.GIBBS_BG_DATA), 1)


@pytest.mark.parametrize("team", TEAMS)
def test_row_without_a_valid_window(team):
    """Every window of sequence 1 holds an N: its scan returns (log2 0, 0) = (-inf, 0) (fs:302-314 from (0., 0)),
    and site 0 then covers masked bases, which the other sequences' counts must skip."""
    seqs = [b"ACGTACGTTGCAAC", b"ANNANNCNNGNNTN", b"TTGCAACGTACGTA", b"ACGGTACGTTGCAT", b"CCGTACGTAGCAAC"]
    k = 3
    S = O.sources(seqs)
    pcv = O.pcv_from_acgt(BG)
    with GibbsEngine(seqs) as eng:
        eng.set_team_warps(team)
        res = eng.run(make_params(k, 1e-4, 5, BG), 4, chain_id_base=2, seed=5)
    for c in range(4):
        rng, keep = O.make_rng(seed=5, chain=2 + c)
        score, pos, _ = O.site_step("do_site_sampling_with_bpv", S, k, 1e-4, pcv=pcv, rng=rng)
        assert score[1] == -np.inf and pos[1] == 0
        assert res.sites[c].tolist() == pos.tolist()
        assert res.scores[c][1] == -np.inf
        ok = np.isfinite(score)
        np.testing.assert_allclose(res.scores[c][ok], score[ok], rtol=LOG2_RTOL)
        assert res.sums[c] == -np.inf


DATA_CASES = [  # (seed, n, lo, hi, k, other symbols, fraction, alphabet_size)
    (11, 6, 30, 60, 6, "NRY*", 0.05, 5),
    (12, 9, 40, 90, 12, "N", 0.02, 5),
    (13, 5, 25, 40, 7, "N-W", 0.10, 4),      # Gap outside a 4-symbol alphabet: a dead row like N
    (14, 8, 200, 300, 16, "NB", 0.01, 5),
    (15, 4, 12, 20, 3, "N", 0.30, 5),        # heavy masking
]


@pytest.mark.parametrize("case", DATA_CASES, ids=lambda c: f"s{c[0]}_n{c[1]}_k{c[4]}")
def test_data_derived_background_with_masked_symbols(case):
    """doSiteSampling (fs:697) on sequences with symbols outside the alphabet: zero-score windows, uncounted site
    bases, and the dead rows that the held-out sequence adds to the background denominator (fs:471, fs:116)."""
    seed, n, lo, hi, k, sym, frac, alen = case
    seqs = _seqs(seed, n, lo, hi, sym, frac)
    S = O.sources(seqs)
    alphabet = _alphabet(alen)
    params = make_params(k, 1e-4, alen, BG, background=_abi.GIBBS_BG_DATA)
    with GibbsEngine(seqs) as eng:
        res = eng.run(params, 4, chain_id_base=60, seed=5 + seed, want_counts=False)
        u = np.random.default_rng(seed).random((1, draws_per_chain(n)))
        inj = eng.run(params, 1, uniforms=u, want_counts=False)
    for c in range(4):
        rng, keep = O.make_rng(seed=5 + seed, chain=60 + c)
        score, pos, _ = O.site_step("do_site_sampling", S, k, 1e-4, rng=rng, alphabet=alphabet)
        assert res.sites[c].tolist() == pos.tolist(), f"chain {c}"
        finite = np.isfinite(score)
        assert np.array_equal(np.isfinite(res.scores[c]), finite)
        np.testing.assert_allclose(res.scores[c][finite], score[finite], rtol=LOG2_RTOL)
    rng, keep = O.make_rng(uniforms=u[0])
    score, pos, _ = O.site_step("do_site_sampling", S, k, 1e-4, rng=rng, alphabet=alphabet)
    assert inj.sites[0].tolist() == pos.tolist()


MOTIF_CASES = [  # (seed, n, lo, hi, k, other symbols, fraction, alphabet_size, pc, cutoff)
    (11, 6, 30, 60, 6, "NRY*", 0.05, 5, 1e-4, 1.0),
    (12, 9, 50, 90, 8, "N", 0.02, 5, 1e-2, 0.0),      # low cut-off: long candidate lists next to masked windows
    (13, 5, 25, 40, 7, "N-W", 0.10, 4, 1e-4, 1.0),    # Gap outside an alphabet of 4
    (14, 8, 200, 260, 12, "NB", 0.006, 5, 1e-4, 2.0), # 16-window chunks: unmasked rows take the ranking pass, masked ones the list
    (15, 4, 12, 20, 3, "N", 0.30, 5, 1.0, -1.0),      # heavy masking, negative cut-off (the sequential roulette walk)
]


def _motif_same(got_sites, got_scores, want):
    assert [([int(x)] if x >= 0 else []) for x in got_sites] == [p for _, p in want]
    np.testing.assert_allclose(got_scores, [v for v, _ in want], rtol=LOG2_RTOL)


@pytest.mark.parametrize("team", [1, 4])
@pytest.mark.parametrize("case", MOTIF_CASES, ids=lambda c: f"s{c[0]}_n{c[1]}_k{c[4]}")
def test_motif_sampler_with_masked_symbols(case, team):
    """doMotifSamplingWithPCV (fs:876) over sequences that hold symbols outside the alphabet: a window over one is never a
    candidate (log2 0 = -inf), its background entry has probability 0, a site that covers one counts it in a dead row."""
    seed, n, lo, hi, k, sym, frac, alen, pc, cutoff = case
    seqs = _seqs(seed, n, lo, hi, sym, frac)
    S = O.sources(seqs)
    pcv = O.pcv_from_acgt(BG)
    alphabet = _alphabet(alen)
    n_chains = 4
    with GibbsEngine(seqs) as eng:
        eng.set_team_warps(team)
        params = make_params(k, pc, alen, BG, cutoff=cutoff, sampler=_abi.GIBBS_MOTIF_SAMPLER)
        res = eng.run(params, n_chains, chain_id_base=11, seed=90 + seed, want_counts=False)
        for c in range(n_chains):
            rng, _ = O.make_rng(seed=90 + seed, chain=11 + c)
            want, _ = O.motif_step("do_motif_sampling", 0, S, 1, k, pc, cutoff, pcv=pcv, rng=rng, alphabet=alphabet)
            _motif_same(res.sites[c], res.scores[c], want)
        # the roulette primitive on states that include site-less sequences and sites over masked symbols
        rng_np = np.random.default_rng(seed)
        for trial in range(3):
            sites = np.array([rng_np.integers(0, len(q) - k + 1) for q in seqs], dtype=np.int32)
            if trial == 2:
                sites[0] = -1
            for h in (0, n - 1):
                pfm = np.zeros((O.NSLOT, k), dtype=np.int32)
                for i, p0 in enumerate(sites):
                    if i != h and p0 >= 0:
                        for j in range(k):
                            pfm[seqs[i][p0 + j] - 42, j] += 1
                ppm = O.ppm_of_pfm(pfm, n - 1, pc, alphabet)
                cand = O.candidates(seqs[h], k, 1, cutoff, pcv, ppm, alphabet)
                weights = [cd[0] for cd in cand]
                if not (sum(weights) > 0.0) or min(weights) < 0.0:
                    continue                                   # (no mass / negative weights: the walk is exercised by the restarts above)
                for u in (0.0, 0.003, 0.5, 0.999, float(rng_np.random())):
                    idx = O.roulette(weights, u)
                    pwms, site = eng.pick_roulette(sites, h, params, u)
                    assert ([site] if site >= 0 else []) == cand[idx][1], (trial, h, u)
                    assert pwms == pytest.approx(cand[idx][0], rel=LOG2_RTOL)


@pytest.mark.parametrize("team", [1, 4])
@pytest.mark.parametrize("case", MOTIF_CASES[:4], ids=lambda c: f"s{c[0]}_n{c[1]}_k{c[4]}")
def test_data_background_motif_sampler_with_masked_symbols(case, team):
    """doMotifSampling (fs:1034): the background of a held-out sequence takes the others' alphabet symbols outside their
    sites (fuse keeps alphabet rows, fs:67-69) plus EVERY symbol of the held-out sequence (fs:953) -- its symbols outside
    the alphabet count in the denominator (fs:117)."""
    seed, n, lo, hi, k, sym, frac, alen, pc, cutoff = case
    seqs = _seqs(seed, n, lo, hi, sym, frac)
    S = O.sources(seqs)
    alphabet = _alphabet(alen)
    n_chains = 4
    with GibbsEngine(seqs) as eng:
        eng.set_team_warps(team)
        params = make_params(k, pc, alen, [0.25] * 4, cutoff=cutoff, sampler=_abi.GIBBS_MOTIF_SAMPLER, background=_abi.GIBBS_BG_DATA)
        res = eng.run(params, n_chains, chain_id_base=3, seed=40 + seed, want_counts=False)
        for c in range(n_chains):
            rng, _ = O.make_rng(seed=40 + seed, chain=3 + c)
            want, _ = O.motif_step("do_motif_sampling", 1, S, 1, k, pc, cutoff, rng=rng, alphabet=alphabet)
            _motif_same(res.sites[c], res.scores[c], want)


def test_motif_sampler_masked_set_equals_clean_set_when_masks_are_never_touched():
    """A set whose only symbol outside the alphabet sits in ONE sequence: every other sequence's updates must equal what the
    plain instantiation computes on the same sequences -- checked through the whole-restart result against the oracle on a
    larger shape (1 masked base in 40 x 150 bp)."""
    rng = np.random.default_rng(5)
    seqs = ["".join(rng.choice(list("ACGT"), size=150)).encode() for _ in range(40)]
    s7 = bytearray(seqs[7]); s7[75] = ord("N"); seqs[7] = bytes(s7)
    S = O.sources(seqs)
    pcv = O.pcv_from_acgt(BG)
    k, pc, cutoff = 10, 1e-4, 1.0
    with GibbsEngine(seqs) as eng:
        res = eng.run(make_params(k, pc, 5, BG, cutoff=cutoff, sampler=_abi.GIBBS_MOTIF_SAMPLER), 3, seed=8, want_counts=False)
    for c in range(3):
        r, _ = O.make_rng(seed=8, chain=c)
        want, _ = O.motif_step("do_motif_sampling", 0, S, 1, k, pc, cutoff, pcv=pcv, rng=r)
        _motif_same(res.sites[c], res.scores[c], want)
