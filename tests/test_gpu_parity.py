"""GPU parity: the sm_100a path, called through the C ABI (ctypes), against the CPU oracle.

Bars (BASELINE.json north_star): PWM counts bit-exact; raw float64 window products bit-exact
(same IEEE operations in the same order); log2 window scores within 1e-5 relative; picked sites
bit-exact when both sides consume the same uniform stream.
"""
import numpy as np
import pytest

import oracle_lib as O
from gibbssampling_b200 import _abi
from gibbssampling_b200.engine import GibbsEngine, draws_per_chain, make_params
from gibbssampling_b200.synthetic import background_of, planted_motif_set

pytestmark = pytest.mark.gpu

LOG2_RTOL = 1e-5  # north_star: "window log-odds agree within 1e-5 relative"


def _cases():
    # (n_seqs, max_len, min_len or None, k, alphabet_size, pc, seed)
    return [
        (4, 21, None, 6, 5, 1e-4, 1),
        (20, 100, None, 8, 5, 1e-4, 2),       # BASELINE config C1 shape
        (12, 60, 33, 7, 5, 1e-4, 3),          # ragged, odd k
        (9, 300, 120, 12, 4, 1e-2, 4),        # CH=8/16 mix, alphabet of 4
        (6, 700, None, 16, 5, 1e-4, 5),
        (5, 150, 90, 20, 5, 1e-3, 6),
        (5, 200, None, 31, 5, 1e-4, 7),       # 3-word k-mers, odd k
        (7, 130, None, 32, 5, 1e-4, 8),
        (8, 40, 12, 1, 5, 1e-4, 9),           # k = 1
        (3, 50, None, 2, 5, 1.0, 10),
    ]


def _setup(case):
    n, L, Lmin, k, alen, pc, seed = case
    ps = planted_motif_set(n, L, k, seed=seed, min_length=Lmin)
    seqs = ps.sequences()
    bg = background_of(ps.ascii, pc, alen)
    alphabet = b"ATGC-"[:alen] if alen <= 5 else b"ATGC-N"
    return ps, seqs, bg, alphabet


def _oracle_site_update(S, seqs, sites, h, k, pc, bg, alphabet):
    pfm = O.loo_pfm(S, sites, h, k)
    ppm = O.ppm_of_pfm(pfm, S.n - 1, pc, alphabet)
    pcv = O.pcv_from_acgt(bg)
    raw = O.window_scores_bpv(seqs[h], k, pcv, ppm, alphabet)
    score, pos = O.best_pwms_with_bpv(seqs[h], k, pcv, ppm, alphabet)
    return O.acgt_counts(pfm), raw, score, pos


TEAMS = [1, 4, 8, 16]  # warps per chain: every kernel variant must be bit-identical to the oracle


@pytest.mark.parametrize("team", TEAMS)
@pytest.mark.parametrize("case", _cases(), ids=lambda c: f"n{c[0]}_L{c[1]}_k{c[3]}")
def test_primitives_match_oracle(case, team):
    ps, seqs, bg, alphabet = _setup(case)
    n, L, Lmin, k, alen, pc, seed = case
    S = O.sources(seqs)
    rng = np.random.default_rng(seed)
    with GibbsEngine(seqs) as eng:
        eng.set_team_warps(team)
        params = make_params(k, pc, alen, bg)
        for trial in range(3):
            sites = np.array([rng.integers(0, len(s) - k + 1) for s in seqs], dtype=np.int32)
            if trial == 2:
                sites = ps.truth.copy()
            for h in sorted({0, n // 2, n - 1}):
                counts, raw, score, pos = _oracle_site_update(S, seqs, sites, h, k, pc, bg, alphabet)
                got_counts = eng.loo_counts(sites, h, k)
                assert got_counts.tolist() == counts.tolist()                      # bit-exact integers
                got_raw, got_log2 = eng.window_scores(sites, h, params)
                assert got_raw.tobytes() == raw.tobytes()                          # bit-exact float64 products
                np.testing.assert_allclose(got_log2, np.log(raw) / np.log(2.0), rtol=LOG2_RTOL)
                got_score, got_pos = eng.pick_argmax(sites, h, params)
                assert got_pos == pos
                assert got_score == pytest.approx(score, rel=LOG2_RTOL)


@pytest.mark.parametrize("team", TEAMS)
def test_argmax_first_maximum_on_ties(team):
    # identical k-mers everywhere: every window ties; the reference keeps the FIRST strict maximum (fs:312)
    seqs = [b"ACGTACGTACGTACGTACGTACGTACGT", b"ACGTACGTACGTACGTACGTACGT", b"ACGTACGTACGTACGTACGT", b"AAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAA"]
    k, pc, alen = 4, 1e-4, 5
    bg = [0.25, 0.25, 0.25, 0.25]
    S = O.sources(seqs)
    with GibbsEngine(seqs) as eng:
        eng.set_team_warps(team)
        params = make_params(k, pc, alen, bg)
        for sites in ([0, 0, 0, 0], [4, 8, 12, 3], [1, 2, 3, 0]):
            for h in range(len(seqs)):
                _, raw, score, pos = _oracle_site_update(S, seqs, sites, h, k, pc, bg, b"ATGC-")
                assert eng.pick_argmax(sites, h, params)[1] == pos
                assert eng.window_scores(sites, h, params)[0].tobytes() == raw.tobytes()


@pytest.mark.parametrize("team", TEAMS)
def test_zero_pseudocount_uses_exact_path(team):
    # pc = 0 makes odds ratios 0 (log2 = -inf): the ranking pass is disabled and every window is
    # scored in float64; all-zero scores give (-inf, 0) like the reference
    seqs = [b"ACGTTGCAACGT", b"TTTTTTTTTTTT", b"ACGTACGTACGT", b"GGGGGGGGGGGG"]
    k, pc, alen = 4, 0.0, 5
    bg = [0.25, 0.25, 0.25, 0.25]
    S = O.sources(seqs)
    with GibbsEngine(seqs) as eng:
        eng.set_team_warps(team)
        params = make_params(k, pc, alen, bg)
        for sites in ([0, 0, 0, 0], [8, 3, 4, 1]):
            for h in range(len(seqs)):
                _, raw, score, pos = _oracle_site_update(S, seqs, sites, h, k, pc, bg, b"ATGC-")
                gs, gp = eng.pick_argmax(sites, h, params)
                assert gp == pos
                assert gs == score or gs == pytest.approx(score, rel=LOG2_RTOL)
                assert eng.window_scores(sites, h, params)[0].tobytes() == raw.tobytes()


def _oracle_chain(S, k, pc, pcv, alphabet, seed=None, chain=0, uniforms=None, name="do_site_sampling_with_bpv", state=None):
    rng, keep = O.make_rng(uniforms=uniforms, seed=seed or 0, chain=chain)
    return O.site_step(name, S, k, pc, pcv=pcv, rng=rng, state=state, alphabet=alphabet)


@pytest.mark.parametrize("team", TEAMS)
@pytest.mark.parametrize("case", _cases(), ids=lambda c: f"n{c[0]}_L{c[1]}_k{c[3]}")
def test_chains_match_oracle_philox(case, team):
    """Whole restarts (fs:691-695), several chains per launch, Philox stream shared with the oracle."""
    ps, seqs, bg, alphabet = _setup(case)
    n, L, Lmin, k, alen, pc, seed = case
    S = O.sources(seqs)
    pcv = O.pcv_from_acgt(bg)
    n_chains, base = 6, 1000 * seed
    with GibbsEngine(seqs) as eng:
        eng.set_team_warps(team)
        res = eng.run(make_params(k, pc, alen, bg), n_chains, chain_id_base=base, seed=0xB200 + seed)
    total_updates = 0
    for c in range(n_chains):
        score, pos, st = _oracle_chain(S, k, pc, pcv, alphabet, seed=0xB200 + seed, chain=base + c)
        assert res.sites[c].tolist() == pos.tolist(), f"chain {c}"
        np.testing.assert_allclose(res.scores[c], score, rtol=LOG2_RTOL)
        assert res.sums[c] == pytest.approx(float(np.sum(score)), rel=1e-9)
        total_updates += st.site_updates
    assert res.stats["site_updates"] == total_updates
    assert res.stats["team_warps"] == team
    best = int(np.argmax(res.sums))
    assert res.best_chain == best
    want_counts = np.zeros((k, 4), dtype=np.int64)
    for i, s in enumerate(seqs):
        for j in range(k):
            want_counts[j, "ACGT".index(chr(s[res.sites[best][i] + j]))] += 1
    assert res.counts.tolist() == want_counts.tolist()


INIT_PATHS = {"chain": _abi.GIBBS_INIT_CHAIN, "wide": _abi.GIBBS_INIT_WIDE, "smem": _abi.GIBBS_INIT_SMEM,
              "tiled": _abi.GIBBS_INIT_TILED}


@pytest.mark.parametrize("path", sorted(INIT_PATHS))
@pytest.mark.parametrize("shape", [(70, 64, 40, 9), (300, 90, None, 12), (1100, 40, 30, 16), (90, 700, 500, 20), (40, 130, None, 31)],
                         ids=lambda s: f"n{s[0]}_L{s[1]}_k{s[3]}")
def test_random_starts_same_on_every_init_path(shape, path):
    """Random starts (fs:412-430) run inside the chain kernel, as the grid-wide kernel gathering from global memory, as
    the grid-wide kernel with the packed set in shared memory, or as the one that streams it through shared memory in
    tiles (gibbs_api.cu picks by shape;
    gibbs_set_option(GIBBS_OPT_INIT_PATH) forces one). All must consume the uniform stream exactly as the oracle does:
    N(N-1) draws, bit-sliced base counters with several nibble spills and (N = 1100) warp flushes, one- and two-word k-mers."""
    n, L, Lmin, k = shape
    ps = planted_motif_set(n, L, k, seed=21, min_length=Lmin)
    seqs = ps.sequences()
    bg = background_of(ps.ascii, 1e-4, 5)
    S = O.sources(seqs)
    pcv = O.pcv_from_acgt(bg)
    n_chains = 3
    with GibbsEngine(seqs) as eng:
        eng.set_option(_abi.GIBBS_OPT_INIT_PATH, INIT_PATHS[path])
        res = eng.run(make_params(k, 1e-4, 5, bg, phase_mask=_abi.PHASE_INIT), n_chains, chain_id_base=7, seed=99)
        full = eng.run(make_params(k, 1e-4, 5, bg), n_chains, chain_id_base=7, seed=99)
    assert res.stats["init_path"] == INIT_PATHS[path]
    assert res.stats["kernel_launches"] >= (1 if path == "chain" else 2)
    for c in range(n_chains):
        score, pos, _ = _oracle_chain(S, k, 1e-4, pcv, b"ATGC-", seed=99, chain=7 + c, name="random_starts_with_bpv")
        assert res.sites[c].tolist() == pos.tolist(), f"chain {c}"
        np.testing.assert_allclose(res.scores[c], score, rtol=LOG2_RTOL)
        if n <= 300 or c == 0:   # the oracle's full restart is O(N^2 L k): seconds at N = 1100
            score, pos, _ = _oracle_chain(S, k, 1e-4, pcv, b"ATGC-", seed=99, chain=7 + c)
            assert full.sites[c].tolist() == pos.tolist(), f"chain {c}"


@pytest.mark.parametrize("tile_rows", [1, 3, 7, 32, 33, 250])
@pytest.mark.parametrize("shape", [(70, 64, 40, 9), (300, 90, None, 12), (1100, 40, 30, 16), (90, 300, 200, 20), (40, 130, None, 31)],
                         ids=lambda s: f"n{s[0]}_L{s[1]}_k{s[3]}")
def test_tiled_random_starts_for_any_tile_size(shape, tile_rows):
    """init_tiled_kernel streams the set through shared memory in tiles; a tile holds a contiguous range of every item's
    Philox stream, cut at arbitrary places (inside Philox blocks, at the held-out sequence, tiles of one sequence, a
    short last tile). GIBBS_OPT_TILE_ROWS forces small tiles; the result must not depend on the tile size: compared
    bit for bit with the chain kernel's own random starts and with the oracle."""
    n, L, Lmin, k = shape
    ps = planted_motif_set(n, L, k, seed=23, min_length=Lmin)
    seqs = ps.sequences()
    bg = background_of(ps.ascii, 1e-4, 5)
    n_chains = 5
    prm = make_params(k, 1e-4, 5, bg, phase_mask=_abi.PHASE_INIT)
    with GibbsEngine(seqs) as eng:
        eng.set_option(_abi.GIBBS_OPT_INIT_PATH, _abi.GIBBS_INIT_CHAIN)
        want = eng.run(prm, n_chains, chain_id_base=2, seed=5)
        eng.set_option(_abi.GIBBS_OPT_INIT_PATH, _abi.GIBBS_INIT_TILED)
        eng.set_option(_abi.GIBBS_OPT_TILE_ROWS, tile_rows)
        got = eng.run(prm, n_chains, chain_id_base=2, seed=5)
    assert got.stats["init_path"] == _abi.GIBBS_INIT_TILED
    assert got.sites.tobytes() == want.sites.tobytes()
    assert got.scores.tobytes() == want.scores.tobytes()
    assert got.stats["site_updates"] == want.stats["site_updates"] == n_chains * n
    S = O.sources(seqs)
    pcv = O.pcv_from_acgt(bg)
    score, pos, _ = _oracle_chain(S, k, 1e-4, pcv, b"ATGC-", seed=5, chain=2, name="random_starts_with_bpv")
    assert got.sites[0].tolist() == pos.tolist()


@pytest.mark.parametrize("team", TEAMS)
def test_chains_match_oracle_injected_uniforms(team):
    """Same comparison with an injected stream of doubles (the parity definition of north_star)."""
    case = (10, 80, 50, 9, 5, 1e-4, 11)
    ps, seqs, bg, alphabet = _setup(case)
    n, k, pc, alen = case[0], case[3], case[5], case[4]
    S = O.sources(seqs)
    pcv = O.pcv_from_acgt(bg)
    n_chains = 4
    rng = np.random.default_rng(5)
    u = rng.random((n_chains, draws_per_chain(n)))
    u[0, :5] = [0.0, 0.999999999999, 0.5, 0.25, 1.0 - 2.0 ** -32]   # range boundaries
    with GibbsEngine(seqs) as eng:
        eng.set_team_warps(team)
        res = eng.run(make_params(k, pc, alen, bg), n_chains, uniforms=u)
    for c in range(n_chains):
        score, pos, _ = _oracle_chain(S, k, pc, pcv, alphabet, uniforms=u[c])
        assert res.sites[c].tolist() == pos.tolist()
        np.testing.assert_allclose(res.scores[c], score, rtol=LOG2_RTOL)


@pytest.mark.parametrize("team", TEAMS)
def test_phases_match_reference_functions(team):
    """Each reference function on its own (fs:412, fs:381, fs:350, fs:318), chained through host state."""
    case = (11, 90, 60, 8, 5, 1e-4, 12)
    ps, seqs, bg, alphabet = _setup(case)
    n, k, pc, alen = case[0], case[3], case[5], case[4]
    S = O.sources(seqs)
    pcv = O.pcv_from_acgt(bg)
    steps = [("random_starts_with_bpv", _abi.PHASE_INIT), ("find_best_motif_with_start_position", _abi.PHASE_GREEDY),
             ("left_shifted_with_bpv", _abi.PHASE_LEFT), ("right_shifted_with_bpv", _abi.PHASE_RIGHT)]
    with GibbsEngine(seqs) as eng:
        eng.set_team_warps(team)
        state = None
        for name, mask in steps:
            o_score, o_pos, _ = _oracle_chain(S, k, pc, pcv, alphabet, seed=99, chain=3, name=name, state=state)
            if state is not None:
                eng.set_start_state(state[1], state[0])
            res = eng.run(make_params(k, pc, alen, bg, phase_mask=mask), 1, chain_id_base=3, seed=99)
            assert res.sites[0].tolist() == o_pos.tolist(), name
            np.testing.assert_allclose(res.scores[0], o_score, rtol=LOG2_RTOL)
            state = (o_score, o_pos)   # feed the ORACLE's state forward: scores only known as log2


def test_chain_results_do_not_depend_on_batching():
    """Chain c gives the same result alone, inside a batch, or on another shard (multi-GPU invariant)."""
    ps = planted_motif_set(30, 120, 10, seed=21)
    seqs = ps.sequences()
    bg = background_of(ps.ascii, 1e-4, 5)
    params = make_params(10, 1e-4, 5, bg)
    with GibbsEngine(seqs) as eng:
        eng.set_team_warps(1)
        solo = eng.run(params, 16, chain_id_base=0, seed=5)
        eng.set_team_warps(4)
        full = eng.run(params, 16, chain_id_base=0, seed=5)
        assert solo.sites.tolist() == full.sites.tolist() and solo.scores.tobytes() == full.scores.tobytes()
        eng.set_team_warps(8)
        wide = eng.run(params, 16, chain_id_base=0, seed=5)
        assert wide.sites.tolist() == full.sites.tolist() and wide.scores.tobytes() == full.scores.tobytes()
        assert wide.stats["site_updates"] == full.stats["site_updates"]
        many8 = eng.run(params, 1500, chain_id_base=0, seed=5)      # more chains than fit at once: several waves
        eng.set_team_warps(4)
        many4 = eng.run(params, 1500, chain_id_base=0, seed=5)
        assert many8.sites[:16].tolist() == full.sites.tolist()
        assert many8.sites.tolist() == many4.sites.tolist() and many8.sums.tobytes() == many4.sums.tobytes()
        eng.set_team_warps(0)
        part = eng.run(params, 5, chain_id_base=9, seed=5)
        assert part.sites.tolist() == full.sites[9:14].tolist()
        assert part.scores.tobytes() == full.scores[9:14].tobytes()
        again = eng.run(params, 16, chain_id_base=0, seed=5)
        assert again.sites.tolist() == full.sites.tolist() and again.sums.tobytes() == full.sums.tobytes()


def test_pinned_fetch_returns_the_same_results():
    """fetch(pinned=True) copies into page-locked buffers from gibbs_host_alloc and reuses them."""
    ps = planted_motif_set(16, 90, 8, seed=3)
    seqs = ps.sequences()
    bg = background_of(ps.ascii, 1e-4, 5)
    p = make_params(8, 1e-4, 5, bg)
    with GibbsEngine(seqs) as eng:
        a = eng.run(p, 5, seed=11)
        b = eng.run(p, 5, seed=11, pinned=True)
        assert a.sites.tolist() == b.sites.tolist() and a.scores.tolist() == b.scores.tolist()
        assert a.sums.tolist() == b.sums.tolist() and a.best_chain == b.best_chain
        first = b.sites.ctypes.data
        keep = b.sites.copy()
        c = eng.run(p, 3, seed=12, pinned=True)            # smaller run: same buffer, overwritten
        assert c.sites.ctypes.data == first and c.sites.shape == (3, 16)
        d = eng.run(p, 9, seed=11, pinned=True)            # larger run: buffer regrown
        assert d.sites[:5].tolist() == keep.tolist()


def test_errors_cross_the_boundary_as_status_codes():
    with pytest.raises(_abi.GibbsSymbolError):
        GibbsEngine([b"ACGT[ACGT", b"ACGTACGT"])    # '[' = 91: outside the 49-slot tables (fs:17-20)
    with pytest.raises(_abi.GibbsSymbolError):
        GibbsEngine([b"ACGT-ACGT", b"acgtacgt"])    # lower case never reaches the tables (the parser upper-cases)
    with GibbsEngine([b"ACGTACGT", b"ACG"]) as eng:
        with pytest.raises(_abi.GibbsShortSequenceError):
            eng.run(make_params(4, 1e-4, 5, [0.25] * 4), 1)
        with pytest.raises(_abi.GibbsArgumentError):
            eng.run(make_params(0, 1e-4, 5, [0.25] * 4), 1)
        with pytest.raises(_abi.GibbsArgumentError):
            eng.run(make_params(33, 1e-4, 5, [0.25] * 4), 1)
        with pytest.raises(_abi.GibbsArgumentError):
            eng.run(make_params(2, 1e-4, 5, [0.25, 0.25, 0.0, 0.5]), 1)
        with pytest.raises(_abi.GibbsArgumentError):
            eng.loo_counts([0, 5], 0, 2)          # site leaves its sequence
        with pytest.raises(_abi.GibbsArgumentError):
            eng.run(make_params(2, 1e-4, 5, [0.25] * 4, phase_mask=_abi.PHASE_GREEDY), 1)  # no start state
        ok = eng.run(make_params(3, 1e-4, 5, [0.25] * 4), 2, seed=1)
        assert ok.sites.shape == (2, 2)
        with pytest.raises(_abi.GibbsUnsupportedError):   # the primitives are the WithBPV functions: fixed background only
            eng.pick_argmax([0, 0], 0, make_params(3, 1e-4, 5, [0.25] * 4, background=_abi.GIBBS_BG_DATA))
        with pytest.raises(_abi.GibbsUnsupportedError):
            eng.window_scores([0, 0], 0, make_params(3, 1e-4, 5, [0.25] * 4, background=_abi.GIBBS_BG_DATA))


def test_single_sequence_and_exact_length():
    # N = 1: no other sequences, the PPM is all pseudocount; L = k: one window
    with GibbsEngine([b"ACGTTGCA"]) as eng:
        r = eng.run(make_params(8, 1e-4, 5, [0.25] * 4), 2, seed=3)
        assert r.sites.tolist() == [[0], [0]]
    seqs = [b"ACGTTGCA", b"TTGCAACG", b"GGGGACGT"]
    S = O.sources(seqs)
    pcv = O.pcv_from_acgt([0.25] * 4)
    with GibbsEngine(seqs) as eng:
        r = eng.run(make_params(8, 1e-4, 5, [0.25] * 4), 1, seed=3)
    score, pos, _ = _oracle_chain(S, 8, 1e-4, pcv, b"ATGC-", seed=3, chain=0)
    assert r.sites[0].tolist() == pos.tolist() == [0, 0, 0]
    np.testing.assert_allclose(r.scores[0], score, rtol=LOG2_RTOL)
