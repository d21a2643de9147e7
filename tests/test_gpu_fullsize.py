"""Full BASELINE.json sizes on the GPU, against the oracle. The oracle's faithful mode needs hours at these sizes; its
incremental mode (same float64 operations on the same integer counts, proven bit-identical on CPU by
tests/test_oracle_fast.py) generated tests/golden/fullsize_v1.json: every one of the 1024 chains of C2, four chains of
C3, one whole restart of C4 with phase shifts and its random starts. The same mode also runs live here on a sample of
chains, and the primitives are checked at full N against the faithful functions. Beside that: size-independent
properties (determinism, independence of batching / sharding / init path, idempotence of a converged phase, site
ranges, sums = left-to-right sums of the scores, best chain = first largest sum, PWM counts = recount of the sites)."""
import hashlib
import json
import os

import numpy as np
import pytest

import oracle_lib as O
from gibbssampling_b200 import SiteSampler, _abi
from gibbssampling_b200.engine import GibbsEngine, make_params
from gibbssampling_b200.synthetic import background_of, planted_motif_set

pytestmark = pytest.mark.gpu
PC, ALEN = 1e-4, 5


def _recount(ps, sites, k):
    codes = np.zeros(256, dtype=np.int64)
    codes[ord("C")], codes[ord("G")], codes[ord("T")] = 1, 2, 3
    idx = (ps.offsets[:-1] + sites)[:, None] + np.arange(k)[None, :]
    b = codes[ps.ascii[idx]]
    out = np.zeros((k, 4), dtype=np.int64)
    for j in range(k):
        out[j] = np.bincount(b[:, j], minlength=4)
    return out


def _check_invariants(ps, res, k):
    lens = np.diff(ps.offsets)
    assert (res.sites >= 0).all() and (res.sites <= (lens - k)[None, :]).all()
    for c in range(res.sites.shape[0]):
        acc = 0.0
        for v in res.scores[c]:
            acc = acc + float(v)
        assert res.sums[c] == acc                       # Array.sum order (fs:445)
    best = 0
    for c in range(1, len(res.sums)):
        if res.sums[c] > res.sums[best]:
            best = c
    assert res.best_chain == best                       # strict >, first wins (fs:450)
    assert res.counts.tolist() == _recount(ps, res.sites[best], k).tolist()
    assert res.counts.sum() == k * ps.n
    assert res.stats["capped_chains"] == 0 and res.stats["fast_path"] == 1


def _golden():
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fullsize_v1.json")) as f:
        return json.load(f)


def _digest(sites) -> str:
    return hashlib.sha1(np.ascontiguousarray(sites, dtype="<i4").tobytes()).hexdigest()[:20]


def _assert_matches_golden(res, g, chains):
    """sites bit-exact (digest), sums and stored scores within the 1e-5 of north_star (device log vs glibc log differ
    in the last bit at most; the sums agree far tighter than that), counters exact."""
    assert [_digest(res.sites[c]) for c in range(chains)] == g["sites_sha1"]
    np.testing.assert_allclose(res.sums, [float.fromhex(x) for x in g["sums"]], rtol=1e-12)
    head = np.array([[float.fromhex(x) for x in row] for row in g["scores_head"]])
    np.testing.assert_allclose(res.scores[:, :4], head, rtol=1e-5)
    assert res.stats["sweeps"] == g["sweeps"] and res.stats["site_updates"] == g["site_updates"]
    assert res.stats["window_scores"] == g["window_scores"]


def test_c2_every_chain_matches_the_oracle():
    """BASELINE configs[1]: 1000 x 500 bp, k = 12, ALL 1024 chains of the benchmarked step against the oracle
    (golden file generated with the oracle's incremental mode, tests/golden/make_fullsize_golden.py), and a live
    sample of chains against the oracle running on this box."""
    g = _golden()["C2"]
    n, L, k, chains = g["n"], g["L"], g["k"], g["n_chains"]
    assert (n, L, k, chains) == (1000, 500, 12, 1024)
    ps = planted_motif_set(n, L, k)
    seqs = ps.sequences()
    bg = background_of(ps.ascii, PC, ALEN)
    params = make_params(k, PC, ALEN, bg)
    with GibbsEngine(seqs) as eng:
        res = eng.run(params, chains, chain_id_base=g["chain_base"], seed=g["seed"])
        _check_invariants(ps, res, k)
        _assert_matches_golden(res, g, chains)
        # the planted motif is recovered: the best chain puts most sites on the planted positions
        assert (res.sites[res.best_chain] == ps.truth).mean() > 0.5   # 10 % planted-base mutation; chance windows win some
        # live: the oracle (incremental mode) on this box for a spread of chains, full score vectors
        S = O.sources(seqs)
        pcv = O.pcv_from_acgt(bg)
        live = [0, 1, 2, 3, 255, 256, 517, 1000, 1022, 1023]
        for c in live:
            rng, _ = O.make_rng(seed=g["seed"], chain=g["chain_base"] + c)
            score, pos, st, _ = O.fast_site_pipeline(0, S, k, PC, pcv=pcv, rng=rng)
            assert res.sites[c].tolist() == pos.tolist(), f"chain {c}"
            np.testing.assert_allclose(res.scores[c], score, rtol=1e-5)
        # every init path gives the same chains
        for path in (_abi.GIBBS_INIT_CHAIN, _abi.GIBBS_INIT_WIDE, _abi.GIBBS_INIT_SMEM, _abi.GIBBS_INIT_TILED):
            eng.set_option(_abi.GIBBS_OPT_INIT_PATH, path)
            part = eng.run(params, 160, chain_id_base=g["chain_base"] + 512, seed=g["seed"], want_counts=False)
            assert part.stats["init_path"] == path
            assert part.sites.tobytes() == res.sites[512:672].tobytes()
            assert part.scores.tobytes() == res.scores[512:672].tobytes()
        eng.set_option(_abi.GIBBS_OPT_INIT_PATH, _abi.GIBBS_INIT_AUTO)
        # determinism and independence of batching (what sharding over GPUs relies on)
        again = eng.run(params, chains, chain_id_base=g["chain_base"], seed=g["seed"])
        assert again.sites.tobytes() == res.sites.tobytes() and again.scores.tobytes() == res.scores.tobytes()
        # the restart loop on the device over all 1024 restarts
        best = eng.fetch_best(chains - 1)
        want = SiteSampler.replay_restart_loop(chains - 1, res.scores, res.sites, res.sums)
        assert [(float(a), int(b)) for a, b in zip(best.scores, best.sites)] == want
        # idempotence: the last phase (right shifts) applied to its own output changes nothing
        eng.set_start_state(res.sites[:64], res.scores[:64])
        redo = eng.run(make_params(k, PC, ALEN, bg, phase_mask=_abi.PHASE_RIGHT), 64, seed=g["seed"])
        assert redo.sites.tobytes() == res.sites[:64].tobytes()
        np.testing.assert_allclose(redo.scores, res.scores[:64], rtol=1e-12)
        assert redo.stats["sweeps"] == 64               # one quiet sweep per chain


def test_c3_chains_match_the_oracle():
    """BASELINE configs[2]: 10k promoter-length (1 kb) sequences, k = 16; four of the 8192 restarts, end to end."""
    g = _golden()["C3"]
    n, L, k, chains = g["n"], g["L"], g["k"], g["n_chains"]
    assert (n, L, k) == (10000, 1000, 16) and chains >= 4
    ps = planted_motif_set(n, L, k)
    bg = background_of(ps.ascii, PC, ALEN)
    params = make_params(k, PC, ALEN, bg)
    with GibbsEngine(ps.sequences()) as eng:
        res = eng.run(params, chains, chain_id_base=g["chain_base"], seed=g["seed"])
        _check_invariants(ps, res, k)
        _assert_matches_golden(res, g, chains)
        solo = eng.run(params, 1, chain_id_base=g["chain_base"] + 2, seed=g["seed"])
        assert solo.sites[0].tobytes() == res.sites[2].tobytes() and solo.scores[0].tobytes() == res.scores[2].tobytes()
        eng.set_team_warps(4)
        narrow = eng.run(params, chains, chain_id_base=g["chain_base"], seed=g["seed"])
        assert narrow.sites.tobytes() == res.sites.tobytes() and narrow.scores.tobytes() == res.scores.tobytes()
        assert (res.sites[res.best_chain] == ps.truth).mean() > 0.5   # 10 % planted-base mutation; chance windows win some
        # one chain live against the oracle on this box (about 10 s of CPU)
        S = O.sources(ps.sequences())
        rng, _ = O.make_rng(seed=g["seed"], chain=g["chain_base"] + 1)
        score, pos, st, _ = O.fast_site_pipeline(0, S, k, PC, pcv=O.pcv_from_acgt(bg), rng=rng, threads=os.cpu_count() or 1)
        assert res.sites[1].tolist() == pos.tolist()
        np.testing.assert_allclose(res.scores[1], score, rtol=1e-5)


def test_c4_restart_and_primitives_match_the_oracle():
    """BASELINE configs[3]: 100k ChIP-seq-peak-sized (200 bp) sequences, k = 20, with phase-shift moves. One restart
    draws N(N-1) ~ 1e10 initial sites, so the draw counter leaves 32 bits. The random starts and the whole restart are
    compared with the oracle's golden results; the primitives are compared live at full N."""
    g = _golden()["C4"]
    n, L, k = g["n"], g["L"], g["k"]
    assert (n, L, k) == (100000, 200, 20)
    ps = planted_motif_set(n, L, k)
    seqs = ps.sequences()
    bg = background_of(ps.ascii, PC, ALEN)
    S = O.sources(seqs)
    pcv = O.pcv_from_acgt(bg)
    with GibbsEngine(seqs) as eng:
        # the random starts alone: sites are those of the Philox draws n(N-1) + rank (fs:418-426)
        init = eng.run(make_params(k, PC, ALEN, bg, phase_mask=_abi.PHASE_INIT), 1, chain_id_base=g["chain"], seed=g["seed"])
        assert init.stats["site_updates"] == n and init.stats["sweeps"] == 1
        assert _digest(init.sites[0]) == g["init_sites_sha1"]
        for i, want in g["init_sites_sample"].items():
            assert int(init.sites[0][int(i)]) == want
        np.testing.assert_allclose(init.sums[0], float.fromhex(g["init_sum"]), rtol=1e-12)
        np.testing.assert_allclose(init.scores[0][:4], [float.fromhex(x) for x in g["init_scores_head"]], rtol=1e-5)
        # one whole restart with phase shifts (and a second chain beside it, so batching is exercised too)
        res = eng.run(make_params(k, PC, ALEN, bg), 2, chain_id_base=g["chain"], seed=g["seed"])
        _check_invariants(ps, res, k)
        assert _digest(res.sites[0]) == g["sites_sha1"]
        np.testing.assert_allclose(res.sums[0], float.fromhex(g["sum"]), rtol=1e-12)
        np.testing.assert_allclose(res.scores[0][:4], [float.fromhex(x) for x in g["scores_head"]], rtol=1e-5)
        assert (res.sites[res.best_chain] == ps.truth).mean() > 0.5   # 10 % planted-base mutation; chance windows win some
        # primitives at full N against the faithful oracle functions, for several held-out sequences and two states
        params = make_params(k, PC, ALEN, bg)
        for sites in (init.sites[0], res.sites[0]):
            for h in (0, 1, 54321, n - 1):
                pfm = O.loo_pfm(S, sites, h, k)
                assert eng.loo_counts(sites, h, k).tolist() == O.acgt_counts(pfm).tolist()
                ppm = O.ppm_of_pfm(pfm, n - 1, PC)
                want_raw = O.window_scores_bpv(S.seq(h), k, pcv, ppm)
                raw, lg = eng.window_scores(sites, h, params)
                assert raw.tobytes() == want_raw.tobytes()                       # float64 products, bit for bit
                np.testing.assert_allclose(lg, np.log(want_raw) / np.log(2.0), rtol=1e-5)
                want_score, want_pos = O.best_pwms_with_bpv(S.seq(h), k, pcv, ppm)
                score, pos = eng.pick_argmax(sites, h, params)
                assert pos == want_pos and score == pytest.approx(want_score, rel=1e-5)


@pytest.mark.parametrize("L", [100, 1000, 10000])
@pytest.mark.parametrize("k", [6, 12, 20, 30])
def test_c5_sweep_of_widths_and_lengths_matches_oracle(k, L):
    """BASELINE configs[4]: k in {6, 12, 20, 30} x L in {100 bp, 1 kb, 10 kb}. Few sequences so that the oracle's
    from-scratch rebuilds finish in seconds; every chunk geometry of the ranking pass (4 / 8 / 16 windows per lane,
    several rounds per lane at 10 kb) and 2- and 3-word k-mers are covered."""
    n = 6
    ps = planted_motif_set(n, L, k, seed=900 + k + L)
    seqs = ps.sequences()
    bg = background_of(ps.ascii, PC, ALEN)
    S = O.sources(seqs)
    pcv = O.pcv_from_acgt(bg)
    with GibbsEngine(seqs) as eng:
        res = eng.run(make_params(k, PC, ALEN, bg), 2, chain_id_base=50, seed=k + L)
    _check_invariants(ps, res, k)
    for c in range(2):
        rng, keep = O.make_rng(seed=k + L, chain=50 + c)
        score, pos, _ = O.site_step("do_site_sampling_with_bpv", S, k, PC, pcv=pcv, rng=rng)
        assert res.sites[c].tolist() == pos.tolist(), f"chain {c}"
        np.testing.assert_allclose(res.scores[c], score, rtol=1e-5)


def test_c5_sweep_of_chain_counts():
    """BASELINE configs[4]: 1 ... 65536 chains (powers of 4). A chain's result depends on (seed, chain id) only,
    whatever the launch shape: all three hand-over stages, the grid-wide random starts and one-wave launches agree."""
    n, L, k = 24, 100, 8
    ps = planted_motif_set(n, L, k, seed=31)
    seqs = ps.sequences()
    bg = background_of(ps.ascii, PC, ALEN)
    params = make_params(k, PC, ALEN, bg)
    with GibbsEngine(seqs) as eng:
        big = eng.run(params, 65536, seed=17)
        _check_invariants(ps, big, k)
        assert big.stats["site_updates"] >= 65536 * n * 4
        for chains in (1, 4, 16, 64, 256, 1024, 4096, 16384):
            res = eng.run(params, chains, seed=17, want_counts=False)
            assert res.sites.tobytes() == big.sites[:chains].tobytes(), chains
            assert res.scores.tobytes() == big.scores[:chains].tobytes(), chains
            tail = eng.run(params, 3, chain_id_base=chains - 1, seed=17, want_counts=False)   # an unaligned slice
            assert tail.sites[0].tobytes() == big.sites[chains - 1].tobytes()
    S = O.sources(seqs)
    pcv = O.pcv_from_acgt(bg)
    for c in (0, 1, 4095, 65535):
        rng, keep = O.make_rng(seed=17, chain=c)
        score, pos, _ = O.site_step("do_site_sampling_with_bpv", S, k, PC, pcv=pcv, rng=rng)
        assert big.sites[c].tolist() == pos.tolist(), f"chain {c}"
