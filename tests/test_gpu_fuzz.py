"""Differential fuzz: random shapes, widths, pseudocounts, alphabets, symbols, backgrounds, team sizes and phase masks,
GPU (through the C ABI) against the oracle. Seeds are fixed, so a failure names a reproducible case."""
import numpy as np
import pytest

import oracle_lib as O
from gibbssampling_b200 import _abi
from gibbssampling_b200.engine import GibbsEngine, draws_per_chain, make_params

pytestmark = pytest.mark.gpu
LOG2_RTOL = 1e-5


def _case(seed):
    rng = np.random.default_rng(10_000 + seed)
    n = int(rng.integers(1, 14))
    k = int(rng.choice([1, 2, 3, 5, 6, 7, 8, 11, 12, 13, 16, 17, 20, 24, 29, 31, 32]))
    lo = k + int(rng.integers(0, 4))
    hi = lo + int(rng.choice([0, 3, 20, 90, 400]))
    alen = int(rng.choice([4, 5]))
    masked = rng.random() < 0.35
    symbols = "N*RY" + ("-" if alen == 4 else "")
    seqs = []
    for _ in range(n):
        L = int(rng.integers(lo, hi + 1))
        s = rng.choice(list("ACGT"), size=L, p=rng.dirichlet([2.0] * 4))
        if masked:
            m = rng.random(L) < rng.choice([0.01, 0.05, 0.3])
            s[m] = rng.choice(list(symbols), size=int(m.sum()))
        seqs.append("".join(s).encode())
    pc = float(rng.choice([1e-4, 1e-4, 1e-2, 0.5, 1.0, 0.0]))
    data = rng.random() < 0.4
    bg = rng.dirichlet([5.0] * 4).tolist()
    team = int(rng.choice([0, 0, 1, 4, 8, 16]))
    return n, k, alen, seqs, pc, data, bg, team, rng


@pytest.mark.parametrize("seed", range(160))
def test_random_case_matches_oracle(seed):
    n, k, alen, seqs, pc, data, bg, team, rng = _case(seed)
    alphabet = b"ATGC" if alen == 4 else b"ATGC-"
    S = O.sources(seqs)
    pcv = None if data else O.pcv_from_acgt(bg)
    name = "do_site_sampling" if data else "do_site_sampling_with_bpv"
    params = make_params(k, pc, alen, bg, background=_abi.GIBBS_BG_DATA if data else _abi.GIBBS_BG_FIXED)
    n_chains = int(rng.integers(1, 5))
    injected = rng.random() < 0.3
    u = rng.random((n_chains, max(draws_per_chain(n), 1))) if injected else None
    with GibbsEngine(seqs) as eng:
        eng.set_team_warps(team)
        res = eng.run(params, n_chains, chain_id_base=seed, seed=seed * 7 + 1, uniforms=u)
    for c in range(n_chains):
        r, keep = O.make_rng(uniforms=u[c]) if injected else O.make_rng(seed=seed * 7 + 1, chain=seed + c)
        score, pos, st = O.site_step(name, S, k, pc, pcv=pcv, rng=r, alphabet=alphabet)
        assert res.sites[c].tolist() == pos.tolist(), f"chain {c}"
        nan = np.isnan(score)
        assert np.array_equal(np.isnan(res.scores[c]), nan)
        fin = np.isfinite(score)
        np.testing.assert_allclose(res.scores[c][fin], score[fin], rtol=LOG2_RTOL)
        inf = ~fin & ~nan
        assert np.array_equal(res.scores[c][inf], score[inf])


def _motif_case(seed):
    rng = np.random.default_rng(20_000 + seed)
    n = int(rng.integers(2, 10))
    k = int(rng.choice([2, 3, 5, 6, 8, 11, 12, 16, 20]))
    lo = k + int(rng.integers(0, 4))
    hi = lo + int(rng.choice([0, 5, 30, 120]))
    seqs = ["".join(rng.choice(list("ACGT"), size=int(rng.integers(lo, hi + 1)), p=rng.dirichlet([2.0] * 4))).encode()
            for _ in range(n)]
    pc = float(rng.choice([1e-4, 1e-2, 0.5]))
    cutoff = float(rng.choice([-5.0, 0.0, 1.0, 4.0, 50.0]))
    return n, k, seqs, pc, cutoff, rng.random() < 0.5, rng.dirichlet([5.0] * 4).tolist(), rng


@pytest.mark.parametrize("seed", range(60))
def test_random_motif_sampler_case_matches_oracle(seed):
    """MotifSampler m = 1 (doMotifSamplingWithPCV fs:876 / doMotifSampling fs:1034): sequential roulette walk over
    [W background entries] ++ [candidates above the cut-off], then greedy sweeps."""
    n, k, seqs, pc, cutoff, data, bg, rng = _motif_case(seed)
    S = O.sources(seqs)
    params = make_params(k, pc, 5, bg, cutoff=cutoff, sampler=_abi.GIBBS_MOTIF_SAMPLER,
                         background=_abi.GIBBS_BG_DATA if data else _abi.GIBBS_BG_FIXED)
    with GibbsEngine(seqs) as eng:
        try:
            res = eng.run(params, 3, chain_id_base=seed, seed=seed + 3, want_counts=False)
            failed = None
        except _abi.GibbsRouletteError as e:     # a pick beyond the accumulated mass: the reference throws too (fs:753)
            failed = e
    for c in range(3):
        r, keep = O.make_rng(seed=seed + 3, chain=seed + c)
        try:
            want, st = O.motif_step("do_motif_sampling", 1 if data else 0, S, 1, k, pc, cutoff,
                                    pcv=None if data else O.pcv_from_acgt(bg), rng=r)
        except O.OracleError as e:
            assert failed is not None and e.code == O.ERR_ROULETTE
            return
        assert failed is None
        got_pos = [[int(p)] if p >= 0 else [] for p in res.sites[c]]
        assert got_pos == [list(p) for _, p in want], f"chain {c}"
        np.testing.assert_allclose(res.scores[c], [s for s, _ in want], rtol=1e-5)


def _masked_motif_case(seed):
    rng = np.random.default_rng(30_000 + seed)
    n = int(rng.integers(2, 10))
    k = int(rng.choice([2, 3, 5, 6, 8, 11, 12, 16, 20]))
    lo = k + int(rng.integers(0, 4))
    hi = lo + int(rng.choice([0, 5, 30, 120, 300]))
    alen = int(rng.choice([4, 5]))
    symbols = "N*RY" + ("-" if alen == 4 else "")
    frac = float(rng.choice([0.005, 0.02, 0.1, 0.3]))
    seqs = []
    for _ in range(n):
        L = int(rng.integers(lo, hi + 1))
        q = rng.choice(list("ACGT"), size=L, p=rng.dirichlet([2.0] * 4))
        m = rng.random(L) < frac
        q[m] = rng.choice(list(symbols), size=int(m.sum()))
        seqs.append("".join(q).encode())
    pc = float(rng.choice([1e-4, 1e-2, 0.5]))
    cutoff = float(rng.choice([-5.0, 0.0, 1.0, 4.0]))
    team = int(rng.choice([0, 1, 4]))
    return n, k, alen, seqs, pc, cutoff, rng.random() < 0.5, rng.dirichlet([5.0] * 4).tolist(), team, rng


@pytest.mark.parametrize("seed", range(60))
def test_random_masked_motif_sampler_case_matches_oracle(seed):
    """The MotifSampler (m = 1, both backgrounds) over sequences with symbols outside the alphabet."""
    n, k, alen, seqs, pc, cutoff, data, bg, team, rng = _masked_motif_case(seed)
    alphabet = b"ATGC" if alen == 4 else b"ATGC-"
    S = O.sources(seqs)
    params = make_params(k, pc, alen, bg, cutoff=cutoff, sampler=_abi.GIBBS_MOTIF_SAMPLER,
                         background=_abi.GIBBS_BG_DATA if data else _abi.GIBBS_BG_FIXED)
    with GibbsEngine(seqs) as eng:
        eng.set_team_warps(team)
        try:
            res = eng.run(params, 3, chain_id_base=seed, seed=seed + 5, want_counts=False)
            failed = None
        except _abi.GibbsRouletteError as e:     # a pick beyond the accumulated mass (e.g. every window masked: 0 / 0)
            failed = e
    for c in range(3):
        r, keep = O.make_rng(seed=seed + 5, chain=seed + c)
        try:
            want, st = O.motif_step("do_motif_sampling", 1 if data else 0, S, 1, k, pc, cutoff,
                                    pcv=None if data else O.pcv_from_acgt(bg), rng=r, alphabet=alphabet)
        except O.OracleError as e:
            assert failed is not None and e.code == O.ERR_ROULETTE
            return
        assert failed is None
        got_pos = [[int(p)] if p >= 0 else [] for p in res.sites[c]]
        assert got_pos == [list(p) for _, p in want], f"chain {c}"
        want_s = np.array([v for v, _ in want])
        fin = np.isfinite(want_s)
        np.testing.assert_allclose(res.scores[c][fin], want_s[fin], rtol=1e-5)
        assert np.array_equal(res.scores[c][~fin], want_s[~fin], equal_nan=True)
