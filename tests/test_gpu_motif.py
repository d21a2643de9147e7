"""GPU parity of the MotifSampler path (fixed pcv, motifAmount = 1): candidate list + roulette pick,
the synchronous stochastic sweep (fs:828), greedy sweeps (fs:788) and the whole restart (fs:876-879)."""
import numpy as np
import pytest

import oracle_lib as O
from gibbssampling_b200 import MotifSampler, _abi
from gibbssampling_b200.CompositeVector import ProbabilityCompositeVector
from gibbssampling_b200.engine import GibbsEngine, draws_per_chain, make_params
from gibbssampling_b200.synthetic import background_of, planted_motif_set

pytestmark = pytest.mark.gpu
RTOL = 1e-5
DNA = list("ATGC-")


def _cases():
    # (n, L, Lmin, k, pc, cutoff, seed)
    return [
        (5, 27, None, 6, 1e-4, 1.0, 1),      # the script's call shape (fsx:407) with m = 1
        (8, 60, 40, 7, 1e-4, 1.0, 2),
        (6, 90, None, 8, 1.0, 0.0, 3),       # large pseudocount, low cut-off: many candidates
        (12, 120, 70, 10, 1e-2, 2.0, 4),
        (4, 200, None, 12, 1e-4, 5.0, 5),    # high cut-off: mostly "no site"
    ]


def _setup(case):
    n, L, Lmin, k, pc, cutoff, seed = case
    ps = planted_motif_set(n, L, k, seed=100 + seed, min_length=Lmin)
    seqs = ps.sequences()
    bg = background_of(ps.ascii, pc, 5)
    return ps, seqs, bg


def _same_state(got_sites, got_scores, want):
    assert [([int(s)] if s >= 0 else []) for s in got_sites] == [p for _, p in want]
    np.testing.assert_allclose(got_scores, [v for v, _ in want], rtol=RTOL)


@pytest.mark.parametrize("case", _cases(), ids=lambda c: f"n{c[0]}_L{c[1]}_k{c[3]}_cut{c[5]}")
def test_roulette_pick_matches_oracle(case):
    n, L, Lmin, k, pc, cutoff, seed = case
    ps, seqs, bg = _setup(case)
    S = O.sources(seqs)
    pcv = O.pcv_from_acgt(bg)
    rng = np.random.default_rng(seed)
    with GibbsEngine(seqs) as eng:
        params = make_params(k, pc, 5, bg, cutoff=cutoff, sampler=_abi.GIBBS_MOTIF_SAMPLER)
        for trial in range(3):
            sites = np.array([rng.integers(0, len(s) - k + 1) for s in seqs], dtype=np.int32) if trial else ps.truth.copy()
            if trial == 2:
                sites[1] = -1                                  # a sequence without a site (Positions = [])
            for h in (0, n - 1):
                pos_lists = [[int(x)] if x >= 0 else [] for x in sites]
                pfm = np.zeros((O.NSLOT, k), dtype=np.int32)
                for i, pl in enumerate(pos_lists):
                    if i != h:
                        for p0 in pl:
                            for j in range(k):
                                pfm[seqs[i][p0 + j] - 42, j] += 1
                ppm = O.ppm_of_pfm(pfm, n - 1, pc)
                cand = O.candidates(seqs[h], k, 1, cutoff, pcv, ppm)
                weights = [c[0] for c in cand]
                for u in (0.0, 1e-7, 0.003, 0.25, 0.5, 0.75, 0.999, float(rng.random())):
                    idx = O.roulette(weights, u)
                    pwms, site = eng.pick_roulette(sites, h, params, u)
                    assert ([site] if site >= 0 else []) == cand[idx][1], (trial, h, u)
                    assert pwms == pytest.approx(cand[idx][0], rel=RTOL)


@pytest.mark.parametrize("case", _cases(), ids=lambda c: f"n{c[0]}_L{c[1]}_k{c[3]}_cut{c[5]}")
def test_motif_restarts_match_oracle(case):
    n, L, Lmin, k, pc, cutoff, seed = case
    ps, seqs, bg = _setup(case)
    S = O.sources(seqs)
    pcv = O.pcv_from_acgt(bg)
    n_chains = 5
    with GibbsEngine(seqs) as eng:
        params = make_params(k, pc, 5, bg, cutoff=cutoff, sampler=_abi.GIBBS_MOTIF_SAMPLER)
        res = eng.run(params, n_chains, chain_id_base=40, seed=77 + seed, want_counts=False)
    for c in range(n_chains):
        rng, _ = O.make_rng(seed=77 + seed, chain=40 + c)
        want, st = O.motif_step("do_motif_sampling", 0, S, 1, k, pc, cutoff, pcv=pcv, rng=rng)
        assert rng.next == draws_per_chain(n, _abi.GIBBS_MOTIF_SAMPLER)
        _same_state(res.sites[c], res.scores[c], want)
    assert res.stats["site_updates"] > 0


@pytest.mark.parametrize("tile_rows", [5, 64])
def test_motif_restarts_on_tiled_random_starts(tile_rows):
    """The MotifSampler's random starts (fs:868) on init_tiled_kernel: same restarts as on the default path."""
    case = _cases()[0]
    n, L, Lmin, k, pc, cutoff, seed = case
    ps, seqs, bg = _setup(case)
    params = make_params(k, pc, 5, bg, cutoff=cutoff, sampler=_abi.GIBBS_MOTIF_SAMPLER)
    with GibbsEngine(seqs) as eng:
        want = eng.run(params, 6, chain_id_base=40, seed=77 + seed, want_counts=False)
        eng.set_option(_abi.GIBBS_OPT_INIT_PATH, _abi.GIBBS_INIT_TILED)
        eng.set_option(_abi.GIBBS_OPT_TILE_ROWS, tile_rows)
        got = eng.run(params, 6, chain_id_base=40, seed=77 + seed, want_counts=False)
    assert got.stats["init_path"] == _abi.GIBBS_INIT_TILED and want.stats["init_path"] != _abi.GIBBS_INIT_TILED
    assert got.sites.tobytes() == want.sites.tobytes()
    assert got.scores.tobytes() == want.scores.tobytes()


def test_motif_phases_and_injected_uniforms():
    case = (7, 80, 50, 8, 1e-4, 1.0, 9)
    n, L, Lmin, k, pc, cutoff, seed = case
    ps, seqs, bg = _setup(case)
    S = O.sources(seqs)
    pcv49 = O.pcv_from_acgt(bg)
    pcv = ProbabilityCompositeVector.ofACGT(*bg)
    u = np.random.default_rng(3).random(draws_per_chain(n, _abi.GIBBS_MOTIF_SAMPLER))
    # whole restart with an injected stream
    got = MotifSampler.doMotifSamplingWithPCV(1, k, pc, cutoff, DNA, seqs, pcv, uniforms=u)
    rng, _ = O.make_rng(uniforms=u)
    want, _ = O.motif_step("do_motif_sampling", 0, S, 1, k, pc, cutoff, pcv=pcv49, rng=rng)
    assert [list(m.Positions) for m in got] == [p for _, p in want]
    np.testing.assert_allclose([m.PWMS for m in got], [v for v, _ in want], rtol=RTOL)
    # the two sweep functions on their own, from a hand-made MotifIndex[] including an empty entry
    start = [MotifSampler.createMotifIndex(1.5, [3]) for _ in range(n)]
    start[2] = MotifSampler.createMotifIndex(1e-6, [])
    ostart = [(m.PWMS, list(m.Positions)) for m in start]
    picks = np.random.default_rng(4).random(n)
    full = np.concatenate([np.zeros(n * (n - 1)), picks])       # draw index of the sweep = N(N-1) + n
    got_s = MotifSampler.findBestMotifPositionsWithStartPositionsByPCV(1, k, pc, cutoff, DNA, seqs, pcv, start, uniforms=full)
    rng, _ = O.make_rng(uniforms=picks)
    want_s, _ = O.motif_step("stochastic", 0, S, 1, k, pc, cutoff, pcv=pcv49, state=ostart, rng=rng)
    assert [list(m.Positions) for m in got_s] == [p for _, p in want_s]
    got_g = MotifSampler.findBestMotifPositionsWithStartPositionByPCV(1, k, pc, cutoff, DNA, seqs, pcv, start)
    want_g, _ = O.motif_step("greedy", 0, S, 1, k, pc, cutoff, pcv=pcv49, state=ostart)
    assert [list(m.Positions) for m in got_g] == [p for _, p in want_g]
    np.testing.assert_allclose([m.PWMS for m in got_g], [v for v, _ in want_g], rtol=RTOL)


def test_motif_restart_loop_and_scope():
    case = (5, 27, None, 6, 1e-4, 1.0, 1)
    n, L, Lmin, k, pc, cutoff, seed = case
    ps, seqs, bg = _setup(case)
    S = O.sources(seqs)
    pcv = ProbabilityCompositeVector.ofACGT(*bg)
    for reps in (0, 1, 3):
        got = MotifSampler.findBestInormationContentContainingMotifsWithPCV(reps, 1, k, pc, cutoff, DNA, seqs, pcv, seed=5, chain=9)
        rng, _ = O.make_rng(seed=5, chain=9)
        want, _ = O.best_motif_information_content(0, reps, S, 1, k, pc, cutoff, rng, pcv=O.pcv_from_acgt(bg))
        assert [list(m.Positions) for m in got] == [p for _, p in want]
        np.testing.assert_allclose([m.PWMS for m in got], [v for v, _ in want], rtol=RTOL)
    with pytest.raises(_abi.GibbsUnsupportedError):    # one and two sites per sequence are built (test_gpu_motif2.py), not three
        MotifSampler.doMotifSamplingWithPCV(3, k, pc, cutoff, DNA, seqs, pcv)
    with pytest.raises(_abi.GibbsArgumentError):       # null PPM = ArgumentNullException; the family itself is built
        MotifSampler.doMotifSamplingWithPPM(1, k, pc, cutoff, DNA, seqs, None)
    with pytest.raises(_abi.GibbsUnsupportedError):
        MotifSampler.doMotifSamplingWithPPM(3, k, pc, cutoff, DNA, seqs, np.full((k, 4), 0.25))


@pytest.mark.parametrize("case", _cases(), ids=lambda c: f"n{c[0]}_L{c[1]}_k{c[3]}_cut{c[5]}")
def test_data_background_motif_restarts_match_oracle(case):
    """doMotifSampling (fs:1034): random starts with the drifting background, then sweeps whose background is
    rebuilt once per held-out sequence from the others' non-site bases (fs:896-905)."""
    n, L, Lmin, k, pc, cutoff, seed = case
    ps, seqs, bg = _setup(case)
    S = O.sources(seqs)
    n_chains = 4
    with GibbsEngine(seqs) as eng:
        params = make_params(k, pc, 5, [0.25] * 4, cutoff=cutoff, sampler=_abi.GIBBS_MOTIF_SAMPLER, background=_abi.GIBBS_BG_DATA)
        res = eng.run(params, n_chains, chain_id_base=70, seed=31 + seed, want_counts=False)
    for c in range(n_chains):
        rng, _ = O.make_rng(seed=31 + seed, chain=70 + c)
        want, st = O.motif_step("do_motif_sampling", 1, S, 1, k, pc, cutoff, rng=rng)
        _same_state(res.sites[c], res.scores[c], want)


def test_data_background_motif_functions(golden):
    """The script's second call shape (fsx:407) with motifAmount = 1, on its own multi-sample toy input (fsx:49-57)."""
    seqs, k, pc, cutoff = golden["multi_sequences"], 6, 1e-4, 1.0
    S = O.sources(seqs)
    for reps in (1, 3):
        got = MotifSampler.getMotifsWithBestInformationContents(reps, 1, k, pc, cutoff, DNA, seqs, seed=8, chain=2)
        rng, _ = O.make_rng(seed=8, chain=2)
        want, _ = O.best_motif_information_content(1, reps, S, 1, k, pc, cutoff, rng)
        assert [list(m.Positions) for m in got] == [p for _, p in want]
        np.testing.assert_allclose([m.PWMS for m in got], [v for v, _ in want], rtol=RTOL)
    n = len(seqs)
    start = [MotifSampler.createMotifIndex(1.5, [3]) for _ in range(n)]
    start[4] = MotifSampler.createMotifIndex(1e-9, [])          # the poly-T sequence has no site
    ostart = [(m.PWMS, list(m.Positions)) for m in start]
    picks = np.random.default_rng(6).random(n)
    full = np.concatenate([np.zeros(n * (n - 1)), picks])
    got_s = MotifSampler.findBestMotifIndicesByWithStartPositions(1, k, pc, cutoff, DNA, seqs, start, uniforms=full)
    rng, _ = O.make_rng(uniforms=picks)
    want_s, _ = O.motif_step("stochastic", 1, S, 1, k, pc, cutoff, state=ostart, rng=rng)
    assert [list(m.Positions) for m in got_s] == [p for _, p in want_s]
    np.testing.assert_allclose([m.PWMS for m in got_s], [v for v, _ in want_s], rtol=RTOL)
    got_g = MotifSampler.findBestMotifIndicesWithStartPositions(1, k, pc, cutoff, DNA, seqs, start)
    want_g, _ = O.motif_step("greedy", 1, S, 1, k, pc, cutoff, state=ostart)
    assert [list(m.Positions) for m in got_g] == [p for _, p in want_g]
    np.testing.assert_allclose([m.PWMS for m in got_g], [v for v, _ in want_g], rtol=RTOL)


@pytest.mark.parametrize("data", [False, True], ids=["fixed", "data"])
@pytest.mark.parametrize("shape", [(40, 150, None, 8, 0.0), (120, 400, 250, 12, 1.0), (30, 90, None, 20, -3.0), (25, 64, 40, 32, 0.0)],
                         ids=lambda s: f"n{s[0]}_L{s[1]}_k{s[3]}")
def test_greedy_ranking_pass_equals_the_all_windows_candidate_list(shape, data):
    """The greedy sweeps rank windows in fixed point and re-score the candidates; gibbs_set_option(GIBBS_OPT_EXACT_SCANS) forces the
    all-windows float64 candidate list. Both must give the same MotifIndex arrays, bit for bit, at sizes the oracle
    cannot reach in seconds (the small cases above and the fuzz test pin both against the oracle)."""
    n, L, Lmin, k, cutoff = shape
    ps = planted_motif_set(n, L, k, seed=88, min_length=Lmin)
    seqs = ps.sequences()
    bg = background_of(ps.ascii, 1e-4, 5)
    params = make_params(k, 1e-4, 5, bg, cutoff=cutoff, sampler=_abi.GIBBS_MOTIF_SAMPLER,
                         background=_abi.GIBBS_BG_DATA if data else _abi.GIBBS_BG_FIXED)
    with GibbsEngine(seqs) as eng:
        fast = eng.run(params, 6, chain_id_base=3, seed=9, want_counts=False)
        eng.set_option(_abi.GIBBS_OPT_EXACT_SCANS, 1)
        exact = eng.run(params, 6, chain_id_base=3, seed=9, want_counts=False)
    assert fast.sites.tolist() == exact.sites.tolist()
    assert fast.scores.tobytes() == exact.scores.tobytes()
    assert fast.sums.tobytes() == exact.sums.tobytes()


@pytest.mark.parametrize("data", [False, True], ids=["fixed", "data"])
@pytest.mark.parametrize("shape", [(20, 60, 40, 6, 1.0), (40, 150, None, 10, 0.0)], ids=lambda s: f"n{s[0]}_L{s[1]}_k{s[3]}")
def test_handover_stages_equal_one_warp_per_restart(shape, data):
    """More restarts than two per SM: the run starts with teams of 4 warps, the restarts still running when few are left pause at
    a sweep boundary and continue on teams of 8, then 16 warps (motif_stages, gibbs_api.cu). Every restart must end exactly
    where the one-warp kernel ends, and the device-side restart loop must return the same array."""
    n, L, Lmin, k, cutoff = shape
    ps = planted_motif_set(n, L, k, seed=21, min_length=Lmin)
    seqs = ps.sequences()
    bg = background_of(ps.ascii, 1e-4, 5)
    params = make_params(k, 1e-4, 5, bg, cutoff=cutoff, sampler=_abi.GIBBS_MOTIF_SAMPLER,
                         background=_abi.GIBBS_BG_DATA if data else _abi.GIBBS_BG_FIXED)
    n_chains = 700
    with GibbsEngine(seqs) as eng:
        staged = eng.run(params, n_chains, chain_id_base=5, seed=13, want_counts=False)
        best_staged = eng.fetch_best(n_chains - 1)
        eng.set_option(_abi.GIBBS_OPT_STAGE2_AT, 4)         # hand over earlier: more restarts travel through all three stages
        eng.set_option(_abi.GIBBS_OPT_STAGE3_AT, 3)
        early = eng.run(params, n_chains, chain_id_base=5, seed=13, want_counts=False)
        eng.set_team_warps(1)
        single = eng.run(params, n_chains, chain_id_base=5, seed=13, want_counts=False)
        best_single = eng.fetch_best(n_chains - 1)
    assert staged.stats["kernel_launches"] > single.stats["kernel_launches"]
    for got in (staged, early):
        assert got.sites.tolist() == single.sites.tolist()
        assert got.scores.tobytes() == single.scores.tobytes()
        assert got.sums.tobytes() == single.sums.tobytes()
        assert got.stats["site_updates"] == single.stats["site_updates"]
        assert got.stats["sweeps"] == single.stats["sweeps"]
    assert best_staged.restart == best_single.restart
    assert best_staged.sites.tolist() == best_single.sites.tolist()
