"""GPU parity of the data-derived-background SiteSampler (doSiteSampling fs:697, getBestPWMSs fs:462): the
background counts drift from window to window (quirk A.6-1). This is the family the reference script calls
(fsx:384: getMotifsWithBestInformationContent 1 6 0.0001 dnaBases bioTests)."""
import numpy as np
import pytest

import oracle_lib as O
from gibbssampling_b200 import SiteSampler, _abi
from gibbssampling_b200.engine import GibbsEngine, draws_per_chain, make_params
from gibbssampling_b200.synthetic import planted_motif_set

pytestmark = pytest.mark.gpu
RTOL = 1e-5
DNA = list("ATGC-")


def _cases():
    # (n, L, Lmin, k, alen, pc, seed)
    return [
        (4, 21, None, 6, 5, 1e-4, 1),
        (9, 70, 40, 7, 5, 1e-4, 2),
        (12, 130, None, 12, 4, 1e-2, 3),
        (6, 300, 150, 16, 5, 1e-4, 4),
        (5, 90, None, 31, 5, 1.0, 5),
        (7, 33, 9, 1, 5, 1e-4, 6),
    ]


@pytest.mark.parametrize("case", _cases(), ids=lambda c: f"n{c[0]}_L{c[1]}_k{c[3]}")
def test_data_background_restarts_match_oracle(case):
    n, L, Lmin, k, alen, pc, seed = case
    ps = planted_motif_set(n, L, k, seed=300 + seed, min_length=Lmin)
    seqs = ps.sequences()
    S = O.sources(seqs)
    alphabet = b"ATGC-"[:alen]
    params = make_params(k, pc, alen, [0.25] * 4, background=_abi.GIBBS_BG_DATA)
    with GibbsEngine(seqs) as eng:
        res = eng.run(params, 5, chain_id_base=10, seed=seed, want_counts=False)
    total = 0
    for c in range(5):
        rng, _ = O.make_rng(seed=seed, chain=10 + c)
        score, pos, st = O.site_step("do_site_sampling", S, k, pc, rng=rng, alphabet=alphabet)
        assert res.sites[c].tolist() == pos.tolist(), f"chain {c}"
        np.testing.assert_allclose(res.scores[c], score, rtol=RTOL)
        total += st.site_updates
    assert res.stats["site_updates"] == total and res.stats["fast_path"] == (1 if pc > 0 else 0)


def test_script_call_and_phase_functions(golden):
    """The reference script's live call shape on its own toy input (fsx:29-35, fsx:384)."""
    seqs, k, pc = golden["sequences"], golden["k"], golden["pc"]
    S = O.sources(seqs)
    n = len(seqs)
    for reps in (1, 4):
        got = SiteSampler.getMotifsWithBestInformationContent(reps, k, pc, DNA, seqs, seed=2024, chain=3)
        rng, _ = O.make_rng(seed=2024, chain=3)
        ws, wp, _ = O.best_information_content(1, reps, S, k, pc, rng)
        assert [p for _, p in got] == wp.tolist()
        np.testing.assert_allclose([s for s, _ in got], ws, rtol=RTOL)
    # KAT-1 of SURVEY Appendix B: the converged state scores 11.81 / 11.98 / 11.85 / 11.84 with the drifting background
    u = np.random.default_rng(1).random(draws_per_chain(n))
    steps = [("random_starts", SiteSampler.getPWMOfRandomStarts), ("best_pwms_with_start_positions", SiteSampler.getBestPWMSsWithStartPositions),
             ("left_shifted", SiteSampler.getLeftShiftedBestPWMSs), ("right_shifted", SiteSampler.getRightShiftedBestPWMSs)]
    state = None
    for name, fn in steps:
        rng, _ = O.make_rng(uniforms=u)
        o_score, o_pos, _ = O.site_step(name, S, k, pc, state=state, rng=rng)
        got = fn(k, pc, DNA, seqs, uniforms=u) if state is None else fn(k, pc, DNA, seqs, [(float(a), int(b)) for a, b in zip(*state)])
        assert [p for _, p in got] == o_pos.tolist(), name
        np.testing.assert_allclose([s for s, _ in got], o_score, rtol=RTOL)
        state = (o_score, o_pos)
    kat = golden["kat1"]
    start = [(0.0, p) for p in kat["sites"]]
    got = SiteSampler.getBestPWMSsWithStartPositions(k, pc, DNA, seqs, start)
    assert [p for _, p in got] == kat["sites"]
    np.testing.assert_allclose([s for s, _ in got], [v for v, _ in kat["best_drifting"]], rtol=1e-12)


@pytest.mark.parametrize("shape", [(60, 200, None, 8), (200, 500, 300, 12), (40, 120, None, 20), (30, 64, 40, 32)],
                         ids=lambda s: f"n{s[0]}_L{s[1]}_k{s[3]}")
def test_ranking_pass_equals_the_all_windows_float64_scan(shape):
    """The float32 ranking pass + exact re-scoring of the candidates must return what the scan of every window in
    float64 returns (gibbs_set_option GIBBS_OPT_EXACT_SCANS forces the latter): same sites, bit-identical scores, at sizes the CPU
    oracle cannot reach in seconds. The small cases above pin both against the oracle."""
    n, L, Lmin, k = shape
    ps = planted_motif_set(n, L, k, seed=77, min_length=Lmin)
    seqs = ps.sequences()
    params = make_params(k, 1e-4, 5, [0.25] * 4, background=_abi.GIBBS_BG_DATA)
    with GibbsEngine(seqs) as eng:
        fast = eng.run(params, 6, chain_id_base=3, seed=9, want_counts=False)
        eng.set_option(_abi.GIBBS_OPT_EXACT_SCANS, 1)
        exact = eng.run(params, 6, chain_id_base=3, seed=9, want_counts=False)
    assert fast.stats["fast_path"] == 1 and exact.stats["fast_path"] == 0
    assert exact.stats["exact_rescans"] == exact.stats["site_updates"]
    assert fast.stats["exact_rescans"] < 0.05 * fast.stats["site_updates"]
    assert fast.sites.tolist() == exact.sites.tolist()
    assert fast.scores.tobytes() == exact.scores.tobytes()
    assert fast.sums.tobytes() == exact.sums.tobytes()


def test_zero_pseudocount_takes_the_exact_scan():
    ps = planted_motif_set(8, 60, 6, seed=5)
    seqs = ps.sequences()
    S = O.sources(seqs)
    params = make_params(6, 0.0, 5, [0.25] * 4, background=_abi.GIBBS_BG_DATA)
    with GibbsEngine(seqs) as eng:
        res = eng.run(params, 3, seed=4, want_counts=False)
    assert res.stats["fast_path"] == 0
    for c in range(3):
        rng, keep = O.make_rng(seed=4, chain=c)
        score, pos, _ = O.site_step("do_site_sampling", S, 6, 0.0, rng=rng)
        assert res.sites[c].tolist() == pos.tolist()


def test_hand_over_stages_do_not_change_a_chain():
    """More than two chains per SM: stage 1 (4 warps) pauses the last chains and stage 2 (8 warps) continues them.
    A chain's result depends on (seed, chain id) only, whatever the launch shape; spot-checked against the oracle."""
    n, L, k = 24, 100, 8
    ps = planted_motif_set(n, L, k, seed=41)
    seqs = ps.sequences()
    params = make_params(k, 1e-4, 5, [0.25] * 4, background=_abi.GIBBS_BG_DATA)
    with GibbsEngine(seqs) as eng:
        big = eng.run(params, 2000, seed=23, want_counts=False)
        assert big.stats["kernel_launches"] >= 3                     # two chain-kernel stages + best chain
        for chains, base in ((1, 0), (5, 1234), (300, 900), (64, 1936)):
            res = eng.run(params, chains, chain_id_base=base, seed=23, want_counts=False)
            assert res.sites.tobytes() == big.sites[base:base + chains].tobytes(), (chains, base)
            assert res.scores.tobytes() == big.scores[base:base + chains].tobytes(), (chains, base)
    S = O.sources(seqs)
    for c in (0, 777, 1999):
        rng, _ = O.make_rng(seed=23, chain=c)
        score, pos, _ = O.site_step("do_site_sampling", S, k, 1e-4, rng=rng)
        assert big.sites[c].tolist() == pos.tolist(), f"chain {c}"
        np.testing.assert_allclose(big.scores[c], score, rtol=RTOL)


@pytest.mark.parametrize("wide", ["0", "1"])
@pytest.mark.parametrize("shape", [(60, 64, 40, 9), (300, 90, None, 12)], ids=lambda s: f"n{s[0]}_L{s[1]}_k{s[3]}")
def test_random_starts_same_on_both_kernels(shape, wide):
    """getPWMOfRandomStarts (fs:589-611) inside the chain kernel or as the grid-wide init kernel: both consume the
    uniform stream as the oracle does."""
    n, L, Lmin, k = shape
    ps = planted_motif_set(n, L, k, seed=22, min_length=Lmin)
    seqs = ps.sequences()
    S = O.sources(seqs)
    with GibbsEngine(seqs) as eng:
        eng.set_option(_abi.GIBBS_OPT_INIT_PATH, _abi.GIBBS_INIT_WIDE if wide == "1" else _abi.GIBBS_INIT_CHAIN)
        init = eng.run(make_params(k, 1e-4, 5, [0.25] * 4, background=_abi.GIBBS_BG_DATA, phase_mask=_abi.PHASE_INIT), 3,
                       chain_id_base=5, seed=31, want_counts=False)
        full = eng.run(make_params(k, 1e-4, 5, [0.25] * 4, background=_abi.GIBBS_BG_DATA), 3, chain_id_base=5, seed=31,
                       want_counts=False)
    assert init.stats["kernel_launches"] >= (2 if wide == "1" else 1)
    for c in range(3):
        rng, _ = O.make_rng(seed=31, chain=5 + c)
        score, pos, _ = O.site_step("random_starts", S, k, 1e-4, rng=rng)
        assert init.sites[c].tolist() == pos.tolist(), f"chain {c}"
        np.testing.assert_allclose(init.scores[c], score, rtol=RTOL)
        rng, _ = O.make_rng(seed=31, chain=5 + c)
        score, pos, _ = O.site_step("do_site_sampling", S, k, 1e-4, rng=rng)
        assert full.sites[c].tolist() == pos.tolist(), f"chain {c}"
