"""The oracle against every known-answer value that exists for this path.

The reference has no tests and cannot run here, so Appendix B (derived from the reference's code by
an independent model during the survey) plus the published Random123 Philox vectors are the pins.
"""
import numpy as np
import pytest

import oracle_lib as O


def _src(golden):
    return O.sources(golden["sequences"])


def test_philox_known_answers(golden):
    for kat in golden["philox4x32_10"]:
        out = O.philox4x32_10(kat["ctr"], kat["key"])
        assert [f"{x:08x}" for x in out] == kat["out"]


def test_uniform_stream_layout():
    # draw d = word d%4 of block d/4, counter = (block, chain), key = seed
    seed, chain = 0x123456789ABCDEF0, 0x0FEDCBA987654321
    for d in (0, 1, 2, 3, 4, 7, 1 << 33):
        blk = d >> 2
        out = O.philox4x32_10([blk & 0xFFFFFFFF, blk >> 32, chain & 0xFFFFFFFF, chain >> 32],
                              [seed & 0xFFFFFFFF, seed >> 32])
        assert O.uniform_at(seed, chain, d) == out[d & 3] / 4294967296.0


def test_background(golden):
    pcv = O.pcv_of_sources(_src(golden), golden["pc"])
    for ch, v in golden["q"].items():
        assert pcv[ord(ch) - 42] == v


def test_kat1_counts_ppm_and_best(golden):
    S, k, pc = _src(golden), golden["k"], golden["pc"]
    kat = golden["kat1"]
    pcv = O.pcv_of_sources(S, pc)
    pfm = O.loo_pfm(S, kat["sites"], kat["heldout"], k)
    for ch, row in kat["counts"].items():
        assert pfm[ord(ch) - 42].tolist() == row
    assert pfm.sum() == 3 * k
    ppm = O.ppm_of_pfm(pfm, S.n - 1, pc)
    assert ppm[ord("C") - 42, 0] == kat["P_C0"]
    assert ppm[ord("A") - 42, 0] == kat["P_A0"]
    for h in range(S.n):
        ppm = O.ppm_of_pfm(O.loo_pfm(S, kat["sites"], h, k), S.n - 1, pc)
        score, pos = O.best_pwms_with_bpv(S.seq(h), k, pcv, ppm)
        assert pos == kat["best_fixed"][h][1]
        assert score == pytest.approx(kat["best_fixed"][h][0], rel=1e-13)
        raw = O.window_scores_bpv(S.seq(h), k, pcv, ppm)
        assert raw.max() == pytest.approx(kat["max_raw_approx"], rel=1e-5)
        s2, p2, _, _ = O.best_pwms(S.seq(h), k, pc, O.loo_fcv(S, kat["sites"], h, k), ppm)
        assert p2 == kat["best_drifting"][h][1]
        assert s2 == pytest.approx(kat["best_drifting"][h][0], rel=1e-13)


def test_kat2_non_converged(golden):
    S, k, pc = _src(golden), golden["k"], golden["pc"]
    kat = golden["kat2"]
    pcv = O.pcv_of_sources(S, pc)
    pfm = O.loo_pfm(S, kat["sites"], kat["heldout"], k)
    for ch, row in kat["counts"].items():
        assert pfm[ord(ch) - 42].tolist() == row
    ppm = O.ppm_of_pfm(pfm, S.n - 1, pc)
    seq = S.seq(kat["heldout"])
    score, pos = O.best_pwms_with_bpv(seq, k, pcv, ppm)
    assert (score, pos) == tuple(kat["best_fixed"])
    lg = np.log(O.window_scores_bpv(seq, k, pcv, ppm)) / np.log(2.0)
    np.testing.assert_allclose(lg, kat["log2_scores_fixed"], rtol=1e-13)
    f0 = O.loo_fcv(S, kat["sites"], kat["heldout"], k)
    assert {ch: int(f0[ord(ch) - 42]) for ch in "ACGT"} == kat["F0"]
    s2, p2, raw, _ = O.best_pwms(seq, k, pc, f0, ppm)
    assert (s2, p2) == tuple(kat["best_drifting"])
    lg2 = np.log(raw) / np.log(2.0)
    np.testing.assert_allclose(lg2[:3], kat["log2_scores_drifting_first3"], rtol=1e-13)
    assert lg2[-1] == pytest.approx(kat["log2_scores_drifting_last"], rel=1e-13)


def test_kat3_roulette(golden):
    S, k, pc = _src(golden), golden["k"], golden["pc"]
    kat = golden["kat3"]
    pcv = O.pcv_of_sources(S, pc)
    ppm = O.ppm_of_pfm(O.loo_pfm(S, golden["kat1"]["sites"], 0, k), S.n - 1, pc)
    cand = O.candidates(S.seq(0), k, 1, kat["cutoff"], pcv, ppm)
    assert len(cand) == kat["n_background"] + 1
    assert all(c[1] == [] for c in cand[:16])
    assert cand[0][0] == kat["first_background"]
    assert sum(c[0] for c in cand[:16]) == pytest.approx(kat["background_mass_approx"], rel=1e-5)
    assert cand[16] == (kat["candidate"][0], kat["candidate"][1])
    pw = [c[0] for c in cand]
    for u in kat["picks_item0"]:
        assert O.roulette(pw, u) == 0
    for u in kat["picks_item16"]:
        assert O.roulette(pw, u) == 16
    p1 = kat["pc1"]
    pcv1 = O.pcv_of_sources(S, p1["pc"])
    ppm1 = O.ppm_of_pfm(O.loo_pfm(S, golden["kat1"]["sites"], 0, k), S.n - 1, p1["pc"])
    cand1 = O.candidates(S.seq(0), k, 1, p1["cutoff"], pcv1, ppm1)
    tail = cand1[16:]
    assert [c[1] for c in tail] == [c[1] for c in p1["candidates"]]
    for got, want in zip(tail, p1["candidates"]):
        assert got[0] == pytest.approx(want[0], rel=1e-13)
    pw1 = [c[0] for c in cand1]
    assert cand1[O.roulette(pw1, 0.9)][1] == [p1["u_0.9"]]
    assert cand1[O.roulette(pw1, 0.999999)][1] == [p1["u_0.999999"]]


def test_kat4_draw_to_position(golden):
    kat = golden["kat4"]
    for u, want in kat["draws"]:
        assert O.draw_to_position(u, kat["len"], kat["k"]) == want


def test_planted_truth(golden):
    for s, p in zip(golden["sequences"], golden["planted"]):
        assert s.find("CACGTG") == p
    for s, ps in zip(golden["multi_sequences"], golden["planted_multi"]):
        found, start = [], 0
        while (i := s.find("CACGTG", start)) >= 0:
            found.append(i)
            start = i + 1
        assert found == ps


def test_pipeline_finds_planted_sites(golden):
    """doSiteSamplingWithBPV on the script's toy input (fsx:29-35, call shape of fsx:384) converges to the
    planted CACGTG sites for a good share of streams; the last phase (right shifts) must be idempotent."""
    S, k, pc = _src(golden), golden["k"], golden["pc"]
    pcv = O.pcv_of_sources(S, pc)
    hits = 0
    for chain in range(16):
        rng, _ = O.make_rng(seed=7, chain=chain)
        score, pos, st = O.site_step("do_site_sampling_with_bpv", S, k, pc, pcv=pcv, rng=rng)
        assert rng.next == S.n * (S.n - 1)
        assert st.restarts == 1 and st.site_updates >= 4 * S.n
        s2, p2, _ = O.site_step("right_shifted_with_bpv", S, k, pc, pcv=pcv, state=(score, pos))
        assert p2.tolist() == pos.tolist() and s2.tolist() == score.tolist()
        hits += pos.tolist() == golden["planted"]
    assert hits >= 4
