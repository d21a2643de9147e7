"""The ...OfPPM / ...WithPPM family (fs:644-689, fs:703, fs:1001-1032): the random starts are scored against a
caller-supplied PositionProbabilityMatrix; the random sites only shape the drifting background. GPU vs oracle."""
import numpy as np
import pytest

import oracle_lib as O
from gibbssampling_b200 import MotifSampler, SiteSampler, _abi
from gibbssampling_b200.engine import GibbsEngine, draws_per_chain, make_params
from gibbssampling_b200.synthetic import planted_motif_set

pytestmark = pytest.mark.gpu
RTOL = 1e-5
DNA = list("ATGC-")


def _ppm49(rng, k, zero_some=False):
    """A 49 x k PositionProbabilityMatrix like createPPMOf + normalizePPM would leave it: alphabet rows hold
    probabilities, every other row 0."""
    m = np.zeros((49, k))
    acgt = rng.dirichlet([0.6] * 4, size=k)            # [k][4], peaked columns
    if zero_some:
        acgt[0, 1] = 0.0
    for b, ch in enumerate("ACGT"):
        m[ord(ch) - 42, :] = acgt[:, b]
    m[ord("-") - 42, :] = 1e-6
    return m, acgt


CASES = [(5, 40, None, 6, 1), (9, 90, 50, 12, 2), (6, 200, None, 16, 3), (4, 30, 12, 3, 4)]


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"n{c[0]}_L{c[1]}_k{c[3]}")
def test_site_sampler_ppm_family_matches_oracle(case):
    n, L, Lmin, k, seed = case
    ps = planted_motif_set(n, L, k, seed=500 + seed, min_length=Lmin)
    seqs = ps.sequences()
    S = O.sources(seqs)
    rng = np.random.default_rng(seed)
    ppm49, acgt = _ppm49(rng, k, zero_some=(seed == 3))
    u = rng.random(draws_per_chain(n))
    # fs:644-661 alone
    r, keep = O.make_rng(uniforms=u)
    score, pos, _ = O.site_step("motifs_with_best_pwms_of_ppm", S, k, 1e-4, ppm=ppm49, rng=r)
    got = SiteSampler.getMotifsWithBestPWMSOfPPM(k, 1e-4, DNA, seqs, ppm49, uniforms=u)
    assert [p for _, p in got] == pos.tolist()
    np.testing.assert_allclose([s for s, _ in got], score, rtol=RTOL)
    # fs:703-707, [k][4] form of the same matrix
    r, keep = O.make_rng(uniforms=u)
    score, pos, _ = O.site_step("do_site_sampling_with_ppm", S, k, 1e-4, ppm=ppm49, rng=r)
    got = SiteSampler.doSiteSamplingWithPPM(k, 1e-4, DNA, seqs, acgt, uniforms=u)
    assert [p for _, p in got] == pos.tolist()
    np.testing.assert_allclose([s for s, _ in got], score, rtol=RTOL)
    # fs:664-689 with the Philox stream: restarts 0..reps as chains
    reps = 3
    r, keep = O.make_rng(seed=11, chain=0)
    ws, wp, _ = O.best_information_content(2, reps, S, k, 1e-4, r, ppm=ppm49)
    got = SiteSampler.getBestInformationContentOfPPM(reps, k, 1e-4, DNA, seqs, ppm49, seed=11)
    assert [p for _, p in got] == wp.tolist()
    np.testing.assert_allclose([s for s, _ in got], ws, rtol=RTOL)


def test_ppm_is_cleared_after_the_call_and_rejected_with_a_fixed_background():
    ps = planted_motif_set(6, 50, 6, seed=8)
    seqs = ps.sequences()
    rng = np.random.default_rng(1)
    ppm49, acgt = _ppm49(rng, 6)
    with GibbsEngine(seqs) as eng:
        plain = eng.run(make_params(6, 1e-4, 5, [0.25] * 4, background=_abi.GIBBS_BG_DATA), 2, seed=3)
        a = SiteSampler.doSiteSamplingWithPPM(6, 1e-4, DNA, seqs, ppm49, seed=3, engine=eng)
        again = eng.run(make_params(6, 1e-4, 5, [0.25] * 4, background=_abi.GIBBS_BG_DATA), 2, seed=3)
        assert again.sites.tolist() == plain.sites.tolist()          # the PPM did not stick to the handle
        eng.set_start_ppm(acgt)
        with pytest.raises(_abi.GibbsArgumentError):
            eng.run(make_params(6, 1e-4, 5, [0.25] * 4), 1, seed=3)   # fixed background: no such reference function
        with pytest.raises(_abi.GibbsArgumentError):
            eng.run(make_params(7, 1e-4, 5, [0.25] * 4, background=_abi.GIBBS_BG_DATA), 1, seed=3)   # k mismatch
        eng.set_start_ppm(None)
        with pytest.raises(_abi.GibbsArgumentError):
            eng.set_start_ppm(np.zeros((6, 3)))
    assert len(a) == 6


@pytest.mark.parametrize("case", CASES[:3], ids=lambda c: f"n{c[0]}_L{c[1]}_k{c[3]}")
def test_motif_sampler_ppm_family_matches_oracle(case):
    n, L, Lmin, k, seed = case
    ps = planted_motif_set(n, L, k, seed=600 + seed, min_length=Lmin)
    seqs = ps.sequences()
    S = O.sources(seqs)
    rng = np.random.default_rng(40 + seed)
    ppm49, acgt = _ppm49(rng, k)
    u = rng.random(draws_per_chain(n, _abi.GIBBS_MOTIF_SAMPLER))
    r, keep = O.make_rng(uniforms=u)
    want, _ = O.motif_step("do_motif_sampling", 2, S, 1, k, 1e-4, 0.0, ppm=ppm49, rng=r)
    got = MotifSampler.doMotifSamplingWithPPM(1, k, 1e-4, 0.0, DNA, seqs, ppm49, uniforms=u)
    assert [tuple(g.Positions) for g in got] == [tuple(p) for _, p in want]
    np.testing.assert_allclose([g.PWMS for g in got], [s for s, _ in want], rtol=RTOL)


def test_script_flow_profile_from_gap_padded_consensus_then_site_sampling_with_ppm():
    """fsx:505-512: a PPM built from aligned, gap-padded consensus sequences (dnaBases, Gap is a member) drives
    doSiteSamplingWithPPM over gap-free genes. Only the A,C,G,T rows of the PPM can matter there."""
    from gibbssampling_b200 import PositionMatrix as PM
    from gibbssampling_b200.BioArray import ofNucleotideString
    consensus = [ofNucleotideString(s) for s in ("-----cGTCcaGAA", "gGGAagCTCtgGAA", "tGAAgcTACagGAC", "cGAAggGCCgcGAC")]
    k = len(consensus[0])
    ppm49 = PM.getPositionProbabilityMatrix(len(consensus), DNA, 1e-4,
                                           PM.fusePositionFrequencyMatrices(k, [PM.createPFMOf(s) for s in consensus]))
    ps = planted_motif_set(7, 120, k, seed=77)
    genes = ps.sequences()
    S = O.sources(genes)
    u = np.random.default_rng(8).random(draws_per_chain(7))
    r, keep = O.make_rng(uniforms=u)
    score, pos, _ = O.site_step("do_site_sampling_with_ppm", S, k, 1e-4, ppm=ppm49, rng=r)
    got = SiteSampler.doSiteSamplingWithPPM(k, 1e-4, DNA, genes, ppm49, uniforms=u)
    assert [p for _, p in got] == pos.tolist()
    np.testing.assert_allclose([s for s, _ in got], score, rtol=RTOL)
