"""Host-side PositionMatrix helpers (fs:173-293) against the oracle's restatement of the same functions."""
import numpy as np
import pytest

import oracle_lib as O
from gibbssampling_b200 import PositionMatrix as PM
from gibbssampling_b200.BioArray import ofNucleotideString
from gibbssampling_b200.CompositeVector import ProbabilityCompositeVector

DNA = list("ATGC-")


def test_pfm_fuse_ppm_pwm_and_segment_score_match_the_oracle():
    rng = np.random.default_rng(3)
    k, n = 7, 6
    seqs = ["".join(rng.choice(list("ACGT-N"), size=40, p=[.23, .23, .23, .23, .05, .03])).encode() for _ in range(n)]
    S = O.sources(seqs)
    sites = rng.integers(0, 40 - k + 1, size=n).astype(np.int32)
    want_pfm = O.loo_pfm(S, sites, 2, k)                                   # [49][k] over the other five sites
    got_pfm = PM.fusePositionFrequencyMatrices(k, [PM.createPFMOf(s[p:p + k]) for i, (s, p) in enumerate(zip(seqs, sites)) if i != 2])
    assert got_pfm.tolist() == np.asarray(want_pfm).reshape(49, k).tolist()
    want_ppm = O.ppm_of_pfm(want_pfm, n - 1, 1e-4, b"ATGC-")
    got_ppm = PM.getPositionProbabilityMatrix(n - 1, DNA, 1e-4, got_pfm)
    assert np.array_equal(got_ppm, np.asarray(want_ppm).reshape(49, k))
    bg = [0.3, 0.2, 0.2, 0.3]
    pcv = ProbabilityCompositeVector.ofACGT(*bg)
    pcv["-"] = 0.05
    pcv49 = O.pcv_from_acgt(bg)
    pcv49[ord("-") - 42] = 0.05
    pwm = PM.createPositionWeightMatrix(DNA, pcv, got_ppm)
    want_raw = O.window_scores_bpv(seqs[2], k, pcv49, want_ppm, b"ATGC-")
    got_raw = [PM.calculateSegmentScoreBy(pwm, seqs[2][w:w + k]) for w in range(40 - k + 1)]
    assert got_raw == want_raw.tolist()                                     # same IEEE operations, same order


def test_normalize_ppm_works_in_place_like_the_reference():
    pfm = PM.createPFMOf(b"ACGT")
    ppm = PM.createPPMOf(pfm)
    out = PM.normalizePPM(1, DNA, 0.5, ppm)
    assert out is ppm                                                       # fs:256 wraps the argument's array
    assert ppm[ord("A") - 42, 0] == (1 + 0.5) / (1 + 5 * 0.5) and ppm[ord("C") - 42, 0] == 0.5 / 3.5


def test_fuse_rejects_a_wider_matrix_and_pads_a_narrower_one():
    with pytest.raises(IndexError):
        PM.fusePositionFrequencyMatrices(3, [PM.createPFMOf(b"ACGT")])
    m = PM.fusePositionFrequencyMatrices(5, [PM.createPFMOf(b"ACG"), PM.createPFMOf(b"AAAAA")])
    assert m[ord("A") - 42].tolist() == [2, 1, 1, 1, 1]


def test_profile_from_gap_padded_consensus_like_the_script():
    """fsx:81-127 + fsx:505-508: aligned, gap-padded consensus sequences -> PPM with dnaBases (Gap is a member)."""
    consensus = [ofNucleotideString(s) for s in ("-----cGTCcaGAAgg", "gGGAagCTCtgGAAgg", "tGAAgcTACagGACt-")]
    k = len(consensus[0])
    ppm = PM.getPositionProbabilityMatrix(len(consensus), DNA, 1e-4, PM.fusePositionFrequencyMatrices(k, [PM.createPFMOf(s) for s in consensus]))
    assert ppm.shape == (49, k)
    col0 = {ch: ppm[ord(ch) - 42, 0] for ch in "ACGT-"}
    den = 3 + 5 * 1e-4
    assert col0["-"] == (1 + 1e-4) / den and col0["G"] == (1 + 1e-4) / den and col0["T"] == (1 + 1e-4) / den and col0["A"] == 1e-4 / den


def test_composite_vector_helpers_reproduce_the_drifting_background():
    """createFCVWithout / increaseInPlaceFCVOf / substractSegmentCountsFrom / createNormalizedPCVOfFCV composed as
    getBestPWMSs composes them (fs:470-474) give the oracle's drifting-background window scores."""
    from gibbssampling_b200 import CompositeVector as CV
    rng = np.random.default_rng(5)
    n, L, k, pc = 5, 30, 6, 1e-4
    seqs = ["".join(rng.choice(list("ACGT"), size=L)).encode() for _ in range(n)]
    S = O.sources(seqs)
    sites = rng.integers(0, L - k + 1, size=n).astype(np.int32)
    h = 1
    pfm = O.loo_pfm(S, sites, h, k)
    ppm49 = np.asarray(O.ppm_of_pfm(pfm, n - 1, pc, b"ATGC-")).reshape(49, k)
    fcv = CV.fuseFrequencyVectors(DNA, [CV.createFCVWithout(k, int(sites[i]), seqs[i]) for i in range(n) if i != h])
    best = (0.0, 0)
    for w in range(L - k + 1):                                            # fs:466-478
        seg = seqs[h][w:w + k]
        CV.increaseInPlaceFCVOf(seqs[h], fcv)
        tmp = CV.substractSegmentCountsFrom(seg, fcv)
        assert tmp.Array is fcv.Array                                      # the aliasing of fs:85
        pcv = CV.createNormalizedPCVOfFCV(DNA, pc, tmp)
        score = PM.calculateSegmentScoreBy(PM.createPositionWeightMatrix(DNA, pcv, ppm49), seg)
        if score > best[0]:
            best = (score, w)
    raws = []
    fcv2 = CV.fuseFrequencyVectors(DNA, [CV.createFCVWithout(k, int(sites[i]), seqs[i]) for i in range(n) if i != h])
    for w in range(L - k + 1):
        CV.increaseInPlaceFCVOf(seqs[h], fcv2)
        pcv = CV.createNormalizedPCVOfFCV(DNA, pc, CV.substractSegmentCountsFrom(seqs[h][w:w + k], fcv2))
        raws.append(PM.calculateSegmentScoreBy(PM.createPositionWeightMatrix(DNA, pcv, ppm49), seqs[h][w:w + k]))
    want_score, want_pos, want_raw, _ = O.best_pwms(seqs[h], k, pc, O.loo_fcv(S, sites, h, k), ppm49)
    assert raws == want_raw.tolist()                                        # bit for bit, window by window
    assert best[1] == want_pos and np.log(best[0]) / np.log(2.0) == pytest.approx(want_score, rel=1e-12)
    assert CV.calculateSegmentScoreBy(CV.ProbabilityCompositeVector.ofACGT(.1, .2, .3, .4), b"ACGTA") == ((((1.0 * .1) * .2) * .3) * .4) * .1
