"""GPU parity of the MotifSampler with motifAmount = 2 (fs:727-742 driven by fs:778-782): up to two non-overlapping
sites per sequence; candidate list = background entries ++ single windows ++ pairs in the reference's order, roulette
pick in its sequential float64 order, greedy head of the stable sort. Checked against the oracle's `combos` for both
backgrounds, phase by phase and as whole restarts, including the reference script's own call (fsx:407)."""
import numpy as np
import pytest

import oracle_lib as O
from gibbssampling_b200 import MotifSampler, _abi
from gibbssampling_b200.CompositeVector import ProbabilityCompositeVector
from gibbssampling_b200.engine import GibbsEngine, make_params
from gibbssampling_b200.synthetic import background_of, planted_motif_set

pytestmark = pytest.mark.gpu
RTOL = 1e-5
DNA = list("ATGC-")
# bioTestsWithMultipleSamples, fsx:49-57: sequence 0 holds CACGTG twice (10, 21), sequence 4 is poly-T
SCRIPT_SET = ["GTGGCTGCACCACGTGTATGCCACGTG", "ACATCGCATCACGTGACCAGTTAGTTG", "CCTCGCACGTGGTGGTACAGTCGTACG",
              "GCATAAAGGACCATCACGTGAAGCTGC", "TTTTTTTTTTTTTTTTTTTTTTTTTTT"]


def _as_lists(got):
    return [list(m.Positions) for m in got], [m.PWMS for m in got]


def _check(got, want):
    pos, pw = _as_lists(got)
    assert pos == [list(p) for _, p in want]
    np.testing.assert_allclose(pw, [v for v, _ in want], rtol=RTOL)


def _cases():
    # (n, L, Lmin, k, pc, cutoff, seed)
    return [
        (5, 27, None, 6, 1e-4, 1.0, 1),
        (7, 70, 45, 6, 1e-2, 0.5, 2),
        (6, 90, None, 8, 1.0, 0.0, 3),        # large pseudocount, low cut-off: many singles and pairs
        (9, 110, 80, 9, 1e-3, 3.0, 4),
        (4, 64, None, 12, 1e-4, -2.0, 5),     # negative cut-off: pairs whose second window alone is below it
    ]


@pytest.mark.parametrize("data", [False, True], ids=["fixed", "data"])
@pytest.mark.parametrize("case", _cases(), ids=lambda c: f"n{c[0]}_L{c[1]}_k{c[3]}_cut{c[5]}")
def test_whole_restarts_match_oracle(case, data):
    n, L, Lmin, k, pc, cutoff, seed = case
    ps = planted_motif_set(n, L, k, seed=500 + seed, min_length=Lmin)
    seqs = ps.sequences()
    bg = background_of(ps.ascii, pc, 5)
    S = O.sources(seqs)
    pcv = O.pcv_from_acgt(bg)
    params = make_params(k, pc, 5, bg, cutoff=cutoff, sampler=_abi.GIBBS_MOTIF_SAMPLER, motif_amount=2,
                         background=_abi.GIBBS_BG_DATA if data else _abi.GIBBS_BG_FIXED)
    n_chains = 4
    with GibbsEngine(seqs) as eng:
        res = eng.run(params, n_chains, chain_id_base=9, seed=31 + seed, want_counts=False)
        pos = eng.fetch_positions(2)
    for c in range(n_chains):
        rng, _ = O.make_rng(seed=31 + seed, chain=9 + c)
        want, st = O.motif_step("do_motif_sampling", 1 if data else 0, S, 2, k, pc, cutoff, pcv=None if data else pcv, rng=rng)
        got = MotifSampler._to_motif_array(res.scores[c], pos[c])
        _check(got, want)
        assert res.sites[c].tolist() == [p[0] if p else -1 for _, p in want]   # sites = the newest position of each list
    assert any(len(p) == 2 for c in range(n_chains) for p in [list(x[x >= 0]) for x in pos[c]]) or cutoff > 2.5


@pytest.mark.parametrize("data", [False, True], ids=["fixed", "data"])
def test_phases_with_two_site_start_states_and_injected_uniforms(data):
    n, L, k, pc, cutoff = 8, 80, 6, 1e-2, 0.5
    ps = planted_motif_set(n, L, k, seed=77)
    seqs = ps.sequences()
    bg = background_of(ps.ascii, pc, 5)
    S = O.sources(seqs)
    pcv = O.pcv_from_acgt(bg)
    pv = None if data else ProbabilityCompositeVector.ofACGT(*bg)
    rng = np.random.default_rng(5)
    start = []
    for i in range(n):
        r = rng.integers(0, 3)
        if r == 0:
            start.append(MotifSampler.MotifIndex(float(rng.random() * 1e-4), ()))
        elif r == 1:
            start.append(MotifSampler.MotifIndex(float(rng.normal() * 2), (int(rng.integers(0, L - k + 1)),)))
        else:
            a = int(rng.integers(0, L - 3 * k))
            start.append(MotifSampler.MotifIndex(float(rng.normal() * 2 + 3), (a + 2 * k, a)))     # newest first, more than k apart
    state = [(m.PWMS, list(m.Positions)) for m in start]
    picks = rng.random(n)
    full = np.concatenate([np.zeros(n * (n - 1)), picks])
    variant = 1 if data else 0
    if data:
        got = MotifSampler.findBestMotifIndicesByWithStartPositions(2, k, pc, cutoff, DNA, seqs, start, uniforms=full)
    else:
        got = MotifSampler.findBestMotifPositionsWithStartPositionsByPCV(2, k, pc, cutoff, DNA, seqs, pv, start, uniforms=full)
    r, _ = O.make_rng(uniforms=picks)
    want, _ = O.motif_step("stochastic", variant, S, 2, k, pc, cutoff, pcv=None if data else pcv, state=state, rng=r)
    _check(got, want)
    if data:
        got = MotifSampler.findBestMotifIndicesWithStartPositions(2, k, pc, cutoff, DNA, seqs, start)
    else:
        got = MotifSampler.findBestMotifPositionsWithStartPositionByPCV(2, k, pc, cutoff, DNA, seqs, pv, start)
    want, _ = O.motif_step("greedy", variant, S, 2, k, pc, cutoff, pcv=None if data else pcv, state=state)
    _check(got, want)


def test_the_script_call_with_two_sites():
    """getMotifsWithBestInformationContents 1 2 6 0.0001 1. dnaBases bioTestsWithMultipleSamples (fsx:407-411)."""
    seqs = SCRIPT_SET
    S = O.sources(seqs)
    found_pair = False
    for seed in range(6):
        rng, _ = O.make_rng(seed=seed, chain=0)
        want, _ = O.best_motif_information_content(1, 1, S, 2, 6, 1e-4, 1.0, rng)
        got = MotifSampler.getMotifsWithBestInformationContents(1, 2, 6, 1e-4, 1.0, DNA, seqs, seed=seed)
        _check(got, want)
        found_pair |= any(len(m.Positions) == 2 for m in got)
        rng, _ = O.make_rng(seed=seed, chain=3)
        want, _ = O.motif_step("do_motif_sampling", 1, S, 2, 6, 1e-4, 1.0, rng=rng)
        got = MotifSampler.doMotifSampling(2, 6, 1e-4, 1.0, DNA, seqs, seed=seed, chain=3)
        _check(got, want)
    assert found_pair           # sequence 0 holds the planted 6-mer twice


def test_restart_loop_and_limits():
    ps = planted_motif_set(6, 60, 6, seed=9)
    seqs = ps.sequences()
    S = O.sources(seqs)
    bg = background_of(ps.ascii, 1e-3, 5)
    pcv = O.pcv_from_acgt(bg)
    pv = ProbabilityCompositeVector.ofACGT(*bg)
    for reps in (1, 3, 6):
        rng, _ = O.make_rng(seed=4, chain=20)
        want, _ = O.best_motif_information_content(0, reps, S, 2, 6, 1e-3, 0.5, rng, pcv=pcv)
        got = MotifSampler.findBestInormationContentContainingMotifsWithPCV(reps, 2, 6, 1e-3, 0.5, DNA, seqs, pv, seed=4, chain=20)
        _check(got, want)
    with pytest.raises(_abi.GibbsUnsupportedError):
        MotifSampler.doMotifSamplingWithPCV(3, 6, 1e-3, 0.5, DNA, seqs, pv, seed=1)
