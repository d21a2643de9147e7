"""MotifSampler -- the reference module's functions (fs:709-1038) behind the CUDA library.

`MotifIndex` = {PWMS: float; Positions: int list} (fs:712-716). Only motifAmount = 1 is in scope
(SURVEY.md section 2, row 8): combinations of m >= 2 windows are an exponential enumeration, not
data-parallel window scoring. Both the fixed-background (`...ByPCV` / `...WithPCV`) family and the
data-derived-background family (fs:885-1038) run on the GPU.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional

import numpy as np

from . import _abi
from .CompositeVector import ProbabilityCompositeVector
from .engine import GibbsEngine, make_params
from .SiteSampler import _bg_of, _engine_for


@dataclass(frozen=True)
class MotifIndex:
    PWMS: float
    Positions: tuple = field(default_factory=tuple)


def createMotifIndex(pwms: float, pos) -> MotifIndex:
    """fs:719-723."""
    return MotifIndex(float(pwms), tuple(int(p) for p in pos))


def _check_m(motifAmount: int) -> None:
    if int(motifAmount) != 1:
        raise _abi.GibbsUnsupportedError(
            _abi.GIBBS_ERR_UNSUPPORTED,
            f"motifAmount = {motifAmount}: only one site per sequence is built; combinations of m >= 2 windows "
            "(calculatePWMsForSegmentCombinations, fs:727-742) are an exponential enumeration, out of scope")


def _to_motif_array(scores: np.ndarray, sites: np.ndarray) -> list:
    return [MotifIndex(float(s), (int(p),) if p >= 0 else ()) for s, p in zip(scores, sites)]


def _split_motif_state(motifMem) -> tuple[np.ndarray, np.ndarray]:
    if motifMem is None:
        raise _abi.GibbsArgumentError(_abi.GIBBS_ERR_ARG, "motifMem is null (ArgumentNullException)")
    scores = np.array([float(m.PWMS) for m in motifMem], dtype=np.float64)
    sites = np.empty(len(motifMem), dtype=np.int32)
    for i, m in enumerate(motifMem):
        if len(m.Positions) > 1:
            _check_m(len(m.Positions))
        sites[i] = m.Positions[0] if m.Positions else -1
    return scores, sites


def _run(phase_mask: int, motifAmount, motifLength, pseudoCount, cutOff, alphabet, sources, pcv, *, start=None,
         n_chains: int = 1, seed: int = 0, chain: int = 0, uniforms=None, engine: Optional[GibbsEngine] = None,
         max_sweeps: int = 0, ppM=None, best_of: Optional[int] = None):
    """pcv given -> the PCV family (fixed background); pcv None -> the data-derived family (fs:885-1038)."""
    _check_m(motifAmount)
    if pcv is None:
        _bg_of(alphabet, ProbabilityCompositeVector.ofACGT(1, 1, 1, 1))   # only checks A,C,G,T are in the alphabet
        bg, background = [0.25] * 4, _abi.GIBBS_BG_DATA
    else:
        bg, background = _bg_of(alphabet, pcv), _abi.GIBBS_BG_FIXED
    eng, own = _engine_for(sources, engine)
    try:
        params = make_params(motifLength, pseudoCount, len(alphabet), bg, cutoff=cutOff, background=background,
                             sampler=_abi.GIBBS_MOTIF_SAMPLER, phase_mask=phase_mask, max_sweeps=max_sweeps)
        if start is not None:
            scores, sites = _split_motif_state(start)
            eng.set_start_state(np.tile(sites, (n_chains, 1)), np.tile(scores, (n_chains, 1)))
        u = None if uniforms is None else np.asarray(uniforms, dtype=np.float64).reshape(n_chains, -1)
        if ppM is not None:
            eng.set_start_ppm(ppM, motifLength)
        try:
            if best_of is None:
                return eng.run(params, n_chains, chain_id_base=chain, seed=seed, uniforms=u, want_counts=False)
            eng.run_device(params, n_chains, chain_id_base=chain, seed=seed, uniforms=u)
            return eng.fetch_best(best_of)   # the restart loop (fs:857-881) runs on the GPU; only the winner comes back
        finally:
            if ppM is not None:
                eng.set_start_ppm(None)
    finally:
        if own:
            eng.close()


def rouletteWheelSelectionOfSites(motifLength, pseudoCount, cutOff, alphabet, sources, pcv, positions, heldOut: int,
                                  pick: float, *, engine: Optional[GibbsEngine] = None) -> MotifIndex:
    """calculateNormalizedSegmentScores (motifAmount 1) |> rouletteWheelSelection pick (fs:759-784, fs:746-754)
    for sources.[heldOut], with the leave-one-out PWM of `positions` (-1 = no site)."""
    bg = _bg_of(alphabet, pcv)
    eng, own = _engine_for(sources, engine)
    try:
        params = make_params(motifLength, pseudoCount, len(alphabet), bg, cutoff=cutOff, sampler=_abi.GIBBS_MOTIF_SAMPLER)
        pwms, site = eng.pick_roulette(positions, heldOut, params, pick)
        return MotifIndex(pwms, (site,) if site >= 0 else ())
    finally:
        if own:
            eng.close()


def findBestMotifPositionsWithStartPositionsByPCV(motifAmount, motifLength, pseudoCount, cutOff, alphabet, sources, pcv,
                                                  motifMem, **kw) -> list:
    """fs:828-853: the synchronous stochastic sweep (one roulette pick per sequence)."""
    res = _run(_abi.PHASE_STOCHASTIC, motifAmount, motifLength, pseudoCount, cutOff, alphabet, sources, pcv,
               start=motifMem, **kw)
    return _to_motif_array(res.scores[0], res.sites[0])


def findBestMotifPositionsWithStartPositionByPCV(motifAmount, motifLength, pseudoCount, cutOff, alphabet, sources, pcv,
                                                 motifMem, **kw) -> list:
    """fs:788-822: greedy in-place sweeps until the positions stop changing."""
    res = _run(_abi.PHASE_MOTIF_GREEDY, motifAmount, motifLength, pseudoCount, cutOff, alphabet, sources, pcv,
               start=motifMem, **kw)
    return _to_motif_array(res.scores[0], res.sites[0])


def doMotifSamplingWithPCV(motifAmount, motifLength, pseudoCount, cutOff, alphabet, sources, pcv, **kw) -> list:
    """One restart of fs:876-879: random starts |> stochastic sweep |> greedy sweeps (the reference inlines
    this pipeline in findBestInormationContentContainingMotifsWithPCV; there is no separate `do` function)."""
    res = _run(0, motifAmount, motifLength, pseudoCount, cutOff, alphabet, sources, pcv, **kw)
    return _to_motif_array(res.scores[0], res.sites[0])


def replay_motif_restart_loop(numberOfRepetitions: int, scores: np.ndarray, sites: np.ndarray, sums: np.ndarray) -> list:
    """fs:857-881 over restarts that already ran (same loop as SiteSampler.replay_restart_loop, over MotifIndex[])."""
    sums = np.asarray(sums, dtype=np.float64).tolist()

    def total(i):
        return 0.0 if (i is None or i < 0) else sums[i]

    def same(a, b):
        if a is None or b is None:
            return a is None and b is None
        if a < 0 or b < 0:
            i = a if b < 0 else b
            if i < 0:
                return True
            return scores.shape[1] == 1 and scores[i][0] == 0.0 and sites[i][0] < 0   # [|{PWMS 0.; Positions []}|]
        if sums[a] != sums[b]:
            return False   # equal MotifIndex arrays have equal left-to-right sums
        return bool(np.array_equal(sites[a], sites[b]) and np.array_equal(scores[a], scores[b]))

    acc, best, r, n = None, -1, 0, 0
    while n <= numberOfRepetitions and not same(acc, best):
        if total(acc) > total(best):
            best = acc if acc is not None else best
            acc = None
        else:
            acc = r
            r += 1
        n += 1
    if best < 0:
        return [MotifIndex(0.0, ())]
    return _to_motif_array(scores[best], sites[best])


def findBestInormationContentContainingMotifsWithPCV(numberOfRepetitions, motifAmount, motifLength, pseudoCount, cutOff,
                                                     alphabet, sources, pcv, *, seed: int = 0, chain: int = 0,
                                                     uniforms=None, engine: Optional[GibbsEngine] = None,
                                                     max_sweeps: int = 0) -> list:
    """fs:856-881: restarts run as parallel chains, the promote-or-restart loop is replayed over their results."""
    n_restarts = max(int(numberOfRepetitions) + 1, 1)
    res = _run(0, motifAmount, motifLength, pseudoCount, cutOff, alphabet, sources, pcv, n_chains=n_restarts,
               seed=seed, chain=chain, uniforms=uniforms, engine=engine, max_sweeps=max_sweeps,
               best_of=int(numberOfRepetitions))
    return _to_motif_array(res.scores, res.sites)


# ---- data-derived background (fs:885-1038): one background per held-out sequence (fs:896-905) -------------
def findBestMotifIndicesByWithStartPositions(motifAmount, motifLength, pseudoCount, cutOff, alphabet, sources, motifMem,
                                             **kw) -> list:
    """fs:935-970: the synchronous stochastic sweep."""
    res = _run(_abi.PHASE_STOCHASTIC, motifAmount, motifLength, pseudoCount, cutOff, alphabet, sources, None,
               start=motifMem, **kw)
    return _to_motif_array(res.scores[0], res.sites[0])


def findBestMotifIndicesWithStartPositions(motifAmount, motifLength, pseudoCount, cutOff, alphabet, sources, motifMem,
                                           **kw) -> list:
    """fs:885-929: greedy in-place sweeps."""
    res = _run(_abi.PHASE_MOTIF_GREEDY, motifAmount, motifLength, pseudoCount, cutOff, alphabet, sources, None,
               start=motifMem, **kw)
    return _to_motif_array(res.scores[0], res.sites[0])


def doMotifSampling(motifAmount, motifLength, pseudoCount, cutOff, alphabet, sources, **kw) -> list:
    """fs:1034-1038."""
    res = _run(0, motifAmount, motifLength, pseudoCount, cutOff, alphabet, sources, None, **kw)
    return _to_motif_array(res.scores[0], res.sites[0])


def getMotifsWithBestInformationContents(numberOfRepetitions, motifAmount, motifLength, pseudoCount, cutOff, alphabet,
                                         sources, *, seed: int = 0, chain: int = 0, uniforms=None,
                                         engine: Optional[GibbsEngine] = None, max_sweeps: int = 0) -> list:
    """fs:973-998 -- the reference script's second live call (fsx:407; there with motifAmount = 2, out of scope)."""
    n_restarts = max(int(numberOfRepetitions) + 1, 1)
    res = _run(0, motifAmount, motifLength, pseudoCount, cutOff, alphabet, sources, None, n_chains=n_restarts,
               seed=seed, chain=chain, uniforms=uniforms, engine=engine, max_sweeps=max_sweeps,
               best_of=int(numberOfRepetitions))
    return _to_motif_array(res.scores, res.sites)


def doMotifSamplingWithPPM(motifAmount, motifLength, pseudoCount, cutOff, alphabet, sources, ppM, **kw) -> list:
    """fs:1028-1032: SiteSampler.getMotifsWithBestPWMSOfPPM as the start, then the data-derived sweeps."""
    if ppM is None:
        raise _abi.GibbsArgumentError(_abi.GIBBS_ERR_ARG, "ppM is null (ArgumentNullException)")
    res = _run(0, motifAmount, motifLength, pseudoCount, cutOff, alphabet, sources, None, ppM=ppM, **kw)
    return _to_motif_array(res.scores[0], res.sites[0])


def getBestPWMSsOfPPM(numberOfRepetitions, motifAmount, motifLength, pseudoCount, cutOff, alphabet, sources, ppM, *,
                      seed: int = 0, chain: int = 0, uniforms=None, engine: Optional[GibbsEngine] = None,
                      max_sweeps: int = 0) -> list:
    """fs:1001-1026: the restart loop over doMotifSamplingWithPPM-shaped restarts."""
    if ppM is None:
        raise _abi.GibbsArgumentError(_abi.GIBBS_ERR_ARG, "ppM is null (ArgumentNullException)")
    n_restarts = max(int(numberOfRepetitions) + 1, 1)
    res = _run(0, motifAmount, motifLength, pseudoCount, cutOff, alphabet, sources, None, n_chains=n_restarts,
               seed=seed, chain=chain, uniforms=uniforms, engine=engine, max_sweeps=max_sweeps, ppM=ppM,
               best_of=int(numberOfRepetitions))
    return _to_motif_array(res.scores, res.sites)
