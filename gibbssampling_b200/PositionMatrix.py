"""Host-side pieces of the reference's PositionMatrix module that sit at the boundary (fs:126-293)."""
from __future__ import annotations

from typing import Sequence


def getBestInformationContent(item: Sequence[Sequence[tuple]]) -> list:
    """fs:156-170: the entry with the largest sum of scores (strict >), starting from [||]."""
    best: list = []
    for cand in item:
        ic_item = 0.0
        for pwms, _ in cand:
            ic_item = pwms + ic_item
        ic_best = 0.0
        for pwms, _ in best:
            ic_best = pwms + ic_best
        if ic_item > ic_best:
            best = list(cand)
    return best


def getRandomNumberInSequence(segmentLength: int, sourceLength: int, u: float) -> int:
    """fs:143-146 with the draw made explicit: rnd.Next(0, L - k + 1) = floor(u * (L - k + 1))."""
    return int(u * float(sourceLength - segmentLength + 1))


# ---------------------------------------------------------------------------------------------------------
# The matrix helpers a caller uses to PREPARE a PositionProbabilityMatrix for the ...OfPPM / ...WithPPM family
# (fsx:505-512: createPFMOf per aligned consensus sequence |> fusePositionFrequencyMatrices |> PPM) and to inspect
# a result. Host side, numpy [49][k] like the reference's Array2D (row = symbol - 42, fs:175-179); the GPU path
# never needs them (its tables live in shared memory), gibbs_set_start_ppm takes the A,C,G,T rows.
# ---------------------------------------------------------------------------------------------------------
import numpy as np  # noqa: E402

from .CompositeVector import NSLOT, _index  # noqa: E402
from .engine import _as_bytes  # noqa: E402


def createPFMOf(source) -> np.ndarray:
    """fs:211-215: one-hot [49][len(source)] count matrix of a sequence."""
    s = _as_bytes(source)
    pfm = np.zeros((NSLOT, len(s)), dtype=np.int32)
    for pos, sym in enumerate(s):
        pfm[_index(sym), pos] += 1
    return pfm


def fusePositionFrequencyMatrices(motifLength: int, countMatrices: Sequence[np.ndarray]) -> np.ndarray:
    """fs:218-226: element-wise sum into a fresh [49][motifLength] matrix; a wider input is the reference's
    IndexOutOfRangeException, a narrower one only fills its own columns."""
    out = np.zeros((NSLOT, int(motifLength)), dtype=np.int32)
    for m in countMatrices:
        m = np.asarray(m)
        if m.shape[0] != NSLOT or m.shape[1] > out.shape[1]:
            raise IndexError("count matrix wider than motifLength (IndexOutOfRangeException, fs:224)")
        out[:, :m.shape[1]] += m
    return out


def createPPMOf(positionFrequencyMatrix: np.ndarray) -> np.ndarray:
    """fs:249-251: int -> float."""
    return np.asarray(positionFrequencyMatrix).astype(np.float64)


def normalizePPM(sourceCount: int, alphabet, pseudoCount: float, positionProbabilityMatrix: np.ndarray) -> np.ndarray:
    """fs:255-261: (value + pc) / (sourceCount + |alphabet| pc) for the alphabet rows only, IN PLACE like the
    reference (its new matrix wraps the argument's array, fs:256); other rows keep the raw count."""
    m = positionProbabilityMatrix
    total = float(sourceCount) + float(len(alphabet)) * pseudoCount
    for item in alphabet:
        r = _index(item)
        for position in range(m.shape[1]):
            m[r, position] = (m[r, position] + pseudoCount) / total
    return m


def getPositionProbabilityMatrix(sourceCount: int, alphabet, pseudoCount: float, positionFrequencyMatrix: np.ndarray) -> np.ndarray:
    """createPPMOf |> normalizePPM, the helper the script calls (fsx:508, fsx:542, fsx:1159)."""
    return normalizePPM(sourceCount, alphabet, pseudoCount, createPPMOf(positionFrequencyMatrix))


def createPositionWeightMatrix(alphabet, pcv, ppMatrix: np.ndarray) -> np.ndarray:
    """fs:282-287: PWM[s, j] = PPM[s, j] / pcv[s] for s in alphabet, 0 elsewhere (an odds ratio, not a logarithm)."""
    pwm = np.zeros((NSLOT, ppMatrix.shape[1]), dtype=np.float64)
    for item in alphabet:
        r = _index(item)
        for position in range(ppMatrix.shape[1]):
            pwm[r, position] = ppMatrix[r, position] / pcv[item]
    return pwm


def calculateSegmentScoreBy(pwMatrix: np.ndarray, bioItems) -> float:
    """fs:290-293: ((1. * pwm[b0, 0]) * pwm[b1, 1]) * ..., left to right."""
    value = 1.0
    for position, sym in enumerate(_as_bytes(bioItems)):
        value = value * float(pwMatrix[_index(sym), position])
    return value
