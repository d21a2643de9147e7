"""Host-side twins of the reference's CompositeVector types (fs:11-124).

Only the *data types that cross the drop-in boundary* live here: a caller of the WithBPV family
hands in a ProbabilityCompositeVector (fs:90), exactly as in the reference. The helpers that build
one from sequences (createFCVOf -> fuseFrequencyVectors -> createNormalizedPCVOfFCV) are
the caller-side preparation the reference leaves to the user; they are not part of the GPU hot path.
"""
from __future__ import annotations

from typing import Iterable, Sequence

import numpy as np

from .engine import _as_bytes, symbol_code

NSLOT = 49  # Array.zeroCreate 49 (fs:20)


def _index(key) -> int:
    i = symbol_code(key) - 42  # (BioItem.symbol key |> int) - 42 (fs:17)
    if not (0 <= i < NSLOT):
        raise IndexError(f"symbol {key!r} is outside the 49-slot table (fs:17)")
    return i


class CompositeVector:
    """One dimensional array with fixed positions for each element (fs:14-30)."""

    dtype = np.float64

    def __init__(self, array=None):
        self.Array = np.zeros(NSLOT, dtype=self.dtype) if array is None else np.asarray(array, dtype=self.dtype)
        if self.Array.shape != (NSLOT,):
            raise ValueError("CompositeVector needs 49 slots")

    def __getitem__(self, key):
        return self.Array[_index(key)]

    def __setitem__(self, key, value):
        self.Array[_index(key)] = value


class FrequencyCompositeVector(CompositeVector):
    dtype = np.int64


class ProbabilityCompositeVector(CompositeVector):
    dtype = np.float64

    @classmethod
    def ofACGT(cls, a: float, c: float, g: float, t: float) -> "ProbabilityCompositeVector":
        v = cls()
        v["A"], v["C"], v["G"], v["T"] = a, c, g, t
        return v

    def acgt(self) -> list[float]:
        return [float(self[ch]) for ch in "ACGT"]


def createFCVOf(resSources) -> FrequencyCompositeVector:
    """fs:60-62: symbol counts of one BioArray."""
    v = FrequencyCompositeVector()
    codes = np.frombuffer(_as_bytes(resSources), dtype=np.uint8).astype(np.int64) - 42
    if codes.size and (codes.min() < 0 or codes.max() >= NSLOT):
        raise IndexError("symbol outside the 49-slot table (fs:17)")
    v.Array += np.bincount(codes, minlength=NSLOT)
    return v


def fuseFrequencyVectors(alphabet: Sequence, bfVectors: Iterable[FrequencyCompositeVector]) -> FrequencyCompositeVector:
    """fs:65-70: sums the alphabet slots only."""
    out = FrequencyCompositeVector()
    for fcv in bfVectors:
        for item in alphabet:
            out[item] = out[item] + fcv[item]
    return out


def createNormalizedPCVOfFCV(alphabet: Sequence, pseudoCount: float, fcv: FrequencyCompositeVector) -> ProbabilityCompositeVector:
    """fs:115-120: (count + pc) / (sum of ALL 49 counts + |alphabet| * pc) for the alphabet slots."""
    pcv = ProbabilityCompositeVector(fcv.Array.astype(np.float64))
    total = float(int(fcv.Array.sum())) + (float(len(alphabet)) * float(pseudoCount))
    for item in alphabet:
        pcv[item] = (pcv[item] + float(pseudoCount)) / total
    return pcv


def createPCVOfSources(alphabet: Sequence, pseudoCount: float, sources: Sequence) -> ProbabilityCompositeVector:
    """The fixed background a user of the WithBPV family would build."""
    fused = fuseFrequencyVectors(alphabet, (createFCVOf(s) for s in sources))
    return createNormalizedPCVOfFCV(alphabet, pseudoCount, fused)


# ---- the remaining functions of the reference module (fs:43-124), host side, for callers and for inspection ----
def increaseInPlaceFCV(bioItem, frequencyCompositeVector: FrequencyCompositeVector) -> FrequencyCompositeVector:
    """fs:43-46."""
    frequencyCompositeVector[bioItem] = frequencyCompositeVector[bioItem] + 1
    return frequencyCompositeVector


def createFCVWithout(motifLength: int, position: int, resSource) -> FrequencyCompositeVector:
    """fs:73-76: counts of a sequence outside the segment [position, position + motifLength)."""
    s = _as_bytes(resSource)
    return createFCVOf(s[:max(position, 0)] + s[position + motifLength:])


def increaseInPlaceFCVOf(resSources, backGroundCounts: FrequencyCompositeVector) -> FrequencyCompositeVector:
    """fs:79-81: adds every symbol of the sequence to the SAME vector (the in-place mutation behind quirk A.6-1)."""
    backGroundCounts.Array += createFCVOf(resSources).Array
    return backGroundCounts


def substractSegmentCountsFrom(source, fcVector: FrequencyCompositeVector) -> FrequencyCompositeVector:
    """fs:84-88: max(count - 1, 0) per symbol of the segment. The result WRAPS the argument's array (fs:85), so the
    argument is mutated too -- the aliasing that makes the background drift from window to window."""
    out = FrequencyCompositeVector.__new__(FrequencyCompositeVector)
    out.Array = fcVector.Array
    for sym in _as_bytes(source):
        c = int(fcVector[sym])
        out[sym] = c - 1 if c - 1 > 0 else 0
    return out


def createPCVOf(caArray: FrequencyCompositeVector) -> ProbabilityCompositeVector:
    """fs:109-112: int -> float copy."""
    return ProbabilityCompositeVector(caArray.Array.astype(np.float64))


def calculateSegmentScoreBy(pcv: ProbabilityCompositeVector, bioItems) -> float:
    """fs:123-124: ((1. * pcv[b0]) * pcv[b1]) * ..., the background-only probability of a window (fs:776)."""
    value = 1.0
    for sym in _as_bytes(bioItems):
        value = value * float(pcv[sym])
    return value
