"""MotifSampler -- the reference module's functions (fs:709-1038) behind the CUDA library.

`MotifIndex` = {PWMS: float; Positions: int list} (fs:712-716). Only motifAmount = 1 is in scope
(SURVEY.md section 2, row 8): combinations of m >= 2 windows are an exponential enumeration, not
data-parallel window scoring.
"""
from __future__ import annotations

from dataclasses import dataclass, field

from . import _abi


@dataclass(frozen=True)
class MotifIndex:
    PWMS: float
    Positions: tuple = field(default_factory=tuple)


def createMotifIndex(pwms: float, pos) -> MotifIndex:
    """fs:719-723."""
    return MotifIndex(float(pwms), tuple(int(p) for p in pos))


def _not_built(name: str, where: str):
    raise _abi.GibbsUnsupportedError(_abi.GIBBS_ERR_UNSUPPORTED, f"{name} ({where}): MotifSampler kernels are not built yet")


def findBestInormationContentContainingMotifsWithPCV(numberOfRepetitions, motifAmount, motifLength, pseudoCount, cutOff,
                                                     alphabet, sources, pcv, **kw):
    """fs:856-881."""
    _not_built("findBestInormationContentContainingMotifsWithPCV", "fs:856")


def doMotifSampling(motifAmount, motifLength, pseudoCount, cutOff, alphabet, sources, **kw):
    """fs:1034-1038 -- data-derived background."""
    _not_built("doMotifSampling", "fs:1034")


def getMotifsWithBestInformationContents(numberOfRepetitions, motifAmount, motifLength, pseudoCount, cutOff, alphabet,
                                         sources, **kw):
    """fs:973-998 -- data-derived background."""
    _not_built("getMotifsWithBestInformationContents", "fs:973")
