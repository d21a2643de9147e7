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
