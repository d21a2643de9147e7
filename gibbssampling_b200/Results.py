"""What the reference's script does with sampler results; host side only.

`GibbsSampling.fsx:386-388` / `fsx:409-411` group restarts by their position vectors and sort the
groups by size; `fsx:400` prints the segment found in every sequence; the result tables
(`fsx:1171-1348`) list positions with their scores. The PWM of a result is `createPPMOf` +
`normalizePPM` (fs:249-261) and `createPositionWeightMatrix` (fs:282-287) over the site counts that
`gibbs_run` already returns (`counts_out`), so nothing here touches the device.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Sequence

import numpy as np

BASES = "ACGT"  # column order of counts_out (include/gibbs_b200.h)


def countByPositions(results: Sequence) -> list[tuple[tuple, int]]:
    """`results |> Array.countBy (Array.map snd) |> Array.sortByDescending snd` (fsx:386-388).

    results: one `(score, position)[]` (SiteSampler) or `MotifIndex[]` (MotifSampler, fsx:409-411)
    per restart. Groups keep first-occurrence order (Array.countBy) and the sort is stable
    (Array.sortByDescending), like FSharp.Core."""
    groups: dict[tuple, int] = {}
    for items in results:
        key = tuple(_position_of(it) for it in items)
        groups[key] = groups.get(key, 0) + 1
    return sorted(groups.items(), key=lambda kv: -kv[1])


def _position_of(item):
    if hasattr(item, "Positions"):
        return tuple(item.Positions)
    return int(item[1])


def getDefinedSegment(subsequenceLength: int, source, startPoint: int):
    """`source |> Array.skip startPoint |> Array.take subsequenceLength` (fs:149-153, used at fsx:400).
    Raises like Array.skip / Array.take when the segment leaves the sequence."""
    if startPoint < 0 or startPoint > len(source):
        raise ValueError("startPoint outside the sequence (Array.skip, fs:151)")
    if subsequenceLength < 0 or startPoint + subsequenceLength > len(source):
        raise ValueError("segment longer than the rest of the sequence (Array.take, fs:152)")
    return source[startPoint:startPoint + subsequenceLength]


def segmentsOf(sources: Sequence, sites: Sequence[int], motifLength: int) -> list:
    """The k-mer every sequence contributes to the motif (the loop of fsx:399-400)."""
    return [getDefinedSegment(motifLength, s, int(p)) for s, p in zip(sources, sites)]


@dataclass
class MotifProfile:
    counts: np.ndarray        # [k, 4] int64, A C G T
    ppm: np.ndarray           # [k, 4] (c + pc) / (n + |A| pc), fs:260
    pwm: np.ndarray           # [k, 4] ppm / background, fs:286 (odds ratio, not a logarithm)
    log2_odds: np.ndarray     # [k, 4]
    information: np.ndarray   # [k] sum_b ppm log2(ppm / background), bits per column
    consensus: str            # most frequent base per column, first of A C G T on ties

    @property
    def total_information(self) -> float:
        return float(self.information.sum())


def motifProfile(counts, sourceCount: int, pseudoCount: float, alphabetSize: int, background: Sequence[float]) -> MotifProfile:
    """PPM / PWM of the sites of one result.

    counts       [k][4] site counts in A C G T order (gibbs_run's counts_out over all sequences, or
                 gibbs_loo_counts for a leave-one-out matrix)
    sourceCount  the divisor `normalizePPM` gets (fs:255: N for all sites, N - 1 leave-one-out)
    background   P(A), P(C), P(G), P(T) of the ProbabilityCompositeVector (fs:90)"""
    c = np.asarray(counts, dtype=np.int64)
    if c.ndim != 2 or c.shape[1] != 4:
        raise ValueError("counts must be [k][4]")
    q = np.asarray(background, dtype=np.float64)
    if q.shape != (4,):
        raise ValueError("background must hold P(A), P(C), P(G), P(T)")
    den = float(sourceCount) + float(alphabetSize) * pseudoCount      # fs:258
    ppm = (c.astype(np.float64) + pseudoCount) / den                  # fs:260
    pwm = ppm / q                                                     # fs:286
    with np.errstate(divide="ignore", invalid="ignore"):
        lg = np.log(pwm) / math.log(2.0)                              # FSharpAux log2 = ln x / ln 2
        info = np.where(ppm > 0, ppm * lg, 0.0).sum(axis=1)
    consensus = "".join(BASES[int(np.argmax(row))] for row in c)
    return MotifProfile(c, ppm, pwm, lg, info, consensus)


def scoreTable(result) -> list[tuple[int, int, float]]:
    """(sequence index, position, score) rows, the layout of the pasted tables fsx:1171-1348."""
    rows = []
    for i, it in enumerate(result):
        if hasattr(it, "Positions"):
            for p in it.Positions:
                rows.append((i, int(p), float(it.PWMS)))
        else:
            rows.append((i, int(it[1]), float(it[0])))
    return rows
