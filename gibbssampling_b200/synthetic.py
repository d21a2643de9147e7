"""Synthetic planted-motif DNA: the workload of every BASELINE.json config.

Bases are i.i.d. uniform over {A,C,G,T} from a fixed-seed counter-based generator (numpy's Philox,
key 0xB200); one consensus k-mer from the same generator is planted once per sequence at a uniform
position in [0, L-k]; each planted base mutates to a uniform other base with p = 0.1.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

ALPHABET = np.frombuffer(b"ACGT", dtype=np.uint8)


@dataclass
class PlantedSet:
    ascii: np.ndarray        # uint8 [N*L] (fixed length) or ragged concatenation
    offsets: np.ndarray      # int64 [N+1]
    truth: np.ndarray        # int32 [N] planted start positions
    consensus: bytes
    k: int

    @property
    def n(self) -> int:
        return len(self.offsets) - 1

    def sequences(self) -> list[bytes]:
        return [self.ascii[self.offsets[i]: self.offsets[i + 1]].tobytes() for i in range(self.n)]


def planted_motif_set(n_seqs: int, length: int, k: int, *, seed: int = 0xB200, mutation: float = 0.1,
                      min_length: int | None = None) -> PlantedSet:
    """N sequences of `length` bp (or uniform in [min_length, length] when min_length is given)."""
    rng = np.random.Generator(np.random.Philox(key=seed))
    consensus = rng.integers(0, 4, size=k, dtype=np.int64)
    if min_length is None:
        lens = np.full(n_seqs, length, dtype=np.int64)
    else:
        lens = rng.integers(min_length, length + 1, size=n_seqs, dtype=np.int64)
    off = np.zeros(n_seqs + 1, dtype=np.int64)
    off[1:] = np.cumsum(lens)
    codes = rng.integers(0, 4, size=int(off[-1]), dtype=np.int64)
    pos = (rng.random(n_seqs) * (lens - k + 1)).astype(np.int64)
    mut = rng.random((n_seqs, k)) < mutation
    shift = rng.integers(1, 4, size=(n_seqs, k), dtype=np.int64)
    motif = np.where(mut, (consensus[None, :] + shift) % 4, consensus[None, :])
    idx = (off[:-1] + pos)[:, None] + np.arange(k)[None, :]
    codes[idx.reshape(-1)] = motif.reshape(-1)
    return PlantedSet(ALPHABET[codes], off, pos.astype(np.int32), ALPHABET[consensus].tobytes(), k)


def background_of(ascii_codes: np.ndarray, pseudocount: float, alphabet_size: int) -> list[float]:
    """pcv.[A,C,G,T] a user of the WithBPV family would pass: whole-set base counts normalised like
    createNormalizedPCVOfFCV (fs:115-120): (count + pc) / (total + |alphabet| * pc)."""
    counts = np.bincount(ascii_codes, minlength=256)
    total = float(int(ascii_codes.size)) + (float(alphabet_size) * float(pseudocount))
    return [(float(counts[ord(ch)]) + float(pseudocount)) / total for ch in "ACGT"]
