"""The step before the hot path: text -> symbols.

The reference's script builds every input with `BioArray.ofNucleotideString` (fsx:47 ...), i.e.
BioFSharp's `charToParsedNucleotideChar`: the character is upper-cased, the 19 nucleotide symbols
`* - A B C D G H I K M N R S T U V W Y` are kept and anything else (whitespace, line breaks of the
multi-line literals fsx:225-229, digits, other letters) is silently dropped.

The result is the byte string `gibbs_create` takes (one `BioItem.symbol` per item, fs:17). On the device
A, C, G, T are 2-bit codes and every other kept symbol sets the mask plane: it is outside the alphabet,
so its PWM row is 0 (fs:283-287) and a window that holds it scores 0.
"""
from __future__ import annotations

from typing import Iterable

NUCLEOTIDE_SYMBOLS = b"*-ABCDGHIKMNRSTUVWY"
_KEEP = bytes(c for c in range(256) if bytes([c]).upper()[0:1] and bytes([c]).upper()[0] in NUCLEOTIDE_SYMBOLS)
_UPPER = bytes(bytes([c]).upper()[0] for c in range(256))
_DROP = bytes(c for c in range(256) if c not in _KEEP)


def ofNucleotideString(s) -> bytes:
    """`BioArray.ofNucleotideString` as symbol bytes."""
    if s is None:
        raise ValueError("string is null (ArgumentNullException)")
    b = s.encode("latin-1", errors="ignore") if isinstance(s, str) else bytes(s)
    return b.translate(_UPPER, _DROP)


def ofNucleotideStrings(strings: Iterable) -> list[bytes]:
    return [ofNucleotideString(s) for s in strings]


def symbolIndex(symbol: int) -> int:
    """`(int (BioItem.symbol a)) - 42` (fs:17, fs:176): row of the 49-slot tables."""
    i = int(symbol) - 42
    if not 0 <= i < 49:
        raise IndexError(f"symbol {symbol!r} outside '*'..'Z' (IndexOutOfRangeException, fs:17-20)")
    return i
