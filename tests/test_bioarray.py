"""BioArray.ofNucleotideString semantics (charToParsedNucleotideChar + Seq.choose)."""
import pytest

from gibbssampling_b200.BioArray import NUCLEOTIDE_SYMBOLS, ofNucleotideString, ofNucleotideStrings, symbolIndex


def test_upper_cases_and_drops_everything_that_is_not_a_nucleotide_symbol():
    assert ofNucleotideString("acgt") == b"ACGT"
    assert ofNucleotideString("AC GT\nAC\tGT\r\n") == b"ACGTACGT"            # the script's multi-line literals, fsx:225-229
    assert ofNucleotideString("acgu-*nrykmswbdhvi") == b"ACGU-*NRYKMSWBDHVI"
    assert ofNucleotideString("ACGT123xzjoq.,;ACGT") == b"ACGTACGT"          # not nucleotide symbols: dropped, not an error
    assert ofNucleotideString(b"acgtn") == b"ACGTN"
    assert ofNucleotideString("") == b""
    with pytest.raises(ValueError):
        ofNucleotideString(None)
    assert ofNucleotideStrings(["ac", "g t"]) == [b"AC", b"GT"]


def test_every_kept_symbol_indexes_the_49_slot_tables():
    idx = {chr(c): symbolIndex(c) for c in NUCLEOTIDE_SYMBOLS}
    assert idx["*"] == 0 and idx["-"] == 3 and idx["A"] == 23 and idx["C"] == 25 and idx["G"] == 29 and idx["T"] == 42
    assert max(idx.values()) == 47 and len(set(idx.values())) == 19
    with pytest.raises(IndexError):
        symbolIndex(ord("a"))
    with pytest.raises(IndexError):
        symbolIndex(ord(" "))
