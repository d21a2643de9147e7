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


def test_parser_properties():
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=200, deadline=None)
    @given(st.text(alphabet=st.characters(min_codepoint=0, max_codepoint=255), max_size=80))
    def check(s):
        out = ofNucleotideString(s)
        assert all(c in NUCLEOTIDE_SYMBOLS for c in out)                   # only nucleotide symbols survive
        assert ofNucleotideString(out) == out                              # idempotent
        assert ofNucleotideString(s.lower()) == ofNucleotideString(s.upper()) or any(ord(c) > 127 for c in s)
        assert len(out) <= len(s)
        keep = [c.upper() for c in s if c.upper().encode("latin-1", "ignore") and c.upper().encode("latin-1", "ignore") in [bytes([x]) for x in NUCLEOTIDE_SYMBOLS]]
        assert out.decode() == "".join(keep)

    check()
