"""
Character maps, tip encoding and bit-exact site-pattern compression
(reference: alignment/charmaps.py, alignment/alignment.py:26-66, tests/test_utils.py).
"""
import numpy as np
import pytest

from phylo_utils_b200.alignment import alphabets, charmaps
from phylo_utils_b200.alignment.alignment import (seq_to_partials, seq_to_codes, alignment_to_numpy, alignment_to_codes,
                                                  compress_codes, invariant_sites, SeqRecord, read_alignment)
from helpers import load, records, CASES, ASC_CASES
from oracle import oracle

# the expectations of /root/reference/tests/test_utils.py:18-204, one row per IUPAC code
IUPAC = {"A": [1, 0, 0, 0], "C": [0, 1, 0, 0], "G": [0, 0, 1, 0], "T": [0, 0, 0, 1], "U": [0, 0, 0, 1],
         "R": [1, 0, 1, 0], "Y": [0, 1, 0, 1], "M": [1, 1, 0, 0], "K": [0, 0, 1, 1], "W": [1, 0, 0, 1],
         "S": [0, 1, 1, 0], "B": [0, 1, 1, 1], "D": [1, 0, 1, 1], "H": [1, 1, 0, 1], "V": [1, 1, 1, 0],
         "N": [1, 1, 1, 1], "-": [1, 1, 1, 1]}


@pytest.mark.parametrize("char", sorted(IUPAC))
def test_dna_iupac_codes(char):
    for c in {char, char.lower()}:
        out = seq_to_partials(c, alphabets.DNA)
        assert out.shape == (1, 4) and out.dtype == np.double and out.flags.c_contiguous
        assert np.array_equal(out[0], IUPAC[char])


def test_charmaps_equal_reference_tables():
    g = load("charmaps")
    for key, alpha in (("dna", alphabets.DNA), ("protein", alphabets.PROTEIN), ("binary", alphabets.BINARY)):
        chars = str(g[key + "_chars"])
        assert np.array_equal(seq_to_partials(chars, alpha), g[key])
    assert set(charmaps.dna_charmap) == set(str(g["dna_chars"]))
    assert set(charmaps.protein_charmap) == set(str(g["protein_chars"]))
    assert set(charmaps.binary_charmap) == set(str(g["binary_chars"]))


def test_unknown_character_raises_keyerror():
    with pytest.raises(KeyError):
        seq_to_partials("ACGTZ", alphabets.DNA)


def test_codes_are_ranks_in_lexicographic_row_order():
    for book in (charmaps.dna_codebook, charmaps.protein_codebook, charmaps.binary_codebook):
        rows = [tuple(r) for r in book.lut]
        assert rows == sorted(rows) and len(set(rows)) == len(rows)
    assert charmaps.dna_codebook.n_codes == 15 and charmaps.protein_codebook.n_codes == 21


@pytest.mark.parametrize("name", sorted(list(CASES) + list(ASC_CASES)))
def test_compression_is_bit_exact_with_reference(name):
    g = load(name)
    aln, sw, ii, names = alignment_to_numpy(records(g), int(g["alphabet"]))
    assert np.array_equal(aln, g["patterns"])
    assert np.array_equal(sw, g["siteweights"]) and sw.dtype == np.int64
    assert np.array_equal(ii, g["inverse_index"]) and ii.dtype == np.int64
    assert [n for n in names] == [str(n) for n in g["names"]]


def test_compression_matches_np_unique_on_random_inputs():
    rng = np.random.default_rng(5)
    for ntax, nsite, ncodes in [(1, 50, 15), (3, 200, 15), (7, 500, 4), (40, 300, 21), (5, 1, 15), (300, 64, 3)]:
        book = charmaps.dna_codebook if ncodes <= 15 else charmaps.protein_codebook
        codes = rng.integers(0, ncodes, size=(ntax, nsite)).astype(np.uint8)
        codes[:, rng.integers(0, nsite, size=nsite // 2)] = codes[:, rng.integers(0, nsite, size=nsite // 2)]
        pat, w, inv = compress_codes(codes)
        rp, rw, rinv = oracle.reference_compress(book.lut[codes])
        assert np.array_equal(book.lut[pat], rp) and np.array_equal(w, rw) and np.array_equal(inv, rinv)
        assert np.array_equal(pat[:, inv], codes) and w.sum() == nsite


def test_compression_edge_cases():
    empty = np.zeros((4, 0), dtype=np.uint8)
    pat, w, inv = compress_codes(empty)
    assert pat.shape == (4, 0) and w.size == 0 and inv.size == 0
    same = np.full((3, 17), 2, dtype=np.uint8)
    pat, w, inv = compress_codes(same)
    assert pat.shape == (3, 1) and w.tolist() == [17] and not inv.any()
    codes, lut, w, inv, names = alignment_to_codes([SeqRecord("a", "ACGT"), SeqRecord("b", "ACGA")], alphabets.DNA,
                                                   compress=False)
    assert w.tolist() == [1, 1, 1, 1] and inv.tolist() == [0, 1, 2, 3]
    with pytest.raises(ValueError):
        alignment_to_codes([SeqRecord("a", "ACGT"), SeqRecord("b", "ACG")], alphabets.DNA)


def test_invariant_sites():
    aln, _, _, _ = alignment_to_numpy([SeqRecord("a", "AACN"), SeqRecord("b", "ACCA"), SeqRecord("c", "ARC-")],
                                      alphabets.DNA, compress=False)
    assert invariant_sites(aln) == [True, False, True, True]


def test_fasta_and_phylip_readers(tmp_path):
    fa = tmp_path / "x.fa"
    fa.write_text(">t1 desc\nACGT\nAC\n>t2\nAC-TNN\n")
    recs = read_alignment(str(fa), "fasta")
    assert [(r.name, r.seq) for r in recs] == [("t1", "ACGTAC"), ("t2", "AC-TNN")]
    ph = tmp_path / "x.phy"
    ph.write_text(" 2 6\nt1  ACG TAC\nt2  AC-TNN\n")
    recs = read_alignment(str(ph), "phylip")
    assert [(r.name, r.seq) for r in recs] == [("t1", "ACGTAC"), ("t2", "AC-TNN")]
    assert seq_to_codes("ACGT", alphabets.DNA).dtype == np.uint8
