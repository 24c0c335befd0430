"""
Character -> 0/1 state-set vectors.

Same content as the literal tables in /root/reference/phylo_utils/alignment/charmaps.py:2-86
(pinned character by character by tests/test_alignment.py against the values that
/root/reference/tests/test_utils.py:18-204 asserts), but derived from the IUPAC rules
instead of being spelled out, and accompanied by the compact *state-set codes* the device
works with: a tip is stored as one uint8 per pattern indexing a small look-up table of
0/1 rows, never as K-fold replicated fp64 vectors (the reference's tip layout would need
256 GB at 1000 taxa x 1M patterns - SURVEY.md 8(a) row a6).
"""
import numpy as np

DNA_STATES = "ACGT"
PROTEIN_STATES = "ARNDCQEGHILKMFPSTWYV"
BINARY_STATES = "01"

_IUPAC_DNA = {
    "A": "A", "C": "C", "G": "G", "T": "T", "U": "T",
    "R": "AG", "Y": "CT", "M": "AC", "K": "GT", "W": "AT", "S": "CG",
    "B": "CGT", "D": "AGT", "H": "ACT", "V": "ACG", "N": "ACGT", "-": "ACGT",
}


def _row(states, members):
    return [1.0 if s in members else 0.0 for s in states]


def _build_dna():
    table = {}
    for sym, members in _IUPAC_DNA.items():
        table[sym] = _row(DNA_STATES, members)
        if sym.isalpha():
            table[sym.lower()] = _row(DNA_STATES, members)
    return table


def _build_protein():
    table = {}
    for aa in PROTEIN_STATES:
        table[aa] = _row(PROTEIN_STATES, aa)
        table[aa.lower()] = _row(PROTEIN_STATES, aa)
    for wild in ("-", "?", "X", "x"):
        table[wild] = _row(PROTEIN_STATES, PROTEIN_STATES)
    return table


def _build_binary():
    return {"0": _row(BINARY_STATES, "0"), "1": _row(BINARY_STATES, "1"),
            "-": _row(BINARY_STATES, "01"), "N": _row(BINARY_STATES, "01")}


dna_charmap = _build_dna()
protein_charmap = _build_protein()
binary_charmap = _build_binary()


class CodeBook(object):
    """
    The distinct 0/1 rows of a charmap, ranked in lexicographic order of the row.

    ``lut[code]`` is the row; ``code_of[char]`` its rank.  Because the rank order equals the
    lexicographic order of the rows, sorting alignment columns by their tuple of codes
    gives exactly the column order ``np.unique(one_hot, axis=1)`` produces in the reference
    (alignment.py:48-51) - that is what makes compression on codes bit-exact.
    """

    def __init__(self, charmap):
        rows = sorted({tuple(v) for v in charmap.values()})
        self.lut = np.ascontiguousarray(rows, dtype=np.double)
        rank = {r: i for i, r in enumerate(rows)}
        self.code_of = {ch: rank[tuple(v)] for ch, v in charmap.items()}
        self.n_codes = len(rows)
        self.n_states = self.lut.shape[1]
        self.byte_table = np.full(256, 255, dtype=np.uint8)
        for ch, code in self.code_of.items():
            self.byte_table[ord(ch)] = code

    def encode(self, seq):
        """str -> uint8 codes; raises KeyError on a character outside the charmap (as the reference's dict lookup does)."""
        raw = np.frombuffer(seq.encode("latin-1"), dtype=np.uint8)
        codes = self.byte_table[raw]
        if codes.size and codes.max() == 255:
            bad = seq[int(np.argmax(codes == 255))]
            raise KeyError(bad)
        return codes


dna_codebook = CodeBook(dna_charmap)
protein_codebook = CodeBook(protein_charmap)
binary_codebook = CodeBook(binary_charmap)
