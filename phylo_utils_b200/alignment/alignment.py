"""
Alignment -> tip data + site-pattern compression.

Mirrors the public functions of /root/reference/phylo_utils/alignment/alignment.py:26-66
(``seq_to_partials``, ``alignment_to_numpy``, ``invariant_sites``, ``read_alignment``) and adds
the compact form the GPU path actually consumes (``alignment_to_codes``): one uint8
state-set code per (taxon, pattern) plus a tiny look-up table.

Compression is *bit-exact* with the reference: the reference runs
``np.unique(one_hot, axis=1, return_inverse=True, return_counts=True)`` on the
``(ntax, nsite, A)`` float array (alignment.py:48-51), i.e. it sorts columns
lexicographically over the taxon-major / state-minor flattening.  Codes are ranks of the
0/1 rows in lexicographic order (charmaps.CodeBook), so sorting columns by their byte
string of codes yields the same pattern order, inverse index and weights - checked against
the reference itself in tests/test_alignment.py and tests/golden/.

File reading needs no Biopython: a small FASTA / relaxed-PHYLIP reader is included.
"""
import numpy as np

from .alphabets import DNA, PROTEIN, BINARY
from .charmaps import (dna_charmap, protein_charmap, binary_charmap,
                       dna_codebook, protein_codebook, binary_codebook)
from ..utils import setup_logger

logger = setup_logger()


class SeqRecord(object):
    """Just enough of Bio.SeqRecord: ``.name``, ``.id`` and ``.seq`` (str-able)."""
    __slots__ = ("name", "id", "seq")

    def __init__(self, name, seq):
        self.name = name
        self.id = name
        self.seq = seq

    def __len__(self):
        return len(self.seq)


def read_alignment(filename, format="fasta", alphabet=None):
    """-> list of SeqRecord (reference: alignment.py:15-17, minus Biopython)."""
    with open(filename) as fh:
        text = fh.read()
    fmt = format.lower()
    if fmt == "fasta":
        records, name, chunks = [], None, []
        for line in text.splitlines():
            line = line.strip()
            if not line:
                continue
            if line.startswith(">"):
                if name is not None:
                    records.append(SeqRecord(name, "".join(chunks)))
                name, chunks = line[1:].split()[0], []
            else:
                chunks.append(line.replace(" ", ""))
        if name is not None:
            records.append(SeqRecord(name, "".join(chunks)))
    elif fmt in ("phylip", "phylip-relaxed", "phylip-sequential"):
        lines = [ln for ln in text.splitlines() if ln.strip()]
        ntax, nsite = (int(v) for v in lines[0].split()[:2])
        records = []
        for ln in lines[1:1 + ntax]:
            parts = ln.split(None, 1)
            records.append(SeqRecord(parts[0], parts[1].replace(" ", "") if len(parts) > 1 else ""))
        for k, ln in enumerate(lines[1 + ntax:]):       # interleaved continuation blocks
            records[k % ntax].seq += ln.replace(" ", "")
        if any(len(r.seq) != nsite for r in records):
            raise ValueError("PHYLIP header says {} sites but a sequence differs".format(nsite))
    else:
        raise ValueError("unsupported alignment format {!r}".format(format))
    if records and len({len(r.seq) for r in records}) != 1:
        raise ValueError("sequences in {} have unequal lengths".format(filename))
    return records


def sample_characters(alignment, seqlen=1000, nseq=100):
    return set("".join(str(rec.seq)[:seqlen] for rec in list(alignment)[:nseq]))


def guess_alphabet(charsample):
    dna_like = set(dna_charmap) | set("-")
    return PROTEIN if len(set(charsample) - dna_like) > 0 else DNA


def _charmap_for(alphabet):
    if alphabet == DNA:
        return dna_charmap
    if alphabet == PROTEIN:
        return protein_charmap
    if alphabet == BINARY:
        return binary_charmap
    logger.warning("Unrecognised alphabet. Trying DNA")
    return dna_charmap


def codebook_for(alphabet):
    if alphabet == DNA:
        return dna_codebook
    if alphabet == PROTEIN:
        return protein_codebook
    if alphabet == BINARY:
        return binary_codebook
    logger.warning("Unrecognised alphabet. Trying DNA")
    return dna_codebook


def seq_to_partials(seq, alphabet):
    """str -> (len, A) C-contiguous float64 0/1 rows (reference: alignment.py:26-37)."""
    book = codebook_for(alphabet)
    return np.ascontiguousarray(book.lut[book.encode(str(seq))])


def seq_to_codes(seq, alphabet):
    return codebook_for(alphabet).encode(str(seq))


def compress_codes(codes):
    """
    uint8 ``(ntax, nsite)`` -> (patterns ``(ntax, npat)``, siteweights ``(npat,)`` int64,
    inverse_index ``(nsite,)`` int64).  Columns are compared as unsigned byte strings,
    which is the lexicographic order over taxa the reference's np.unique(axis=1) uses.
    """
    codes = np.ascontiguousarray(codes, dtype=np.uint8)
    ntax, nsite = codes.shape
    if nsite == 0:
        return codes.copy(), np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int64)
    cols = np.ascontiguousarray(codes.T)
    keys = cols.view(np.dtype((np.void, ntax))).ravel()
    order = np.argsort(keys, kind="stable")
    sorted_cols = cols[order]
    is_new = np.empty(nsite, dtype=bool)
    is_new[0] = True
    np.any(sorted_cols[1:] != sorted_cols[:-1], axis=1, out=is_new[1:])
    group = np.cumsum(is_new) - 1
    inverse = np.empty(nsite, dtype=np.int64)
    inverse[order] = group
    starts = np.flatnonzero(is_new)
    weights = np.diff(np.concatenate([starts, [nsite]])).astype(np.int64)
    patterns = np.ascontiguousarray(sorted_cols[starts].T)
    return patterns, weights, inverse


def compress_codes_gpu(data, byte_table=None, device=0):
    """
    ``compress_codes`` on the GPU (C ABI ``phb_compress_patterns``, csrc/compress.cu): radix sort of the columns,
    run detection, weights, inverse index - bit-identical outputs.  ``data`` is uint8 ``(ntax, nsite)``: codes, or
    raw characters when ``byte_table`` (256 uint8, 255 = unmapped) is given; an unmapped character raises KeyError
    like the reference's charmap look-up.
    """
    import ctypes
    from .._lib import lib, check
    data = np.ascontiguousarray(data, dtype=np.uint8)
    if data.ndim != 2 or data.shape[0] < 1:
        raise ValueError("data must be (ntax, nsite) with at least one row")
    ntax, nsite = data.shape
    patterns = np.empty((ntax, nsite), dtype=np.uint8)
    weights = np.empty(nsite, dtype=np.int64)
    inverse = np.empty(nsite, dtype=np.int64)
    npat, bad = ctypes.c_int64(0), ctypes.c_int64(-1)
    table = None if byte_table is None else np.ascontiguousarray(byte_table, dtype=np.uint8)
    if table is not None and table.shape != (256,):
        raise ValueError("byte_table must have 256 entries")
    i64p = ctypes.POINTER(ctypes.c_int64)
    status = lib().phb_compress_patterns(
        int(device), ctypes.c_void_p(data.ctypes.data), None if table is None else ctypes.c_void_p(table.ctypes.data),
        ntax, nsite, ctypes.c_void_p(patterns.ctypes.data), weights.ctypes.data_as(i64p), inverse.ctypes.data_as(i64p),
        ctypes.byref(npat), ctypes.byref(bad))
    if bad.value >= 0:
        raise KeyError(chr(int(data.reshape(-1)[bad.value])))
    check(status)
    n = int(npat.value)
    return np.ascontiguousarray(patterns.reshape(-1)[:ntax * n].reshape(ntax, n)), weights[:n].copy(), inverse


def alignment_to_codes(alignment, alphabet, compress=True, device=None):
    """
    -> (codes ``(ntax, npat)`` uint8, lut ``(ncodes, A)`` float64, siteweights int64,
        inverse_index int64, names ``{label: row}``)

    ``device`` = a CUDA device index runs the character look-up and the compression on that GPU
    (``compress_codes_gpu``); None keeps them on the host.  Same outputs either way.
    """
    book = codebook_for(alphabet)
    records = list(alignment)
    names = {rec.name: i for i, rec in enumerate(records)}
    if device is not None and compress and records and len(records[0].seq) > 0:
        raw = [np.frombuffer(str(rec.seq).encode("latin-1"), dtype=np.uint8) for rec in records]
        if len({len(r) for r in raw}) != 1:
            raise ValueError("sequences have unequal lengths")
        codes, weights, inverse = compress_codes_gpu(np.stack(raw), book.byte_table, device)
        return codes, book.lut, weights, inverse, names
    rows = [book.encode(str(rec.seq)) for rec in records]
    if rows and len({len(r) for r in rows}) != 1:
        raise ValueError("sequences have unequal lengths")
    codes = np.stack(rows) if rows else np.zeros((0, 0), dtype=np.uint8)
    nsite = codes.shape[1]
    if compress:
        codes, weights, inverse = compress_codes(codes)
    else:
        weights = np.ones(nsite, dtype=np.int64)
        inverse = np.arange(nsite, dtype=np.int64)
    return codes, book.lut, weights, inverse, names


def alignment_to_numpy(alignment, alphabet, compress=True):
    """
    Reference signature and return order (alignment.py:40-57):
    (alignment ``(ntax, npat, A)`` float64, siteweights, inverse_index, names).
    """
    codes, lut, weights, inverse, names = alignment_to_codes(alignment, alphabet, compress)
    return np.ascontiguousarray(lut[codes]), weights, inverse, names


def invariant_sites(alignment):
    """
    Boolean per column: some state is compatible with every taxon (reference: alignment.py:59-66).
    ``alignment`` is the ``(ntax, nsite, A)`` float array.
    """
    aln = np.asarray(alignment)
    return list(np.any(np.all(aln != 0, axis=0), axis=1))
