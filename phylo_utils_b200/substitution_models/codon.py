"""
Goldman-Yang 1994 codon model over the 61 sense codons of the standard genetic code.

NOT in the reference (SURVEY.md headline fact 5) - BASELINE config 4 asks for it, so it is
built on the same ``Model`` / ``Eigen`` surface:

    q_ij = 0                      if codons i, j differ at more than one position
         = pi_j                   synonymous transversion
         = kappa pi_j             synonymous transition
         = omega pi_j             non-synonymous transversion
         = omega kappa pi_j       non-synonymous transition

scaled to one expected substitution per codon.  The model is reversible w.r.t. pi, so the
symmetric eigendecomposition of the other reversible models applies.  "Parity unpinned":
there is no reference implementation to compare Q against; tests check detailed balance,
the rate normalisation and the kappa/omega structure, and engine parity at A=61 is checked
against the oracle with the same P matrices.
"""
import numpy as np

from .abstract import Eigen, Model
from .utils import check_frequencies, compute_q_matrix, get_eigen

_BASES = "TCAG"
_AA = ("FFLLSSSSYY**CC*W" "LLLLPPPPHHQQRRRR" "IIIMTTTTNNKKSSRR" "VVVVAAAADDEEGGGG")
_ALL = [a + b + c for a in _BASES for b in _BASES for c in _BASES]
GENETIC_CODE = dict(zip(_ALL, _AA))
SENSE_CODONS = [cod for cod in sorted(_ALL, key=lambda s: ["ACGT".index(ch) for ch in s]) if GENETIC_CODE[cod] != "*"]
_TRANSITIONS = {frozenset("AG"), frozenset("CT")}


def f3x4(position_freqs):
    """(3, 4) nucleotide frequencies per codon position (order ACGT) -> 61 codon frequencies."""
    pf = np.asarray(position_freqs, dtype=np.double)
    raw = np.array([pf[0, "ACGT".index(c[0])] * pf[1, "ACGT".index(c[1])] * pf[2, "ACGT".index(c[2])]
                    for c in SENSE_CODONS])
    return raw / raw.sum()


class GY94(Model):
    _name = 'GY94'
    _size = 61
    _states = list(SENSE_CODONS)

    def __init__(self, kappa=2.0, omega=0.2, freqs=None, scale_q=True):
        n = self._size
        if freqs is None:
            freqs = np.full(n, 1.0 / n)
        freqs = np.asarray(freqs, dtype=np.double)
        freqs = freqs / freqs.sum()
        self._freqs = check_frequencies(freqs, n)
        self.kappa, self.omega = float(kappa), float(omega)
        exch = np.zeros((n, n))
        for i, ci in enumerate(SENSE_CODONS):
            for j in range(i + 1, n):
                cj = SENSE_CODONS[j]
                diff = [(x, y) for x, y in zip(ci, cj) if x != y]
                if len(diff) != 1:
                    continue
                r = 1.0
                if frozenset(diff[0]) in _TRANSITIONS:
                    r *= self.kappa
                if GENETIC_CODE[ci] != GENETIC_CODE[cj]:
                    r *= self.omega
                exch[i, j] = exch[j, i] = r
        self._rates = exch
        self._q_mtx = compute_q_matrix(exch, self._freqs, scale_q)
        self.eigen = Eigen(*get_eigen(self._q_mtx, self._freqs))

    def __repr__(self):
        return 'Codon model: GY94 kappa={} omega={}\n'.format(self.kappa, self.omega)
