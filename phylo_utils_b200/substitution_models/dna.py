"""
Nucleotide substitution models (state order A, C, G, T).

Constructors, attribute names and parameterisations follow the reference's per-model
modules under /root/reference/phylo_utils/substitution_models/ (jc69.py:8-45, tn93.py:8-82,
k80.py, f81.py, f84.py, hky85.py, gtr.py:9-46, strsym.py:9-43, unrest.py:9-22).

Deliberate supersets of the reference behaviour (SURVEY.md 8(a) "quirks"):
* ``JC69.p(t, rates=None)`` accepts ``rates`` (the reference's closed form does not, so
  ``TreeModel`` cannot drive its JC69 at all); with ``rates=None`` it is the same closed form.
* ``GTR`` also accepts a ready 4x4 exchangeability matrix (raises UnboundLocalError in the
  reference) and never mutates the caller's list.
"""
import numpy as np

from ..data import fixed_equal_nucleotide_rates, fixed_equal_nucleotide_frequencies
from .abstract import Eigen, DNAReversibleModel, DNANonReversibleModel
from .utils import check_frequencies, check_rates, compute_q_matrix, get_eigen, q_to_freqs

_OFFDIAG_ONES = np.ones((4, 4)) - np.eye(4)


class JC69(DNAReversibleModel):
    """Jukes-Cantor: one rate, equal frequencies; eigensystem known in closed form (reference: jc69.py:8-45)."""
    _name = 'JC69'
    _freqs = fixed_equal_nucleotide_frequencies.copy()

    def __init__(self):
        self._rates = check_rates(_OFFDIAG_ONES.copy(), 4)
        self._freqs = fixed_equal_nucleotide_frequencies.copy()
        self._q_mtx = (_OFFDIAG_ONES - 3.0 * np.eye(4)) / 3.0
        evecs = np.ascontiguousarray([[1., 2., 0., .5],
                                      [1., 2., 0., -.5],
                                      [1., -2., .5, 0.],
                                      [1., -2., -.5, 0.]])
        ivecs = np.asfortranarray([[.25, .25, .25, .25],
                                   [.125, .125, -.125, -.125],
                                   [0., 0., 1., -1.],
                                   [1., -1., 0., 0.]])
        evals = np.ascontiguousarray([0., -4. / 3, -4. / 3, -4. / 3])
        self.eigen = Eigen(evecs, evals, ivecs)

    @staticmethod
    def _closed_form(t):
        decay = np.exp(-4.0 * t / 3.0)
        same, diff = 0.25 + 0.75 * decay, 0.25 - 0.25 * decay
        return np.full((4, 4), diff) + (same - diff) * np.eye(4)

    def p(self, t, rates=None):
        if rates is None:
            return self._closed_form(t)
        return np.stack([self._closed_form(t * r) for r in rates], axis=0)


class TN93(DNAReversibleModel):
    """
    Tamura-Nei 1993: transition rates alpha_y (C<->T) and alpha_r (A<->G), transversion rate
    beta.  Q and its eigensystem are analytic (reference: tn93.py:8-82).
    """
    _name = 'TN93'

    def __init__(self, alpha_y, alpha_r, beta=1.0, freqs=None, scale_q=True):
        if freqs is None:
            freqs = fixed_equal_nucleotide_frequencies.copy()
        else:
            freqs = check_frequencies(freqs, 4)
        self._freqs = freqs
        self._alpha_y, self._alpha_r, self._beta = alpha_y, alpha_r, beta
        a, c, g, t = (freqs[i] for i in range(4))
        pur, pyr = a + g, c + t

        exch = np.array([[0, beta, alpha_r, beta],
                         [beta, 0, beta, alpha_y],
                         [alpha_r, beta, 0, beta],
                         [beta, alpha_y, beta, 0]])
        self._rates = check_rates(exch, 4)

        if scale_q:
            norm = 2 * (alpha_y * c * t + beta * a * t + beta * a * c + alpha_r * a * g + beta * g * t + beta * c * g)
        else:
            norm = 1.0

        q = np.ascontiguousarray([
            [-(alpha_r * g + beta * pyr), beta * c, alpha_r * g, beta * t],
            [beta * a, -(alpha_y * t + beta * pur), beta * g, alpha_y * t],
            [alpha_r * a, beta * c, -(alpha_r * a + beta * pyr), beta * t],
            [beta * a, alpha_y * c, beta * g, -(alpha_y * c + beta * pur)]])
        self._q_mtx = q / norm

        evecs = np.ascontiguousarray([[1, -1 / pur, g / pur, 0],
                                      [1, 1 / pyr, 0, -t / pyr],
                                      [1, -1 / pur, -a / pur, 0],
                                      [1, 1 / pyr, 0, c / pyr]], dtype=np.double)
        ivecs = np.asfortranarray([[a, c, g, t],
                                   [-a * pyr, c * pur, -g * pyr, t * pur],
                                   [1, 0, -1, 0],
                                   [0, -1, 0, 1]], dtype=np.double)
        evals = np.ascontiguousarray([0,
                                      -beta,
                                      -(pur * alpha_r + pyr * beta),
                                      -(pyr * alpha_y + pur * beta)], dtype=np.double) / norm
        self.eigen = Eigen(evecs, evals, ivecs)


class K80(TN93):
    _name = 'K80'
    _freqs = fixed_equal_nucleotide_frequencies.copy()

    def __init__(self, kappa, scale_q=True):
        TN93.__init__(self, kappa, kappa, 1, fixed_equal_nucleotide_frequencies.copy(), scale_q=scale_q)


class F81(TN93):
    _name = 'F81'

    def __init__(self, freqs, scale_q=True):
        TN93.__init__(self, 1, 1, 1, freqs, scale_q=scale_q)


class F84(TN93):
    _name = 'F84'

    def __init__(self, kappa, freqs, scale_q=True):
        TN93.__init__(self,
                      1 + kappa / (freqs[1] + freqs[3]),
                      1 + kappa / (freqs[0] + freqs[2]),
                      1, freqs, scale_q=scale_q)


class HKY85(TN93):
    _name = 'HKY85'

    def __init__(self, kappa, freqs, scale_q=True):
        TN93.__init__(self, kappa, kappa, 1, freqs, scale_q=scale_q)


def _exchangeabilities_from_upper(values):
    """[AC, AG, AT, CG, CT, GT] -> symmetric 4x4, zero diagonal."""
    m = np.zeros((4, 4))
    iu = np.triu_indices(4, 1)
    m[iu] = values
    return m + m.T


class GTR(DNAReversibleModel):
    """General time-reversible model; numeric symmetric eigendecomposition (reference: gtr.py:9-46)."""
    _name = 'GTR'

    def __init__(self, rates=None, freqs=None, scale_q=True):
        if rates is None:
            rates_m = fixed_equal_nucleotide_rates.copy()
        else:
            arr = np.asarray(rates, dtype=np.double)
            if arr.shape == (4, 4):
                rates_m = arr.copy()
            elif arr.shape == (6,):
                rates_m = _exchangeabilities_from_upper(arr)
            elif arr.shape == (5,):
                rates_m = _exchangeabilities_from_upper(np.append(arr, 1.0))
            else:
                raise ValueError('GTR rates must be 5 or 6 values (AC,AG,AT,CG,CT[,GT]) or a 4x4 matrix')
        if freqs is None:
            freqs = fixed_equal_nucleotide_frequencies.copy()
        self._rates = check_rates(np.ascontiguousarray(rates_m), self.size)
        self._freqs = check_frequencies(freqs, self.size)
        self._q_mtx = compute_q_matrix(self._rates, self._freqs, scale_q)
        self.eigen = Eigen(*get_eigen(self._q_mtx, self._freqs))

    def square_matrix(self, uppertri):
        return _exchangeabilities_from_upper(np.asarray(uppertri, dtype=np.double))


class _NonReversible(DNANonReversibleModel):
    def _finish(self):
        self._q_mtx = compute_q_matrix(self._rates, None)
        self.eigen = Eigen(*get_eigen(self._q_mtx))

    @property
    def freqs(self):
        return q_to_freqs(self._q_mtx)


class Strsym(_NonReversible):
    """
    Strand-symmetric model: six rates, a->b equals complement(a)->complement(b)
    (reference: strsym.py:9-43).  ``rates`` = [A>C, A>G, A>T, C>A, C>G, C>T].
    """
    _name = 'STRSYM'

    def __init__(self, rates=None):
        if rates is None:
            rates = np.ones(6)
        if len(rates) != 6:
            raise ValueError("Provide a list of 6 rate parameters")
        m = np.zeros((4, 4))
        src = np.array([0, 0, 0, 1, 1, 1])
        dst = np.array([1, 2, 3, 0, 2, 3])
        m[src, dst] = rates
        m[3 - src, 3 - dst] = rates          # complement: A<->T, C<->G is i -> 3 - i
        DNANonReversibleModel.__init__(self, m)
        self._finish()


class Unrest(_NonReversible):
    """Unrestricted 12-rate model (reference: unrest.py:9-22)."""
    _name = 'UNREST'

    def __init__(self, rates=None):
        if rates is None:
            rates = fixed_equal_nucleotide_rates.copy()
        DNANonReversibleModel.__init__(self, np.ascontiguousarray(rates, dtype=np.double))
        self._finish()
