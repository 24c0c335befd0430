"""
Model / Eigen base classes - the object surface the likelihood path reads:
``model.eigen.{evecs,evals,ivecs}``, ``model.freqs``, ``model.size`` and
``model.p / dp_dt / d2p_dt2 (t, rates)`` returning ``(K, A, A)`` C-contiguous stacks
(reference: /root/reference/phylo_utils/substitution_models/abstract.py:11-198).

These host methods are the *API* and the small-problem convenience; the tree path does
not call them per edge - :class:`phylo_utils_b200.tree_model.TreeModel` ships the
eigensystem to the device once and builds P for all edges x categories in one kernel
(csrc/pmatrix.cu).  Models without a real eigensystem (``has_real_eigensystem`` False:
the non-reversible DNA models, which the reference exponentiates with a Taylor scheme)
are exponentiated on the host and uploaded as ready-made matrices.
"""
import numpy as np

from .utils import compute_b_matrix, check_frequencies, check_rates, compute_q_matrix, get_eigen, expm

MIN_BRANCH_LENGTH = 1 / 2 ** 16


class Eigen(object):
    """Q = evecs . diag(evals) . ivecs (reference: abstract.py:88-122)."""
    __slots__ = ['evals', 'evecs', 'ivecs']

    def __init__(self, evecs, evals, ivecs):
        self.evecs = evecs
        self.evals = evals
        self.ivecs = ivecs

    @property
    def values(self):
        return self.evecs, self.evals, self.ivecs

    def fn_apply(self, fn):
        """evecs . diag(fn(evals)) . ivecs"""
        return (self.evecs * fn(self.evals)).dot(self.ivecs)

    def exp(self, t=1.0):
        return (self.evecs * np.exp(self.evals * t)).dot(self.ivecs)

    def reconstitute(self):
        return (self.evecs * self.evals).dot(self.ivecs)


class Model(object):
    _name = None
    _rates = None
    _freqs = None
    _size = None
    _states = None
    has_real_eigensystem = True

    def __init__(self):
        self.eigen = None
        self._q_mtx = None

    name = property(lambda self: self._name)
    rates = property(lambda self: self._rates)
    freqs = property(lambda self: self._freqs)
    size = property(lambda self: self._size)
    states = property(lambda self: self._states)

    def q(self):
        return self._q_mtx

    def b(self):
        return compute_b_matrix(self.q(), np.sqrt(self.freqs))

    # -- transition probabilities and their derivatives ---------------------------------
    def _stack(self, t, rates, fn):
        if rates is None:
            return self.eigen.fn_apply(lambda lam: fn(lam, t))
        return np.stack([self.eigen.fn_apply(lambda lam, s=t * r: fn(lam, s)) for r in rates], axis=0)

    def p(self, t, rates=None):
        """P(t) = exp(Qt); with ``rates`` a (K, A, A) stack of P(t*r_k) (reference: abstract.py:49-59)."""
        if rates is None:
            return self.eigen.exp(t)
        return np.stack([self.eigen.exp(t * r) for r in rates], axis=0)

    def dp_dt(self, t, rates=None):
        """
        Q exp(Q t r_k) per category.  As in the reference (abstract.py:61-68) this is the
        derivative with respect to the *scaled* time t*r_k: the chain-rule factor r_k is
        not applied.  The device derivative kernels apply it (see tree_model.edge_derivatives).
        """
        return self._stack(t, rates, lambda lam, s: lam * np.exp(lam * s))

    def d2p_dt2(self, t, rates=None):
        """Q^2 exp(Q t r_k) per category; same convention as dp_dt (reference: abstract.py:70-77)."""
        return self._stack(t, rates, lambda lam, s: lam * lam * np.exp(lam * s))

    def detailed_balance(self):
        """pi_i q_ij == pi_j q_ji ? (reference: abstract.py:79-85)"""
        flux = self.q().T * self.freqs
        return bool(np.allclose(flux, flux.T))


class ProteinModel(Model):
    _name = 'GenericProtein'
    _size = 20
    _states = list('ARNDCQEGHILKMFPSTWYV')

    def __init__(self, rates, freqs):
        self._rates = check_rates(rates, self.size)
        self._freqs = check_frequencies(freqs, self.size)
        self._q_mtx = compute_q_matrix(self._rates, self._freqs)
        self.eigen = Eigen(*get_eigen(self._q_mtx, self._freqs))

    def __repr__(self):
        return 'Protein model: {}\nFreqs:        {}\n'.format(self._name, self._freqs)


class DNAReversibleModel(Model):
    _name = 'GenericReversibleDNA'
    _size = 4
    _states = list('ACGT')

    def __init__(self, rates, freqs):
        self._rates = check_rates(rates, self.size)
        self._freqs = check_frequencies(freqs, self.size)

    def __repr__(self):
        upper = self._rates[np.triu_indices(4, 1)]
        return 'DNA reversible model: {}\nRel. rates: {}\nFreqs:      {}\n'.format(self._name, upper, self._freqs)


class DNANonReversibleModel(Model):
    """P and its derivatives come from the Taylor ``expm`` (reference: abstract.py:160-198)."""
    _name = 'GenericNonReversibleDNA'
    _size = 4
    _states = list('ACGT')
    has_real_eigensystem = False

    def __init__(self, rates):
        self._rates = check_rates(rates, self.size, symmetry=False)

    def _expm_stack(self, t, rates, left):
        q = self.q()
        if rates is None:
            return left.dot(expm(q * t)) if left is not None else expm(q * t)
        mats = [expm(q * r * t) for r in rates]
        if left is not None:
            mats = [left.dot(m) for m in mats]
        return np.stack(mats, axis=0)

    def p(self, t, rates=None):
        return self._expm_stack(t, rates, None)

    def dp_dt(self, t, rates=None):
        return self._expm_stack(t, rates, self.q())

    def d2p_dt2(self, t, rates=None):
        q = self.q()
        return self._expm_stack(t, rates, q.dot(q))

    def __repr__(self):
        off = self._rates[~np.eye(4, dtype=bool)]
        return 'DNA non-reversible model: {}\nRel. rates: {}\nFreqs:      {}\n'.format(self._name, off, self.freqs)
