"""
ORACLE - TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py cpu_baseline / --impl reference).

Python face of the CPU restatement of the reference likelihood path:

* ctypes bindings of oracle/pruning_oracle.c (built by oracle/Makefile into oracle/_build/liboracle.so):
  ``clv``, ``lnl_node``, ``lnl_branch``, ``lnl_branch_derivs`` with the reference gufuncs' semantics
  (numba_likelihood_engine.py:10-87), and ``OracleTree`` = TreeModel.initialise / compute_partials /
  compute_likelihood_at_edge (tree_model.py:101-217) in the reference's own array layout;
* ``reference_compress`` = the np.unique call of alignment_to_numpy (alignment/alignment.py:48-51);
* ``ref_discrete_gamma`` = the reference's own C file (src/c_discrete_gamma.c) compiled untouched into
  oracle/_ref/libref_discrete_gamma.so.

Parity pinned by tests/test_oracle.py against tests/golden/*.npz, which oracle/make_golden.py wrote
by running the UNMODIFIED reference (numba engine through TreeModel) in the build container.
"""
import ctypes
import os
import subprocess

import numpy as np
from scipy.special import logsumexp

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "liboracle.so")
_REF_GAMMA = os.path.join(_HERE, "_ref", "libref_discrete_gamma.so")
_dp = ctypes.POINTER(ctypes.c_double)
_lp = ctypes.POINTER(ctypes.c_long)
_lib = None


def build():
    subprocess.run(["make", "-C", _HERE], check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(os.path.join(_HERE, "pruning_oracle.c")):
            build()
        _lib = ctypes.CDLL(_LIB)
        _lib.oracle_max_threads.restype = ctypes.c_int
    return _lib


def _d(a):
    return np.ascontiguousarray(a, dtype=np.double)


def _p(a):
    return a.ctypes.data_as(_dp)


def max_threads():
    return int(lib().oracle_max_threads())


# ---- operators ----------------------------------------------------------------------------------
def clv(p1, p2, clv1, clv2, scaler_a, scaler_b, cml_scaler, out=None, n_threads=0):
    p1, p2, clv1, clv2, scaler_a, scaler_b = map(_d, (p1, p2, clv1, clv2, scaler_a, scaler_b))
    K, A = p1.shape[0], p1.shape[1]
    S = int(np.prod(clv1.shape[:-2])) if clv1.ndim > 2 else 1
    if out is None:
        out = np.empty_like(clv1)
    assert cml_scaler.flags.c_contiguous and cml_scaler.dtype == np.double
    lib().oracle_clv(ctypes.c_long(S), K, A, _p(p1), _p(p2), _p(clv1), _p(clv2), _p(scaler_a), _p(scaler_b),
                     _p(cml_scaler), _p(out), int(n_threads))
    return out


def lnl_node(pi, partials, scale, n_threads=0):
    pi, partials, scale = map(_d, (pi, partials, scale))
    K, A = partials.shape[-2:]
    S = int(np.prod(partials.shape[:-2])) if partials.ndim > 2 else 1
    out = np.empty(partials.shape[:-1])
    lib().oracle_lnl_node(ctypes.c_long(S), K, A, _p(pi), _p(partials), _p(scale), _p(out), int(n_threads))
    return out


def _branch(probs, pi, a, b, sa, sb, nd):
    probs, pi, a, b = map(_d, (probs, pi, a, b))
    A = pi.shape[0]
    lead = a.shape[:-1]
    S = int(np.prod(lead)) if lead else 1
    sa = _d(np.broadcast_to(np.asarray(sa, dtype=np.double), lead or (1,)))
    sb = _d(np.broadcast_to(np.asarray(sb, dtype=np.double), lead or (1,)))
    out = np.empty((S, nd + 1))
    lib().oracle_lnl_branch(ctypes.c_long(S), A, nd, _p(probs), _p(pi), _p(a), _p(b), _p(sa), _p(sb), _p(out))
    return out.reshape(lead + ((3,) if nd else ()))


def lnl_branch(probs, pi, a, b, sa, sb):
    return _branch(probs, pi, a, b, sa, sb, 0)


def lnl_branch_derivs(probs, pi, a, b, sa, sb):
    return _branch(probs, pi, a, b, sa, sb, 2)


# ---- tree-level restatement -----------------------------------------------------------------------
class OracleTree(object):
    """
    TreeModel.initialise + compute_partials + compute_likelihood_at_edge in the reference's layout.

    tip_partials: dict node_id -> (S, A) float array (the alignment rows, alignment.py:26-37);
    rows: (n_rows, 3) PAR, CH1, CH2 in post-order; pmats: (n_rows, 2, K, A, A).
    """

    def __init__(self, n_nodes, tip_partials, n_cat, n_threads=0):
        first = next(iter(tip_partials.values()))
        S, A = first.shape
        self.S, self.A, self.K, self.n_threads = S, A, n_cat, int(n_threads)
        # tree_model.py:117-132
        self.partials = np.zeros((n_nodes, S, n_cat, A))
        self.scale = np.zeros((n_nodes, S, n_cat))
        self.root_partials = np.zeros((S, n_cat, A))
        self.root_scale = np.zeros((S, n_cat))
        for node, tp in tip_partials.items():          # tree_model.py:142-148
            for cat in range(n_cat):
                self.partials[node, :, cat, :] = tp

    def compute_partials(self, rows, pmats):
        rows = np.ascontiguousarray(rows, dtype=np.int64)
        pmats = _d(pmats)
        lib().oracle_compute_partials(rows.shape[0], rows.ctypes.data_as(_lp), _p(pmats), _p(self.partials),
                                      _p(self.scale), ctypes.c_long(self.S), self.K, self.A, self.n_threads)

    def likelihood_at_edge(self, a, b, root_pmats, freqs, cat_weights, want_cat=False):
        root_pmats, freqs, cat_weights = map(_d, (root_pmats, freqs, cat_weights))
        pattern = np.empty(self.S)
        cat = np.empty((self.S, self.K)) if want_cat else None
        lib().oracle_likelihood_at_edge(ctypes.c_long(int(a)), ctypes.c_long(int(b)), _p(root_pmats), _p(self.partials),
                                        _p(self.scale), _p(freqs), _p(cat_weights), ctypes.c_long(self.S), self.K,
                                        self.A, _p(self.root_partials), _p(self.root_scale),
                                        _p(cat) if want_cat else None, _p(pattern), self.n_threads)
        return (pattern, cat) if want_cat else pattern


def tree_lnl(traversal, tip_partials, model_p, freqs, rates, cat_weights, n_threads=0, return_tree=False):
    """
    Whole evaluation as bin/phy.py:140-146 does it: post-order, root on traversal.root_edge, mix.
    ``model_p(t, rates) -> (K, A, A)``.  Returns per-pattern lnL.
    """
    rows = np.asarray(traversal.postorder_traversal, dtype=np.int64)
    K = len(rates)
    n_nodes = 2 * len(traversal.names) - 2
    ot = OracleTree(n_nodes, tip_partials, K, n_threads)
    pm = np.empty((len(rows), 2, K, ot.A, ot.A))
    for i, (par, c1, c2) in enumerate(rows):
        pm[i, 0] = model_p(traversal.brlens[(int(par), int(c1))], rates)
        pm[i, 1] = model_p(traversal.brlens[(int(par), int(c2))], rates)
    ot.compute_partials(rows, pm)
    a, b = traversal.root_edge
    length = traversal.brlens[(a, b)]
    root_pm = np.stack([model_p(0, rates), model_p(length, rates)])
    pattern = ot.likelihood_at_edge(a, b, root_pm, freqs, cat_weights)
    return (pattern, ot) if return_tree else pattern


def mix_categories(cat_lnl, cat_weights):
    """tree_model.py:216"""
    return logsumexp(cat_lnl + np.log(cat_weights), axis=1)


# ---- compression ------------------------------------------------------------------------------------
def reference_compress(one_hot):
    """alignment/alignment.py:48-51 verbatim semantics: unique columns of the (ntax, nsite, A) float array."""
    patterns, inverse, counts = np.unique(one_hot, return_inverse=True, return_counts=True, axis=1)
    return patterns, counts, np.asarray(inverse).reshape(-1)


def discrete_gamma_scipy(alpha, ncat):
    """The reference's scipy formulation of Yang's mean-rate categories (gamma.py:4-18); agrees with the C code
    to ~1e-8 relative.  Used by bench.py's CPU arm only when oracle/_ref is absent."""
    from scipy.special import gammaincinv, gammainc
    cuts = gammaincinv(alpha, np.arange(1, ncat) / float(ncat)) / alpha          # quantiles of Gamma(alpha, rate alpha)
    upper = np.concatenate([[0.0], gammainc(alpha + 1.0, cuts * alpha), [1.0]])
    return np.diff(upper) * ncat


# ---- the reference's own C discrete gamma, compiled untouched -----------------------------------------
def have_ref_gamma():
    return os.path.exists(_REF_GAMMA)


def ref_discrete_gamma(alpha, ncat, median=False):
    """src/discrete_gamma.pyx:30-47 calling convention on top of src/c_discrete_gamma.c:285 DiscreteGamma."""
    h = ctypes.CDLL(_REF_GAMMA)
    h.DiscreteGamma.restype = ctypes.c_int
    h.DiscreteGamma.argtypes = [_dp, _dp, ctypes.c_double, ctypes.c_double, ctypes.c_int, ctypes.c_int]
    weights = np.zeros(ncat)
    rates = np.zeros(ncat)
    h.DiscreteGamma(_p(weights), _p(rates), float(alpha), float(alpha), int(ncat), 1 if median else 0)
    return rates
