"""
Rate-matrix algebra shared by the substitution models.  Function names and results follow
/root/reference/phylo_utils/substitution_models/utils.py:5-116; the implementations are
independent.
"""
import numpy as np

SMALL = 2.0 ** -128


def compute_b_matrix(q_matrix, sqrtfreqs):
    """B = D^(1/2) Q D^(-1/2); symmetric iff Q is reversible w.r.t. freqs (reference: utils.py:5-12)."""
    s = np.asarray(sqrtfreqs, dtype=np.double)
    return (s[:, None] * np.asarray(q_matrix, dtype=np.double)) * (1.0 / s)[None, :]


def check_frequencies(freqs, length):
    """Reference: utils.py:15-23 - same three ValueErrors, same (tight) sum tolerance."""
    freqs = np.ascontiguousarray(freqs)
    if len(freqs) != length:
        raise ValueError('Frequencies vector is not the right length (length={})'.format(len(freqs)))
    if np.min(freqs) < 0:
        raise ValueError('Frequencies vector contains negative values')
    total = sum(freqs)
    if not np.allclose(total, 1.0, rtol=1e-16):
        raise ValueError('Frequencies do not add to 1.0 within tolerance (sum={})'.format(total))
    return freqs


def check_rates(rates, size, symmetry=True):
    """Reference: utils.py:26-34."""
    rates = np.ascontiguousarray(rates)
    if rates.shape != (size, size):
        raise ValueError('Rate matrix is not the right shape (length={})'.format(rates.shape))
    if np.min(rates) < 0:
        raise ValueError('Rate matrix contains negative values')
    if symmetry and not np.allclose(rates, rates.T):
        raise ValueError('Rate matrix is not symmetrical')
    return rates


def impose_min_probs(mtx):
    """Reference: utils.py:37-42."""
    if np.min(mtx) >= SMALL:
        return mtx
    clipped = np.clip(mtx, SMALL, 1.0)
    clipped /= clipped.sum(axis=1, keepdims=True)
    return 0.5 * (clipped + clipped.T)


def q_to_freqs(q_matrix):
    """Stationary distribution: least-squares solution of [1..1; Q^T] pi = [1; 0..0] (reference: utils.py:68-79)."""
    n = q_matrix.shape[0]
    lhs = np.vstack([np.ones((1, n)), np.asarray(q_matrix, dtype=np.double).T])
    rhs = np.zeros(n + 1)
    rhs[0] = 1.0
    pi, _, _, _ = np.linalg.lstsq(lhs, rhs, rcond=None)
    return pi


def compute_q_matrix(rates, freqs, scale=True):
    """
    Q_ij = rates_ij * freqs_j (or rates_ij when freqs is None), rows made to sum to zero,
    optionally normalised to one expected substitution per unit time (reference: utils.py:45-65).
    """
    q = np.array(rates, dtype=np.double, copy=True)
    if freqs is not None:
        q = q * np.asarray(freqs, dtype=np.double)[None, :]
    if q.ndim != 2 or q.shape[0] != q.shape[1]:
        raise AssertionError('Q is not square')
    idx = np.arange(q.shape[0])
    q[idx, idx] -= q.sum(axis=1)
    if scale:
        pi = q_to_freqs(q) if freqs is None else np.asarray(freqs, dtype=np.double)
        q /= -(np.diag(q) @ pi)
    return q


def get_eigen(q_matrix, freqs=None):
    """
    -> (evecs C-order, evals, ivecs F-order) with Q = evecs diag(evals) ivecs
    (reference: utils.py:82-98).  Reversible case goes through the symmetric B matrix and
    ``eigh``; otherwise a general ``eig`` sorted by eigenvalue plus an explicit inverse.
    """
    if freqs is not None:
        root = np.sqrt(np.asarray(freqs, dtype=np.double))
        evals, r = np.linalg.eigh(compute_b_matrix(q_matrix, root))
        evecs = (1.0 / root)[:, None] * r
        ivecs = r.T * root[None, :]
    else:
        evals, evecs = np.linalg.eig(q_matrix)
        order = np.argsort(evals)
        evals, evecs = evals[order], evecs[:, order]
        ivecs = np.linalg.inv(evecs)
    return np.ascontiguousarray(evecs), np.ascontiguousarray(evals), np.asfortranarray(ivecs)


def expm(matrix):
    """
    exp(M) by scaling (2^-8), a 4th-order Taylor polynomial and 8 squarings - the RevBayes
    scheme the reference uses for non-reversible models (reference: utils.py:101-116).
    """
    squarings = 8
    m = np.asarray(matrix, dtype=np.double) / float(2 ** squarings)
    m2 = m @ m
    m3 = m @ m2
    m4 = m @ m3
    out = m + (np.eye(m.shape[0]) + m2 / 2.0 + m3 / 6.0 + m4 / 24.0)
    for _ in range(squarings):
        out = out @ out
    return out
