"""
The four likelihood operators under the reference's names and call signatures
(/root/reference/phylo_utils/likelihood/numba_likelihood_engine.py:10-87), executed by CUDA
kernels through the C ABI (csrc/ops.cu).  numpy arrays in, numpy arrays out, leading
(pattern) dimension optional - i.e. the gufunc calling convention of the originals:

    clv(p1, p2, clv1, clv2, scaler_a, scaler_b, cml_scaler[, out]) -> out      (K,A,A)x2, ([S],K,A)x2, ([S],K)x3
    lnl_node(pi, partials, scale[, out]) -> ([S],K)
    lnl_branch(probs(A,A), pi, a, b, sa, sb[, out]) -> ([S])
    lnl_branch_derivs(probs(3,A,A), pi, a, b, sa, sb[, out]) -> ([S],3)

As in the reference, ``cml_scaler`` is written IN PLACE (tree_model.py:176 relies on that), so it
must be a C-contiguous float64 array.  These are per-call host round trips meant for API parity
and small inputs; TreeModel keeps everything resident on the device instead.
"""
import numpy as np

from .._lib import lib, check, dptr

SCALE_THRESHOLD = 1.0 / 2.0 ** 128
_DEVICE = 0


def set_device(index):
    global _DEVICE
    _DEVICE = int(index)


def _c(a):
    return np.ascontiguousarray(a, dtype=np.double)


def _lead(arr, core_ndim):
    """Split gufunc loop dimensions from core dimensions; returns (flat_leading_count, leading_shape)."""
    lead = arr.shape[:arr.ndim - core_ndim]
    n = 1
    for d in lead:
        n *= d
    return n, lead


def clv(p1, p2, clv1, clv2, scaler_a, scaler_b, cml_scaler, out=None):
    p1, p2, clv1, clv2 = _c(p1), _c(p2), _c(clv1), _c(clv2)
    scaler_a, scaler_b = _c(scaler_a), _c(scaler_b)
    if p1.ndim != 3 or p1.shape != p2.shape or p1.shape[1] != p1.shape[2]:
        raise ValueError("p1 and p2 must both be (ncat, nstate, nstate)")
    K, A = p1.shape[0], p1.shape[1]
    if clv1.shape != clv2.shape or clv1.shape[-2:] != (K, A):
        raise ValueError("clv1 and clv2 must both be ([nsites], ncat, nstate)")
    S, lead = _lead(clv1, 2)
    if scaler_a.shape != lead + (K,) or scaler_b.shape != lead + (K,):
        raise ValueError("scalers must be ([nsites], ncat)")
    if not (isinstance(cml_scaler, np.ndarray) and cml_scaler.dtype == np.double and
            cml_scaler.flags.c_contiguous and cml_scaler.shape == lead + (K,)):
        raise ValueError("cml_scaler must be a C-contiguous float64 array of shape ([nsites], ncat); it is written in place")
    if out is None:
        out = np.empty_like(clv1)
    elif not (out.dtype == np.double and out.flags.c_contiguous and out.shape == clv1.shape):
        raise ValueError("out must be C-contiguous float64 with the shape of clv1")
    check(lib().phb_op_clv(_DEVICE, S, K, A, dptr(p1), dptr(p2), dptr(clv1), dptr(clv2), dptr(scaler_a),
                           dptr(scaler_b), dptr(cml_scaler), dptr(out)))
    return out


def lnl_node(pi, partials, scale, out=None):
    pi, partials, scale = _c(pi), _c(partials), _c(scale)
    if partials.ndim < 2 or pi.shape != (partials.shape[-1],):
        raise ValueError("partials must be ([nsites], ncat, nstate) and pi (nstate,)")
    K, A = partials.shape[-2:]
    S, lead = _lead(partials, 2)
    if scale.shape != lead + (K,):
        raise ValueError("scale must be ([nsites], ncat)")
    if out is None:
        out = np.empty(lead + (K,))
    check(lib().phb_op_lnl_node(_DEVICE, S, K, A, dptr(pi), dptr(partials), dptr(scale), dptr(out)))
    return out


def _branch(probs, pi, partials_a, partials_b, scale_a, scale_b, out, n_derivs):
    probs, pi = _c(probs), _c(pi)
    partials_a, partials_b = _c(partials_a), _c(partials_b)
    A = pi.shape[0]
    want = (3, A, A) if n_derivs else (A, A)
    if probs.shape != want:
        raise ValueError("probs must be {}".format(want))
    if partials_a.shape != partials_b.shape or partials_a.shape[-1] != A:
        raise ValueError("partials_a and partials_b must both be ([nsites], nstate)")
    S, lead = _lead(partials_a, 1)
    scale_a = np.ascontiguousarray(np.broadcast_to(np.asarray(scale_a, dtype=np.double).reshape(lead or (1,))
                                                   if np.ndim(scale_a) else np.full(lead or (1,), float(scale_a)),
                                                   lead or (1,)))
    scale_b = np.ascontiguousarray(np.broadcast_to(np.asarray(scale_b, dtype=np.double).reshape(lead or (1,))
                                                   if np.ndim(scale_b) else np.full(lead or (1,), float(scale_b)),
                                                   lead or (1,)))
    tail = (3,) if n_derivs else ()
    if out is None:
        out = np.empty(lead + tail)
    buf = out if out.flags.c_contiguous and out.dtype == np.double else np.empty(lead + tail)
    check(lib().phb_op_lnl_branch(_DEVICE, S, A, 2 if n_derivs else 0, dptr(probs), dptr(pi), dptr(partials_a),
                                  dptr(partials_b), dptr(scale_a), dptr(scale_b), dptr(buf.reshape(-1))))
    if buf is not out:
        out[...] = buf
    return out if out.ndim else float(out)


def lnl_branch(probs, pi, partials_a, partials_b, scale_a, scale_b, out=None):
    return _branch(probs, pi, partials_a, partials_b, scale_a, scale_b, out, 0)


def lnl_branch_derivs(probs, pi, partials_a, partials_b, scale_a, scale_b, out=None):
    return _branch(probs, pi, partials_a, partials_b, scale_a, scale_b, out, 2)


def transition_matrices(eigen, times, order=0):
    """Batched Eigen.exp / fn_apply on the device: (n, A, A) for scaled times ``times``; order 0/1/2 = P, dP, d2P."""
    evecs, evals, ivecs = _c(eigen.evecs), _c(eigen.evals), _c(eigen.ivecs)
    times = _c(np.atleast_1d(times))
    A = evals.shape[0]
    out = np.empty((times.shape[0], A, A))
    check(lib().phb_op_pmatrices(_DEVICE, A, times.shape[0], dptr(evecs), dptr(evals), dptr(ivecs), dptr(times),
                                 int(order), dptr(out)))
    return out
