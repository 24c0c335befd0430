"""
Scalar optimisation helpers under the names of the reference's compiled ``phylo_utils.optimisation`` module
(/root/reference/src/optimisation.pyx): the simplex <-> unconstrained parameter transforms (:14-50), quadratic
interpolation (:62-84) and the one-dimensional minimisers ``brent_wrap`` / ``dbrent_wrap`` (:86-313), plus what the
GPU path needs on top: ``maximise_bracketed`` - a BATCHED bracketing search that drives the maxima of many independent
one-dimensional functions at once, each evaluation being one device launch over all of them
(``TreeModel.edge_derivatives``).  It is the safeguard behind the Newton sweeps of :mod:`phylo_utils_b200.optimise`
for edges whose likelihood curve is not concave where Newton stands.

The minimisers are the textbook algorithms (Brent 1973; Press et al., "Numerical Recipes", section 10.2-10.3) written
for this package; results are returned like the reference's wrappers: ``array([x_min, f(x_min), iterations])``.
"""
import numpy as np
from scipy.special import expit, logit

__all__ = ["simplex_encode", "simplex_decode", "transform_params", "decode_params", "quad_interp", "brent_wrap",
           "dbrent_wrap", "maximise_bracketed"]

ITMAX = 100
CGOLD = 0.3819660112501051
ZEPS = 1.0e-10
TINY = 1e-15


# ---- parameter transforms (optimisation.pyx:14-50) ---------------------------------------------------------------
def simplex_encode(p):
    """p (length N, sums to 1) -> theta (length N-1, each in (0, 1)): stick-breaking fractions."""
    p = np.asarray(p, dtype=np.double)
    remaining = 1.0 - np.concatenate([[0.0], np.cumsum(p[:-2])])
    return p[:-1] / remaining


def simplex_decode(theta):
    theta = np.asarray(theta, dtype=np.double)
    stick = np.concatenate([[1.0], np.cumprod(1.0 - theta)])
    return np.concatenate([theta * stick[:-1], stick[-1:]])


def transform_params(p):
    return logit(simplex_encode(p))


def decode_params(q):
    return simplex_decode(expit(q))


# ---- one-dimensional minimisation ----------------------------------------------------------------------------------
def quad_interp(p, q, r, fp, fq, fr):
    """Abscissa of the turning point of the parabola through (p, fp), (q, fq), (r, fr) (optimisation.pyx:62-84)."""
    num = (q * q - r * r) * fp + (r * r - p * p) * fq + (p * p - q * q) * fr
    div = (q - r) * fp + (r - p) * fq + (p - q) * fr
    if abs(div) < TINY:
        div = -TINY if div < 0 else TINY
    return num / (2.0 * div)


def _result(x, fx, it):
    return np.array([x, fx, float(it)])


def brent_wrap(guess, lbracket, rbracket, fn, tol=1.5e-8):
    """Brent's derivative-free minimiser on [lbracket, rbracket] started at ``guess`` (optimisation.pyx:86-177, :308-313)."""
    a, b = (lbracket, rbracket) if lbracket < rbracket else (rbracket, lbracket)
    x = w = v = guess
    fx = fw = fv = fn(x)
    step = prev_step = 0.0
    for it in range(1, ITMAX + 1):
        mid = 0.5 * (a + b)
        tol1 = tol * abs(x) + ZEPS
        tol2 = 2.0 * tol1
        if abs(x - mid) <= tol2 - 0.5 * (b - a):
            return _result(x, fx, it)
        golden = True
        if abs(prev_step) > tol1:                       # try the parabola through x, w, v
            r = (x - w) * (fx - fv)
            q = (x - v) * (fx - fw)
            p = (x - v) * q - (x - w) * r
            q = 2.0 * (q - r)
            if q > 0.0:
                p = -p
            q = abs(q)
            if abs(p) < abs(0.5 * q * prev_step) and q * (a - x) < p < q * (b - x):
                prev_step, step = step, p / q
                u = x + step
                if u - a < tol2 or b - u < tol2:
                    step = tol1 if mid >= x else -tol1
                golden = False
        if golden:
            prev_step = (a - x) if x >= mid else (b - x)
            step = CGOLD * prev_step
        u = x + step if abs(step) >= tol1 else x + (tol1 if step >= 0 else -tol1)
        fu = fn(u)
        if fu <= fx:
            if u >= x:
                a = x
            else:
                b = x
            v, fv, w, fw, x, fx = w, fw, x, fx, u, fu
        else:
            if u < x:
                a = u
            else:
                b = u
            if fu <= fw or w == x:
                v, fv, w, fw = w, fw, u, fu
            elif fu <= fv or v == x or v == w:
                v, fv = u, fu
    return _result(x, fx, ITMAX + 1)


def dbrent_wrap(guess, lbracket, rbracket, fn, dfn, tol=1.5e-8):
    """Brent's minimiser with first derivatives: secant steps on f' inside the bracket, bisection otherwise
    (optimisation.pyx:179-306)."""
    a, b = (lbracket, rbracket) if lbracket < rbracket else (rbracket, lbracket)
    x = w = v = guess
    fx = fw = fv = fn(x)
    dx = dw = dv = dfn(x)
    step = prev_step = 0.0
    for it in range(1, ITMAX):
        mid = 0.5 * (a + b)
        tol1 = tol * abs(x) + ZEPS
        tol2 = 2.0 * tol1
        if abs(x - mid) <= tol2 - 0.5 * (b - a):
            return _result(x, fx, it)
        bisect = True
        if abs(prev_step) > tol1:
            candidates = []
            for other, d_other in ((w, dw), (v, dv)):   # secant through x and each of the two older points
                if d_other != dx:
                    d = (other - x) * dx / (dx - d_other)
                    u = x + d
                    if (a - u) * (u - b) > 0.0 and dx * d <= 0.0:   # inside the bracket, downhill side
                        candidates.append(d)
            before_last, prev_step = prev_step, step
            if candidates:
                d = min(candidates, key=abs)
                if abs(d) <= abs(0.5 * before_last):
                    step = d
                    u = x + step
                    if u - a < tol2 or b - u < tol2:
                        step = tol1 if mid >= x else -tol1
                    bisect = False
        if bisect:
            prev_step = (a - x) if dx >= 0.0 else (b - x)
            step = 0.5 * prev_step
        if abs(step) >= tol1:
            u = x + step
            fu = fn(u)
        else:
            u = x + (tol1 if step >= 0 else -tol1)
            fu = fn(u)
            if fu > fx:                                  # the smallest step downhill goes uphill: done
                return _result(x, fx, it)
        du = dfn(u)
        if fu <= fx:
            if u >= x:
                a = x
            else:
                b = x
            v, fv, dv, w, fw, dw, x, fx, dx = w, fw, dw, x, fx, dx, u, fu, du
        else:
            if u < x:
                a = u
            else:
                b = u
            if fu <= fw or w == x:
                v, fv, dv, w, fw, dw = w, fw, dw, u, fu, du
            elif fu < fv or v == x or v == w:
                v, fv, dv = u, fu, du
    return _result(x, fn(x), ITMAX + 1)


# ---- batched bracketing search (the GPU path's safeguard) -----------------------------------------------------------
def maximise_bracketed(fn, lo, hi, start, tol=1e-8, max_iter=64):
    """
    Maxima of ``n`` independent functions on [lo, hi], all advanced together (each is taken to have one interior
    maximum at most on the side of the starting point its derivative points to - true of a branch's likelihood curve).

    ``fn(t, idx)`` evaluates functions ``idx`` (an index array) at the abscissas ``t`` (same length) and returns an
    ``(len(idx), >= 2)`` array whose columns 0 and 1 are the value and the first derivative - one call is one device
    launch over all listed functions.  Every function keeps a bracket [a, b] with f'(a) > 0 > f'(b); the next abscissa
    is the secant root of f' over the bracket (Illinois variant: the weight of an end that has not moved twice in a row
    is halved), replaced by the (geometric, on wide brackets) midpoint when it falls within 1 % of an end.  Functions whose derivative does not change
    sign on [lo, hi] end on the boundary it points to.  Returns (x, f(x), f'(x), evaluations).
    """
    start = np.clip(np.asarray(start, dtype=np.double), lo, hi)
    n = start.shape[0]
    everyone = np.arange(n)
    a, b = np.full(n, float(lo)), np.full(n, float(hi))
    # both ends and the starting point in ONE launch (every function listed three times)
    r0 = fn(np.concatenate([a, b, start]), np.concatenate([everyone, everyone, everyone]))
    fa, da, fb, db = r0[:n, 0].copy(), r0[:n, 1].copy(), r0[n:2 * n, 0].copy(), r0[n:2 * n, 1].copy()
    x, fx, dx = start.copy(), r0[2 * n:, 0].copy(), r0[2 * n:, 1].copy()
    evaluations = 1
    # the starting point becomes the end of the bracket its derivative allows
    climbing = dx > 0
    a[climbing], fa[climbing], da[climbing] = start[climbing], fx[climbing], dx[climbing]
    b[~climbing], fb[~climbing], db[~climbing] = start[~climbing], fx[~climbing], dx[~climbing]
    low_end = da <= 0                       # falling all the way from the lower bound to the start
    high_end = ~low_end & (db >= 0)         # still climbing at the upper bound
    take_lo = low_end & (fa >= fx)          # (falling at both ends but higher at the start: stay at the start)
    take_hi = high_end & (fb >= fx)
    x[take_lo], fx[take_lo], dx[take_lo] = a[take_lo], fa[take_lo], da[take_lo]
    x[take_hi], fx[take_hi], dx[take_hi] = b[take_hi], fb[take_hi], db[take_hi]
    active = ~(low_end | high_end)
    stuck = np.zeros(n, dtype=np.int8)      # +k: end a has not moved for k steps, -k: end b
    with np.errstate(divide="ignore", invalid="ignore"):
        trial = (a * db - b * da) / (db - da)
    off = ~np.isfinite(trial) | (trial <= a) | (trial >= b)
    trial[off] = np.where((a > 0) & (b > 16.0 * a), np.sqrt(np.abs(a * b)), 0.5 * (a + b))[off]
    for _ in range(max_iter):
        idx = np.flatnonzero(active)
        if idx.size == 0:
            break
        r = fn(trial[idx], idx)
        evaluations += 1
        f, d = r[:, 0], r[:, 1]
        better = f >= fx[idx]
        upd = idx[better]
        x[upd], fx[upd], dx[upd] = trial[upd], f[better], d[better]
        up = d > 0                           # still climbing: the maximum lies to the right
        ia, ib = idx[up], idx[~up]
        a[ia], da[ia] = trial[ia], d[up]
        b[ib], db[ib] = trial[ib], d[~up]
        stuck[ia] = np.minimum(stuck[ia], 0) - 1
        stuck[ib] = np.maximum(stuck[ib], 0) + 1
        wa = np.where(stuck[idx] >= 2, 0.5 ** (stuck[idx] - 1), 1.0) * da[idx]
        wb = np.where(stuck[idx] <= -2, 0.5 ** (-stuck[idx] - 1), 1.0) * db[idx]
        width = b[idx] - a[idx]
        with np.errstate(divide="ignore", invalid="ignore"):
            nxt = (a[idx] * wb - b[idx] * wa) / (wb - wa)
        bad = ~np.isfinite(nxt) | (nxt < a[idx] + 0.01 * width) | (nxt > b[idx] - 0.01 * width)
        # midpoint instead - geometric while the bracket still spans more than a factor of 16 (branch lengths live on
        # a logarithmic scale: [1e-5, 20] is 21 octaves)
        wide = bad & (a[idx] > 0) & (b[idx] > 16.0 * a[idx])
        nxt[bad] = 0.5 * (a[idx] + b[idx])[bad]
        nxt[wide] = np.sqrt(a[idx] * b[idx])[wide]
        # done: bracket or step below the tolerance, or a stationary point hit exactly
        finished = (width <= tol * np.abs(x[idx]) + ZEPS) | (np.abs(nxt - trial[idx]) <= tol * np.abs(nxt) + ZEPS) | (d == 0)
        trial[idx] = nxt
        active[idx[finished]] = False
    return x, fx, dx, evaluations
