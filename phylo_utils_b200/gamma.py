"""
Discrete-gamma category rates, scipy formulation (reference: /root/reference/phylo_utils/gamma.py:4-18).

NOTE the argument order ``(ncat, alpha)`` - the native routine in
:mod:`phylo_utils_b200.discrete_gamma` takes ``(alpha, ncat)``, exactly as the two
reference functions differ.  The two agree to 1e-10...1e-8 relative only (the PAML series
stops at 1e-8), so always feed both sides of a comparison the *same* rates array.
"""
import numpy as np
from scipy.special import gammaincinv, gammainc


def discrete_gamma(ncat, alpha, beta=None):
    """Mean rate of each of ``ncat`` equiprobable categories of Gamma(alpha, beta=alpha)."""
    if beta is None:
        beta = alpha
    edges = np.arange(ncat + 1, dtype=np.double) / ncat
    cut = gammaincinv(alpha, edges)                 # quantiles times beta
    mass = gammainc(alpha + 1.0, cut)               # integral of x*pdf up to each cut, over the mean
    return ncat * (alpha / beta) * np.diff(mass)
