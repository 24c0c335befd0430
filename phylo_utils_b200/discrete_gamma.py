"""
Native discrete-gamma rates: ``discrete_gamma(alpha, ncat, median_rates=False)``.

Same call signature and the same numbers as the reference's Cython wrapper around PAML's
DiscreteGamma (/root/reference/src/discrete_gamma.pyx:30-47, src/c_discrete_gamma.c:285-321).
Here the arithmetic is a C++ restatement compiled into libphylo_b200.so
(csrc/discrete_gamma.cpp, entry point ``phb_discrete_gamma``) and reached through ctypes -
host code, no GPU needed.

>>> discrete_gamma(0.5, 5)
array([0.02121238, 0.15548577, 0.46708288, 1.10711735, 3.24910162])
"""
import ctypes

import numpy as np

from ._lib import lib, check


def discrete_gamma(alpha, ncat, median_rates=False):
    rates = np.zeros(int(ncat), dtype=np.double)
    weights = np.zeros(int(ncat), dtype=np.double)
    check(lib().phb_discrete_gamma(ctypes.c_double(float(alpha)), ctypes.c_double(float(alpha)),
                                   ctypes.c_int(int(ncat)), ctypes.c_int(1 if median_rates else 0),
                                   rates.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
                                   weights.ctypes.data_as(ctypes.POINTER(ctypes.c_double))))
    return rates
