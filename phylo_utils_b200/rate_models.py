"""
Among-site rate heterogeneity models: ``.ncat``, ``.rates``, ``.weights``
(reference: /root/reference/phylo_utils/rate_models.py:4-121).  The likelihood kernels take any
number of categories, rates that may be zero (invariant class) and unequal weights.
"""
import numpy as np

from .discrete_gamma import discrete_gamma


class RateModel(object):
    _weights = None
    _rates = None
    ncat = 0

    @property
    def weights(self):
        return self._weights

    @property
    def rates(self):
        return self._rates

    def __str__(self):
        return '{!r}\nweights={}\nrates={}'.format(self, self.weights, self.rates)


def _require_pinvar(pinvar):
    if not 0 <= pinvar < 1:
        raise ValueError("pinvar must be in the range [0, 1)")
    return pinvar


def _require_alpha(alpha):
    if not 0.001 <= alpha:
        raise ValueError("alpha must be greater than 0.001")
    return float(alpha)


class GammaRateModel(RateModel):
    """``ncat`` equiprobable categories, mean rates of Gamma(alpha, alpha) (reference: rate_models.py:15-37)."""

    def __init__(self, ncat, alpha=1.0):
        self.ncat = ncat
        self._weights = np.full(ncat, 1.0 / ncat)
        self.alpha = alpha

    def __repr__(self):
        return "GammaRateModel(ncat={},alpha={})".format(self.ncat, self.alpha)

    @property
    def alpha(self):
        return self._alpha

    @alpha.setter
    def alpha(self, value):
        self._alpha = float(value)
        self._rates = discrete_gamma(self._alpha, self.ncat)


class UniformRateModel(RateModel):
    def __init__(self):
        self.ncat = 1
        self._weights = np.array([1.0])
        self._rates = np.array([1.0])

    def __repr__(self):
        return "UniformRateModel()"


class InvariantSitesModel(RateModel):
    """Two classes: rate 0 with weight pinvar, rate 1/(1-pinvar) otherwise (reference: rate_models.py:50-74)."""

    def __init__(self, pinvar):
        self.ncat = 2
        self.pinvar = pinvar

    def __repr__(self):
        return "InvariantSitesModel(pinvar={})".format(self.pinvar)

    @property
    def pinvar(self):
        return self._pinvar

    @pinvar.setter
    def pinvar(self, value):
        self._pinvar = _require_pinvar(value)
        self._weights = np.array([value, 1 - value])
        self._rates = np.array([0, 1 / (1 - value)])


class InvariantGammaModel(RateModel):
    """+I+G: an invariant class in front of ``n_gamma_cat`` gamma classes (reference: rate_models.py:77-121)."""

    def __init__(self, pinvar, n_gamma_cat, alpha=1.0):
        self._pinvar = _require_pinvar(pinvar)
        self._alpha = _require_alpha(alpha)
        self.ncat = n_gamma_cat + 1
        self._refresh()

    def __repr__(self):
        return "InvariantGammaModel(pinvar={},n_gamma_cat={},alpha={})".format(self._pinvar, self.ncat - 1, self._alpha)

    def _compute_rates_and_weights(self, pinvar, ncat, alpha):
        gamma_rates = discrete_gamma(alpha, ncat)
        rates = np.hstack([0, gamma_rates / (1 - pinvar)])
        weights = np.hstack([pinvar, np.ones(ncat) / ncat * (1 - pinvar)])
        return rates, weights

    def _refresh(self):
        self._rates, self._weights = self._compute_rates_and_weights(self._pinvar, self.ncat - 1, self._alpha)

    @property
    def alpha(self):
        return self._alpha

    @alpha.setter
    def alpha(self, value):
        self._alpha = _require_alpha(value)
        self._refresh()

    @property
    def pinvar(self):
        return self._pinvar

    @pinvar.setter
    def pinvar(self, value):
        self._pinvar = _require_pinvar(value)
        self._refresh()
