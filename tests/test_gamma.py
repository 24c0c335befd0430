"""Discrete-gamma rates and rate models (reference: src/discrete_gamma.pyx, gamma.py, rate_models.py)."""
import numpy as np
import pytest

import phylo_utils_b200 as phy
from helpers import load
from oracle import oracle


def test_docstring_known_answer():
    # /root/reference/src/discrete_gamma.pyx:41-42
    want = [0.02121238, 0.15548577, 0.46708288, 1.10711735, 3.24910162]
    assert np.allclose(phy.discrete_gamma.discrete_gamma(0.5, 5), want, atol=5e-9)
    assert np.allclose(phy.gamma.discrete_gamma(5, 0.5), want, atol=5e-9)


def test_native_matches_reference_bit_for_bit():
    g = load("gamma")
    for a in g["alphas"]:
        for k in g["ncats"]:
            mine = phy.discrete_gamma.discrete_gamma(float(a), int(k))
            assert np.array_equal(mine, g["native_a{}_k{}".format(a, k)]), (a, k)
            med = phy.discrete_gamma.discrete_gamma(float(a), int(k), True)
            assert np.array_equal(med, g["median_a{}_k{}".format(a, k)]), (a, k)


def test_scipy_flavour_matches_reference():
    g = load("gamma")
    for a in g["alphas"]:
        for k in g["ncats"]:
            mine = phy.gamma.discrete_gamma(int(k), float(a))
            assert np.allclose(mine, g["scipy_a{}_k{}".format(a, k)], rtol=1e-13, atol=0)


@pytest.mark.skipif(not oracle.have_ref_gamma(), reason="oracle/_ref not built")
def test_native_matches_compiled_reference_c_on_a_grid():
    for a in [0.02, 0.11, 0.37, 0.5, 0.93, 1.0, 1.7, 2.0, 4.4, 9.0, 10.0, 33.3, 120.0]:
        for k in [2, 3, 4, 6, 10]:
            for median in (False, True):
                assert np.array_equal(phy.discrete_gamma.discrete_gamma(a, k, median),
                                      oracle.ref_discrete_gamma(a, k, median)), (a, k, median)


def test_rates_have_unit_mean_and_increase():
    for a in [0.1, 0.5, 1.0, 3.0]:
        r = phy.discrete_gamma.discrete_gamma(a, 4)
        assert np.all(np.diff(r) > 0)
        assert abs(r.mean() - 1.0) < 1e-6


def test_rate_models():
    rm = phy.rate_models
    g = rm.GammaRateModel(4, 0.5)
    assert g.ncat == 4 and np.allclose(g.weights, 0.25)
    assert np.array_equal(g.rates, phy.discrete_gamma.discrete_gamma(0.5, 4))
    g.alpha = 2.0
    assert np.array_equal(g.rates, phy.discrete_gamma.discrete_gamma(2.0, 4))
    u = rm.UniformRateModel()
    assert u.ncat == 1 and u.rates[0] == 1.0 and u.weights[0] == 1.0
    inv = rm.InvariantSitesModel(0.25)
    assert inv.ncat == 2 and np.allclose(inv.rates, [0, 1 / 0.75]) and np.allclose(inv.weights, [0.25, 0.75])
    assert abs(np.dot(inv.rates, inv.weights) - 1.0) < 1e-15
    ig = rm.InvariantGammaModel(0.2, 4, 0.7)
    assert ig.ncat == 5 and ig.rates[0] == 0 and abs(np.dot(ig.rates, ig.weights) - 1.0) < 1e-6
    with pytest.raises(ValueError):
        rm.InvariantSitesModel(1.0)
    with pytest.raises(ValueError):
        rm.InvariantGammaModel(0.2, 4, 0.0001)
    with pytest.raises(ValueError):
        ig.pinvar = -0.1


def test_invalid_gamma_arguments():
    with pytest.raises(ValueError):
        phy.discrete_gamma.discrete_gamma(-1.0, 4)
    with pytest.raises(ValueError):
        phy.discrete_gamma.discrete_gamma(0.5, 0)
