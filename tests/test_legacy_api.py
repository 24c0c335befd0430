"""
The legacy object API (phylo_utils_b200.likelihood.legacy = the reference's likelihood.py:9-299, whose own import of
its `likcalc` engine is commented out).

  * `gpu` tests: the assertions of the reference's tests/test_likelihood.py, restated against the CUDA operators, plus
    LnlModel / GammaMixture over a whole tree against the output of the unmodified reference TreeModel (tests/golden/);
  * `reference` test (runs where the reference checkout is mounted - no GPU there): the reference's OWN test file,
    unmodified, driving the same objects with the CPU oracle operators plugged in and the documented legacy state
    order (T, C, A, G) switched on.
"""
import importlib
import os
import sys
import types
import unittest

import numpy as np
import pytest

import phylo_utils_b200 as phy
from phylo_utils_b200.likelihood import legacy
from helpers import load, records, tree

K80_ANSWER = [0.0764, 0.0378, 0.0011, 0.0011]     # tests/test_likelihood.py:32-34
K80_LNL = -3.5371                                  # tests/test_likelihood.py:47-49


@pytest.fixture
def legacy_state_order():
    legacy.STATE_ORDER = "TCAG"
    yield
    legacy.STATE_ORDER = None


def _three_nodes():
    model = phy.substitution_models.K80(2.)
    root, left, right = (phy.likelihood.LnlNode(model) for _ in range(3))
    left.set_partials(np.array([[1, 0, 0, 0]], dtype=np.double))
    right.set_partials(np.array([[0, 1, 0, 0]], dtype=np.double))
    return model, root, left, right


@pytest.mark.gpu
def test_lnl_node_known_answers_of_the_reference_suite(legacy_state_order):
    model, root, left, right = _three_nodes()
    root.update_transition_probabilities(0.1, 0.2)
    assert np.allclose(root.probs1, model.p(0.1)) and np.allclose(root.probs2, model.p(0.2))
    root.set_partials([1, 0, 0, 0])
    assert root.partials.dtype == np.double and root.partials.shape == (1, 4)
    root.compute_partials(left, right)
    assert np.allclose([K80_ANSWER], root.partials.round(4))
    assert abs(np.log((model.freqs * root.partials).sum()) - K80_LNL) < 5e-5
    assert abs(left.compute_likelihood(right, 0.3) - K80_LNL) < 5e-5
    for n in np.linspace(0.1, 1.0, 10):                                       # symmetry across the edge
        left.compute_edge_sitewise_likelihood(right, n)
        right.compute_edge_sitewise_likelihood(left, n)
        assert np.allclose(left.sitewise, right.sitewise, rtol=1e-14, atol=0)
    root.compute_partials(left, right, scale=False)
    assert np.allclose([K80_ANSWER], root.partials.round(4))


@pytest.mark.gpu
def test_pairwise_distance_optimisers_find_the_same_maximum():
    rng = np.random.default_rng(3)
    k80 = phy.substitution_models.K80(0.7345)
    a = rng.integers(0, 4, 1500)
    b = np.where(rng.random(1500) < 0.35, rng.integers(0, 4, 1500), a)
    sites_a, sites_b = np.eye(4)[a], np.eye(4)[b]
    dist, var = phy.likelihood.optimise(k80, sites_a, sites_b, verbose=False)
    root = phy.likelihood.LnlNode(k80)
    root.set_partials(sites_a)
    dist2, var2 = phy.likelihood.brent_optimise(root, phy.likelihood.Leaf(sites_b), verbose=False)
    assert 0.1 < dist < 2.0 and abs(dist - dist2) < 1e-5 * dist and var > 0
    wrapper = phy.likelihood.OptWrapper(k80, sites_a, sites_b, dist)
    assert abs(wrapper.dlnl) < 1e-5 and wrapper.d2lnl < 0
    h = 1e-5                                                                   # derivative columns are true derivatives in t
    lp = root.compute_likelihood(phy.likelihood.Leaf(sites_b), dist + h)
    lm = root.compute_likelihood(phy.likelihood.Leaf(sites_b), dist - h)
    l0, d1, d2 = root.compute_likelihood(phy.likelihood.Leaf(sites_b), dist, derivatives=True)
    assert abs((lp - lm) / (2 * h) - d1) < 1e-4 and abs((lp - 2 * l0 + lm) / h ** 2 - d2) < 1e-2 * abs(d2)


@pytest.mark.gpu
def test_gamma_mixture_over_a_tree_matches_the_reference_tree_model():
    g = load("cfg1_gtr_g4")
    model = phy.substitution_models.GTR([6., 5., 4., 3., 2., 1.], [0.1, 0.2, 0.3, 0.4])
    recs = records(g)
    partials = {r.name: phy.seq_to_partials(str(r.seq), int(g["alphabet"])) for r in recs}
    mix = phy.likelihood.GammaMixture(0.5, 4)
    mix.init_models(model, partials)
    mix.set_tree(str(g["newick"]))
    mix.run()
    site = mix.mix_likelihoods(mix.get_sitewise_likelihoods())[:, 0]
    full_site = np.asarray(g["site_lnl"])                                      # per original site, from the numba reference
    assert site.shape == full_site.shape
    assert np.allclose(site, full_site, rtol=1e-10, atol=0)
    assert abs(mix.get_likelihood() - float(g["total_lnl"])) <= 1e-10 * abs(float(g["total_lnl"]))


class _OracleOperators(object):
    """clv / lnl_branch / lnl_branch_derivs of the CPU oracle under the operator module's names."""

    def __init__(self):
        from oracle import oracle
        self.clv, self.lnl_branch, self.lnl_branch_derivs = oracle.clv, oracle.lnl_branch, oracle.lnl_branch_derivs


@pytest.mark.reference
def test_the_reference_test_file_runs_unmodified_against_the_legacy_objects(legacy_state_order):
    path = "/root/reference/tests/test_likelihood.py"
    saved = {k: v for k, v in sys.modules.items() if k == "phylo_utils" or k.startswith("phylo_utils.")}
    for k in saved:
        del sys.modules[k]
    previous = legacy.use_operators(_OracleOperators())
    try:
        alias = types.ModuleType("phylo_utils")                  # `import phylo_utils as phy` inside the reference test
        alias.likelihood = legacy
        alias.alignment = phy.alignment
        alias.substitution_models = phy.substitution_models
        sys.modules["phylo_utils"] = alias
        sys.modules["phylo_utils.alignment"] = phy.alignment
        sys.modules["phylo_utils.alignment.alignment"] = phy.alignment.alignment
        sys.modules["phylo_utils.substitution_models"] = phy.substitution_models
        sys.modules["phylo_utils.substitution_models.k80"] = importlib.import_module("phylo_utils_b200.substitution_models.k80")
        spec = importlib.util.spec_from_file_location("reference_test_likelihood", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        suite = unittest.defaultTestLoader.loadTestsFromModule(mod)
        result = unittest.TextTestRunner(stream=open(os.devnull, "w")).run(suite)
        assert result.testsRun == 7
        assert not result.errors and not result.failures, (result.errors + result.failures)[0][1]
    finally:
        legacy.use_operators(previous)
        for k in [k for k in sys.modules if k == "phylo_utils" or k.startswith("phylo_utils.")]:
            del sys.modules[k]
        sys.modules.update(saved)
