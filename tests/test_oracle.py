"""
Pins the CPU oracle (oracle/pruning_oracle.c) against (a) outputs of the unmodified reference stored in
tests/golden/ and (b) the known answers in the reference's own tests.  The GPU parity tests then
compare the CUDA path with this oracle at sizes the golden files do not cover.
"""
import numpy as np
import pytest

import phylo_utils_b200 as phy
from helpers import load, problem, tip_partials, assert_lnl_close, CASES
from oracle import oracle


@pytest.mark.parametrize("name", sorted(CASES))
def test_tree_lnl_matches_reference(name):
    g, tr, codes, lut, sw, ii, names, model, rate = problem(name)
    pattern, ot = oracle.tree_lnl(tr, tip_partials(tr, codes, lut, names), model.p, model.freqs, rate.rates,
                                  rate.weights, return_tree=True)
    assert_lnl_close(pattern[ii], g["site_lnl"], what=name + " per-site lnL")
    assert_lnl_close(pattern[ii].sum(), g["total_lnl"], what=name + " total lnL")
    if "partials" in g:
        # same scaling rule as the reference, so raw partials and scalers must agree, not only their product
        assert np.allclose(ot.partials, g["partials"], rtol=1e-9, atol=1e-300)
        assert np.allclose(ot.scale, g["scale"], rtol=1e-12, atol=1e-12)
    assert np.allclose(ot.root_scale, g["root_scale"], rtol=1e-12, atol=1e-12)


def test_scaling_is_exercised_by_the_deep_cases():
    for name in ("deep300_gtr_g4", "ladder120_k80_g4", "prot150_jtt_g4"):
        g = load(name)
        assert g["root_scale"].min() < -80, name       # at least one rescale below 2^-128


def test_engine_operators_at_61_states():
    g = load("engine_a61")
    sp = np.zeros_like(g["out_scale"])
    out = oracle.clv(g["p1"], g["p2"], g["clv1"], g["clv2"], g["sa"], g["sb"], sp)
    assert np.allclose(out, g["out"], rtol=1e-12, atol=0)
    assert np.allclose(sp, g["out_scale"], rtol=1e-13, atol=1e-13)
    assert (g["out_scale"] != g["sa"] + g["sb"]).any() and (g["out_scale"] == g["sa"] + g["sb"]).any()
    assert np.allclose(oracle.lnl_node(g["pi"], out, sp), g["lnl_node"], rtol=1e-13)


def test_branch_operators():
    g = load("engine_branch")
    d = oracle.lnl_branch_derivs(g["probs"], g["pi"], g["pa"], g["pb"], g["sa"], g["sb"])
    assert np.allclose(d, g["derivs"].reshape(-1, 3), rtol=1e-11, atol=1e-13)
    l0 = oracle.lnl_branch(g["probs"][0], g["pi"], g["pa"], g["pb"], g["sa"], g["sb"])
    assert np.allclose(l0, g["lnl"].reshape(-1), rtol=1e-13)


def test_reference_test_suite_known_answer_k80_pair():
    # /root/reference/tests/test_likelihood.py:30-49 (state order re-mapped)
    k80 = phy.substitution_models.K80(2.)
    c = np.array([[[0., 1., 0., 0.]]])
    t = np.array([[[0., 0., 0., 1.]]])
    sc = np.zeros((1, 1))
    part = oracle.clv(k80.p(0.1)[None], k80.p(0.2)[None], c, t, np.zeros((1, 1)), np.zeros((1, 1)), sc)
    assert sorted(np.round(part.ravel(), 4).tolist(), reverse=True) == [0.0764, 0.0378, 0.0011, 0.0011]
    lnl = oracle.lnl_node(k80.freqs, part, sc)
    assert abs(float(lnl.ravel()[0]) - (-3.5371)) < 5e-5
    g = load("k80_pair")
    assert np.allclose(part, g["partials"], rtol=1e-13) and np.allclose(lnl, g["lnl"], rtol=1e-13)
    # symmetry of the edge likelihood (tests/test_likelihood.py:35-41)
    for n in np.linspace(0.1, 1.0, 10):
        ab = oracle.lnl_branch(k80.p(n), k80.freqs, c[0, 0], t[0, 0], 0.0, 0.0)
        ba = oracle.lnl_branch(k80.p(n), k80.freqs, t[0, 0], c[0, 0], 0.0, 0.0)
        assert abs(float(np.ravel(ab)[0]) - float(np.ravel(ba)[0])) < 1e-14


def test_jc_closed_form_cross_check():
    # two tips, JC69: L = 1/4 (1/4 + 3/4 e^{-4t/3}) for identical states
    jc = phy.substitution_models.JC69()
    a = np.array([[[1., 0., 0., 0.]]])
    sc = np.zeros((1, 1))
    t = 0.3
    part = oracle.clv(jc.p(0.0)[None], jc.p(t)[None], a, a, np.zeros((1, 1)), np.zeros((1, 1)), sc)
    lnl = float(oracle.lnl_node(jc.freqs, part, sc).ravel()[0])
    assert abs(lnl - np.log(0.25 * (0.25 + 0.75 * np.exp(-4 * t / 3)))) < 1e-14


@pytest.mark.reference
def test_oracle_matches_live_reference_on_a_fresh_problem():
    from oracle import ref_shims
    ref = ref_shims.load_reference()
    rng = np.random.default_rng(99)
    tree = phy.tree.random_tree(25, 99)
    names = [l.taxon.label for l in tree.leaf_node_iter()]
    seqs = ["".join(rng.choice(list("ACGT-"), size=300, p=[.24, .24, .24, .24, .04])) for _ in names]
    tm = ref.tree_model.TreeModel()
    tm.set_tree(tree)
    tm.set_alignment([ref_shims.Record(n, s) for n, s in zip(names, seqs)], 0)
    tm.set_rate_model(ref.rate_models.GammaRateModel(4, 0.3))
    tm.set_substitution_model(ref.substitution_models.HKY85(3.0, [0.2, 0.3, 0.3, 0.2]))
    tm.initialise()
    want = tm.compute_likelihood_at_edge(*tm.traversal.root_edge)
    tr = phy.traversal.Traversal(phy.utils.deepcopy_tree(tree))
    codes, lut, sw, ii, nm = phy.alignment.alignment_to_codes([phy.alignment.SeqRecord(n, s) for n, s in zip(names, seqs)], 0)
    model = phy.substitution_models.HKY85(3.0, [0.2, 0.3, 0.3, 0.2])
    rate = phy.rate_models.GammaRateModel(4, 0.3)
    got = oracle.tree_lnl(tr, tip_partials(tr, codes, lut, nm), model.p, model.freqs, rate.rates, rate.weights)
    assert_lnl_close(got[ii], want)
