"""
Edge derivatives and the pre-order pass (BASELINE configs 3 and 5).  The reference only ships the
single-category primitive lnl_branch_derivs; the oracle here is its composition over the Gamma mixture
(helpers.oracle_edge_derivatives), cross-checked by finite differences of the oracle lnL.
"""
import numpy as np
import pytest

import phylo_utils_b200 as phy
from phylo_utils_b200 import _lib
from phylo_utils_b200.tree import random_tree, caterpillar_tree, balanced_tree
from helpers import (problem, records, tree, tip_partials, oracle_up_partials, oracle_edge_derivatives, assert_lnl_close)
from oracle import oracle

pytestmark = pytest.mark.gpu


def setup_case(name, mode):
    g, tr, codes, lut, sw, ii, names, model, rate = problem(name)
    tm = phy.TreeModel(mode=mode, up_partials=True)
    tm.set_tree(tree(g))
    tm.set_alignment(records(g), int(g["alphabet"]))
    tm.set_rate_model(rate)
    tm.set_substitution_model(model)
    tm.initialise()
    tm.compute_up_partials()
    _, ot = oracle.tree_lnl(tr, tip_partials(tr, codes, lut, names), model.p, model.freqs, rate.rates, rate.weights,
                            return_tree=True)
    up = oracle_up_partials(tr, ot, model, rate.rates)
    return tm, tr, ot, up, model, rate, sw


@pytest.mark.parametrize("mode", ["level", "tile", "resident"])
@pytest.mark.parametrize("name", ["cfg1_gtr_g4", "ambig_hky_ig", "prot12_lg_g4", "ladder120_k80_g4"])
def test_all_edges_match_the_composed_oracle(name, mode):
    if mode == "resident" and name in ("ambig_hky_ig", "prot12_lg_g4"):
        pytest.skip("the operand-resident walks cover 4-state models with 1, 2, 4 or 8 categories")
    tm, tr, ot, up, model, rate, sw = setup_case(name, mode)
    nodes = [n for n in range(2 * len(tr.names) - 2) if n != tr.root_edge[1]]
    if len(nodes) > 40:
        nodes = nodes[::7] + list(tr.root_edge[:1])
    lengths = np.array([tm.branch_length_above(n) for n in nodes])
    got = tm.edge_derivatives(nodes, lengths)
    base = tm.lnl()
    for (node, t), row in zip(zip(nodes, lengths), got):
        want = oracle_edge_derivatives(tr, ot, up, model, rate, node, t, sw)
        assert abs(row[0] - want[0]) <= 1e-10 * abs(want[0]), (node, row, want)
        assert abs(row[1] - want[1]) <= 1e-8 * max(1.0, abs(want[1])), (node, row, want)
        assert abs(row[2] - want[2]) <= 1e-8 * max(1.0, abs(want[2])), (node, row, want)
        # the likelihood is the same whichever edge it is evaluated on (pulley principle)
        assert abs(row[0] - base) <= 1e-10 * abs(base)


def test_derivatives_agree_with_finite_differences_of_the_oracle_lnl():
    g, tr, codes, lut, sw, ii, names, model, rate = problem("cfg1_gtr_g4")
    tm, tr, ot, up, model, rate, sw = setup_case("cfg1_gtr_g4", "tile")
    tips = tip_partials(tr, codes, lut, names)
    key = sorted(tr.brlens.keys())[5]
    node = key[0] if tm.traversal.is_leaf(key[0]) or key[0] < key[1] else key[1]
    # node must be the child end of the edge
    rows = tr.postorder_traversal
    child = [c for par, c1, c2 in rows for c in (int(c1), int(c2)) if {int(par), c} == set(key)]
    node = child[0] if child else tr.root_edge[0]
    t0 = tr.brlens[key]
    h = 1e-4

    def lnl_at(t):
        tr.brlens[tr.brlens.canonical_key(key)] = t
        pat = oracle.tree_lnl(tr, tips, model.p, model.freqs, rate.rates, rate.weights)
        return float(np.dot(pat, sw))
    f_plus, f_0, f_minus = lnl_at(t0 + h), lnl_at(t0), lnl_at(t0 - h)
    tr.brlens[tr.brlens.canonical_key(key)] = t0
    got = tm.edge_derivatives([node], [t0])[0]
    assert abs(got[0] - f_0) <= 1e-10 * abs(f_0)
    assert abs(got[1] - (f_plus - f_minus) / (2 * h)) < 1e-4 * max(1.0, abs(got[1]))
    assert abs(got[2] - (f_plus - 2 * f_0 + f_minus) / (h * h)) < 1e-2 * max(1.0, abs(got[2]))


def test_trial_lengths_and_reference_convention_flag():
    tm, tr, ot, up, model, rate, sw = setup_case("cfg1_gtr_g4", "level")
    node = int(tr.postorder_traversal[0][1])
    for t in (0.01, 0.2, 1.5):
        for chain in (True, False):
            got = tm.edge_derivatives([node], [t], chain_rule=chain)[0]
            want = oracle_edge_derivatives(tr, ot, up, model, rate, node, t, sw, chain_rule=chain)
            assert np.allclose(got, want, rtol=1e-8, atol=1e-8)


@pytest.mark.parametrize("name,mode", [("prot12_lg_g4", "tile"), ("prot12_lg_g4", "level"), ("cfg1_gtr_g4", "resident"),
                                       ("ladder120_k80_g4", "resident")])
def test_repeated_passes_read_the_sum_tables(name, mode):
    """Newton iterations evaluate the same edges again at new trial lengths: from the second pass on the kernels read
    the per-edge sum tables the first pass (20 / 61 states) or the pre-order walk (4 states) left in the up blocks."""
    tm, tr, ot, up, model, rate, sw = setup_case(name, mode)
    nodes = [n for n in range(2 * len(tr.names) - 2) if n != tr.root_edge[1]]
    if len(nodes) > 40:
        nodes = nodes[::9] + list(tr.root_edge[:1])
    lengths = np.array([tm.branch_length_above(n) for n in nodes])
    first = tm.edge_derivatives(nodes, lengths)
    again = tm.edge_derivatives(nodes, lengths)
    assert np.allclose(first, again, rtol=1e-12, atol=1e-9)
    for scale in (0.5, 3.0):
        got = tm.edge_derivatives(nodes, lengths * scale)
        for (node, t), row in zip(zip(nodes, lengths * scale), got):
            want = oracle_edge_derivatives(tr, ot, up, model, rate, node, t, sw)
            assert abs(row[0] - want[0]) <= 1e-10 * abs(want[0]), (node, row, want)
            assert abs(row[1] - want[1]) <= 1e-8 * max(1.0, abs(want[1])), (node, row, want)
            assert abs(row[2] - want[2]) <= 1e-8 * max(1.0, abs(want[2])), (node, row, want)
    # a subset, with a node listed twice (no table is written by such a launch, existing ones are still read)
    sub = [nodes[0], nodes[1], nodes[0]]
    got = tm.edge_derivatives(sub, lengths[[0, 1, 0]])
    assert np.allclose(got, first[[0, 1, 0]], rtol=1e-12, atol=1e-9)
    # a new pre-order pass starts over; a subset first, then everything: one launch mixes edges that already have their
    # table with edges that do not
    tm.compute_up_partials()
    half = list(range(0, len(nodes), 2))
    got = tm.edge_derivatives([nodes[i] for i in half], lengths[half])
    assert np.allclose(got, first[half], rtol=1e-12, atol=1e-9)
    assert np.allclose(tm.edge_derivatives(nodes, lengths), first, rtol=1e-12, atol=1e-9)
    assert np.allclose(tm.edge_derivatives(nodes, lengths), first, rtol=1e-12, atol=1e-9)


def test_up_partials_larger_tree_vs_oracle_total():
    rng = np.random.default_rng(12)
    n_taxa, n_pat = 80, 20000
    tr_tree = random_tree(n_taxa, 12)
    names = [l.taxon.label for l in tr_tree.leaf_node_iter()]
    lut = np.vstack([np.eye(4)[::-1], np.ones((1, 4))])
    codes = rng.integers(0, 5, size=(n_taxa, n_pat)).astype(np.uint8)
    model = phy.substitution_models.GTR([6., 5., 4., 3., 2., 1.], [0.1, 0.2, 0.3, 0.4])
    rate = phy.rate_models.GammaRateModel(4, 0.5)
    tm = phy.TreeModel(mode="tile", up_partials=True)
    tm.set_tree(tr_tree)
    tm.set_tip_codes(codes, lut, {n: i for i, n in enumerate(names)})
    tm.set_rate_model(rate)
    tm.set_substitution_model(model)
    tm.initialise()
    tm.compute_up_partials()
    base = tm.lnl()
    nodes = np.arange(2 * n_taxa - 2)
    out = tm.edge_derivatives(nodes)
    assert np.all(np.abs(out[:, 0] - base) <= 1e-10 * abs(base))          # every edge reproduces the same lnL
    assert np.all(np.isfinite(out))


def _walk_vs_two_rows(tree_fn, n_taxa, n_pat, ppt, monkeypatch, n_cat=4, iupac=False):
    rng = np.random.default_rng(n_taxa)
    tr_tree = tree_fn(n_taxa, 7)
    names = [l.taxon.label for l in tr_tree.leaf_node_iter()]
    if iupac:
        # all 15 non-empty state sets: look-up tables of more than 8 rows take the 16-row tip tables
        lut = np.array([[(c >> (3 - i)) & 1 for i in range(4)] for c in range(1, 16)], dtype=float)
        codes = rng.integers(0, 15, size=(n_taxa, n_pat)).astype(np.uint8)
    else:
        lut = np.vstack([np.eye(4)[::-1], np.ones((1, 4))])
        codes = rng.integers(0, 5, size=(n_taxa, n_pat)).astype(np.uint8)
    model = phy.substitution_models.GTR([6., 5., 4., 3., 2., 1.], [0.1, 0.2, 0.3, 0.4])
    rate = phy.rate_models.GammaRateModel(n_cat, 0.5) if n_cat > 1 else phy.rate_models.UniformRateModel()
    weights = rng.integers(1, 5, size=n_pat)
    out = {}
    for label, env in (("walk", None), ("plain_walk", "plain"), ("two_rows", "1")):
        monkeypatch.setenv("PHB_UP_PPT", ppt)
        monkeypatch.delenv("PHB_UP_TWO_ROWS", raising=False)
        monkeypatch.delenv("PHB_UP_PLAIN", raising=False)
        if env == "1":
            monkeypatch.setenv("PHB_UP_TWO_ROWS", "1")
        elif env == "plain":
            monkeypatch.setenv("PHB_UP_PLAIN", "1")       # the walk storing up partials instead of sum tables
        _lib.lib().phb_reload_tuning()
        tm = phy.TreeModel(mode="resident", up_partials=True)
        tm.set_tree(tr_tree)
        tm.set_tip_codes(codes, lut, {n: i for i, n in enumerate(names)}, siteweights=weights)
        tm.set_rate_model(rate)
        tm.set_substitution_model(model)
        tm.initialise()
        tm.compute_up_partials()
        a, b = tm.traversal.root_edge
        nodes = np.asarray([n for n in range(2 * n_taxa - 2) if n != b])
        first = tm.edge_derivatives(nodes)
        second = tm.edge_derivatives(nodes, tm.lengths_above(nodes) * 1.7)
        out[label] = (first, second, tm.lnl())
    walk, walk2, base = out["walk"]
    assert np.all(np.isfinite(walk))
    assert np.all(np.abs(walk[:, 0] - base) <= 1e-10 * abs(base))          # pulley principle on every edge
    for other in ("plain_walk", "two_rows"):
        assert np.allclose(walk, out[other][0], rtol=1e-9, atol=1e-7), other
        assert np.allclose(walk2, out[other][1], rtol=1e-9, atol=1e-7), other


@pytest.mark.parametrize("tree_fn,n_taxa,n_pat", [(random_tree, 150, 20001), (caterpillar_tree, 200, 4099), (random_tree, 3, 70),
                                                  (random_tree, 4, 33), (balanced_tree, 512, 1500)])   # balanced: deepest parking
@pytest.mark.parametrize("ppt", ["1", "2"])
def test_pre_order_walk_matches_the_two_row_form(tree_fn, n_taxa, n_pat, ppt, monkeypatch):
    """up_dna_pair.cu (sum-table form and plain form) against the two-rows-per-parent pass on the same device partials:
    every edge's (lnL, dlnL, d2lnL) agrees to rounding at the current and at other trial lengths, ragged last tile and
    deep scaling (caterpillar) included."""
    _walk_vs_two_rows(tree_fn, n_taxa, n_pat, ppt, monkeypatch)


@pytest.mark.parametrize("n_cat,iupac", [(1, False), (2, False), (2, True), (4, True), (8, False)])   # 8: the walk declines, two rows per parent take over
def test_pre_order_walk_other_category_counts_and_iupac_codes(n_cat, iupac, monkeypatch):
    _walk_vs_two_rows(random_tree, 60, 5003, "1", monkeypatch, n_cat=n_cat, iupac=iupac)


def test_derivatives_need_the_up_pass():
    g, tr, codes, lut, sw, ii, names, model, rate = problem("cfg1_gtr_g4")
    tm = phy.TreeModel()
    tm.set_tree(tree(g))
    tm.set_alignment(records(g), 0)
    tm.set_rate_model(rate)
    tm.set_substitution_model(model)
    tm.initialise()
    with pytest.raises(RuntimeError):
        tm.compute_up_partials()                   # context built without up_partials=True


def test_newton_sweeps_increase_lnl_monotonically_and_agree_with_the_oracle():
    from phylo_utils_b200.optimise import optimise_branch_lengths, edge_nodes
    g, tr, codes, lut, sw, ii, names, model, rate = problem("cfg1_gtr_g4")
    tm = phy.TreeModel(up_partials=True)
    tm.set_tree(tree(g))
    tm.set_alignment(records(g), 0)
    tm.set_rate_model(rate)
    tm.set_substitution_model(model)
    tm.initialise()
    start = tm.lnl()
    res = optimise_branch_lengths(tm, max_sweeps=25, tol=1e-7)
    trace = np.asarray(res["trace"])
    assert np.all(np.diff(trace) >= 0) and trace[0] == start
    assert res["lnl"] > start + 1.0                      # random data on a random tree: far from the optimum at the start
    # the oracle evaluated at the final branch lengths agrees with the device value
    want = oracle.tree_lnl(tm.traversal, tip_partials(tm.traversal, codes, lut, names), model.p, model.freqs, rate.rates,
                           rate.weights)
    assert_lnl_close(res["lnl"], float(np.dot(want, sw)))
    # stationary point: gradients are small relative to the start
    tm.compute_up_partials()
    d = tm.edge_derivatives(edge_nodes(tm.traversal))
    interior = res["lengths"] > 2e-5
    assert np.abs(d[interior, 1]).max() < 1e-2 * 3112.0
