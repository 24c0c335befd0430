"""
GPU parity: the CUDA path (through the C ABI) against the reference's own outputs (tests/golden/) and
against the CPU oracle on seeded inputs.  Tolerance: 1e-10 relative on total and per-site lnL
(BASELINE.json north_star); integer / index outputs bit-exact.
"""
import numpy as np
import pytest

import phylo_utils_b200 as phy
from phylo_utils_b200 import _lib
from phylo_utils_b200.alignment.alignment import SeqRecord
from phylo_utils_b200.likelihood import clv, lnl_node, lnl_branch, lnl_branch_derivs
from phylo_utils_b200.likelihood.cuda_likelihood_engine import transition_matrices
from helpers import load, problem, records, tree, tip_partials, assert_lnl_close, CASES, ASC_CASES, RTOL
from oracle import oracle

pytestmark = pytest.mark.gpu


def make_tm(name, mode="auto", up=False):
    g, tr, codes, lut, sw, ii, names, model, rate = problem(name)
    tm = phy.TreeModel(mode=mode, up_partials=up)
    tm.set_tree(tree(g))
    tm.set_alignment(records(g), int(g["alphabet"]))
    tm.set_rate_model(rate)
    tm.set_substitution_model(model)
    return g, tm


@pytest.mark.parametrize("mode", ["level", "tile"])
@pytest.mark.parametrize("name", sorted(CASES))
def test_site_and_total_lnl_match_reference(name, mode):
    g, tm = make_tm(name, mode)
    tm.initialise()
    site = tm.compute_likelihood_at_edge(*tm.traversal.root_edge)
    assert_lnl_close(site, g["site_lnl"], what=name + " per-site lnL")
    assert_lnl_close(site.sum(), g["total_lnl"], what=name + " total (host sum)")
    assert_lnl_close(tm.lnl(), g["total_lnl"], what=name + " total (device reduction)")
    assert tm.engine.launch_count > 0


@pytest.mark.parametrize("name", ["cfg1_gtr_g4", "ambig_hky_ig", "prot12_lg_g4"])
@pytest.mark.parametrize("mode", ["level", "tile"])
def test_every_node_partial_matches_reference(name, mode):
    g, tm = make_tm(name, mode)
    tm.initialise()
    tm.compute_likelihood_at_edge(*tm.traversal.root_edge)
    want = g["partials"] * np.exp(g["scale"])[..., None]
    got = tm.partials * np.exp(tm.scale)[..., None]
    assert got.shape == want.shape
    assert np.allclose(got, want, rtol=1e-10, atol=1e-300)
    rp, rs = tm.root_partials, tm.root_scale
    assert np.allclose(rp * np.exp(rs)[..., None], g["root_partials"] * np.exp(g["root_scale"])[..., None],
                       rtol=1e-10, atol=1e-300)
    # tips come back as the K-fold replicated 0/1 rows the reference stores
    tip = sorted(tm.traversal.names.values())[0]
    assert np.array_equal(tm.partials[tip], g["partials"][tip])


@pytest.mark.parametrize("name", ["deep300_gtr_g4", "ladder120_k80_g4", "prot150_jtt_g4", "ambig_tn93_inv"])
def test_scaled_partials_agree_in_the_log_domain(name):
    g, tm = make_tm(name, "tile")
    tm.initialise()
    tm.compute_likelihood_at_edge(*tm.traversal.root_edge)
    with np.errstate(divide="ignore", invalid="ignore"):
        got = np.log(tm.partials.max(axis=3)) + tm.scale
    want = g["node_logmax"]
    ok = np.isfinite(want) & (want > -1e4)
    assert np.allclose(got[ok], want[ok], rtol=1e-10, atol=1e-9)
    if name != "ambig_tn93_inv":
        assert tm.scale.min() < -80      # the rescaling branch really ran


@pytest.mark.parametrize("name", ["cfg1_gtr_g4", "ambig_hky_ig", "prot12_wag_g4", "nonrev_unrest_g4"])
def test_per_category_values_match_lnl_node(name):
    g, tm = make_tm(name)
    tm.initialise()
    a, b = tm.traversal.root_edge
    length = tm.traversal.brlens[(a, b)]
    _, pattern, cat = tm.engine.root_lnl(a, b, length, want_pattern=True, want_cat=True, root_pmats=tm._root_pmats(length))
    assert_lnl_close(cat, g["cat_lnl"], what=name + " per-category lnL")
    assert_lnl_close(oracle.mix_categories(cat, tm.rate_model.weights), pattern, rtol=1e-13)


@pytest.mark.parametrize("name", sorted(ASC_CASES))
def test_ascertainment_bias_correction(name):
    g, tm = make_tm(name)
    tm.set_ascertainment_bias_correction()
    tm.initialise()
    site = tm.compute_likelihood_at_edge(*tm.traversal.root_edge)
    assert np.all(np.isfinite(g["site_lnl"]))
    assert_lnl_close(site, g["site_lnl"], what=name)
    assert_lnl_close(tm.lnl(), g["total_lnl"], what=name + " total")


def test_device_transition_matrices_match_reference():
    g = load("models")
    t, rates = float(g["t"]), g["rates"]
    from test_substitution_models import REVERSIBLE
    for name, make in REVERSIBLE.items():
        m = make()
        for order, key in ((0, "_p"), (1, "_dp"), (2, "_d2p")):
            got = transition_matrices(m.eigen, t * rates, order)
            assert np.allclose(got, g[name + key], rtol=1e-10, atol=1e-13), (name, key)


def test_pmatrices_inside_the_context_match_model_p():
    g, tm = make_tm("cfg1_gtr_g4")
    tm.initialise()
    lengths = tm._row_lengths()
    for row in (0, len(lengths) - 1):
        for child in (0, 1):
            want = tm.substitution_model.p(lengths[row, child], tm.rate_model.rates)
            assert np.allclose(tm.engine.get_pmatrix(row, child), want, rtol=1e-12, atol=1e-15)


def test_operator_clv_has_reference_semantics():
    g = load("engine_a61")
    sp = np.zeros_like(g["out_scale"])
    out = clv(g["p1"], g["p2"], g["clv1"], g["clv2"], g["sa"], g["sb"], sp)
    assert np.allclose(out, g["out"], rtol=1e-11, atol=0)
    assert np.allclose(sp, g["out_scale"], rtol=1e-13, atol=1e-12)       # written in place
    assert np.allclose(lnl_node(g["pi"], out, sp), g["lnl_node"], rtol=1e-12)
    # single-site call without the leading dimension, and `out=` reuse
    buf = np.empty_like(g["clv1"][0])
    s1 = np.zeros(4)
    r = clv(g["p1"], g["p2"], g["clv1"][0], g["clv2"][0], g["sa"][0], g["sb"][0], s1, buf)
    assert r is buf and np.allclose(buf, g["out"][0], rtol=1e-11) and np.allclose(s1, g["out_scale"][0], rtol=1e-13)
    with pytest.raises(ValueError):
        clv(g["p1"], g["p2"][:, :10, :10], g["clv1"], g["clv2"], g["sa"], g["sb"], sp)


def test_operator_k80_known_answer():
    k80 = phy.substitution_models.K80(2.)
    c = np.array([[[0., 1., 0., 0.]]])
    t = np.array([[[0., 0., 0., 1.]]])
    sc = np.zeros((1, 1))
    part = clv(k80.p(0.1)[None], k80.p(0.2)[None], c, t, np.zeros((1, 1)), np.zeros((1, 1)), sc)
    assert sorted(np.round(part.ravel(), 4).tolist(), reverse=True) == [0.0764, 0.0378, 0.0011, 0.0011]
    assert abs(float(lnl_node(k80.freqs, part, sc).ravel()[0]) + 3.5371) < 5e-5
    for n in np.linspace(0.1, 1.0, 10):
        assert abs(lnl_branch(k80.p(n), k80.freqs, c[0, 0], t[0, 0], 0.0, 0.0) -
                   lnl_branch(k80.p(n), k80.freqs, t[0, 0], c[0, 0], 0.0, 0.0)) < 1e-14


def test_operator_branch_derivs():
    g = load("engine_branch")
    d = lnl_branch_derivs(g["probs"], g["pi"], g["pa"], g["pb"], g["sa"], g["sb"])
    assert np.allclose(d, g["derivs"].reshape(-1, 3), rtol=1e-10, atol=1e-12)
    l0 = lnl_branch(g["probs"][0], g["pi"], g["pa"], g["pb"], g["sa"], g["sb"])
    assert np.allclose(l0, g["lnl"].reshape(-1), rtol=1e-12)


def test_changing_branch_lengths_and_model_reuses_the_context():
    g, tm = make_tm("cfg1_gtr_g4")
    tm.initialise()
    base = tm.lnl()
    key = sorted(tm.traversal.brlens.keys())[3]
    old = tm.traversal.brlens[key]
    tm.traversal.brlens[key] = old * 3
    tm.compute_partials()
    changed = tm.lnl()
    assert abs(changed - base) > 1e-3
    tm.traversal.brlens[key] = old
    tm.compute_partials()
    assert tm.lnl() == base                                   # deterministic, bit for bit
    tm.set_rate_model(phy.rate_models.GammaRateModel(4, 1.5))
    tm.compute_partials()
    _, tr, codes, lut, sw, ii, names, model, _ = problem("cfg1_gtr_g4")
    rate = phy.rate_models.GammaRateModel(4, 1.5)
    want = oracle.tree_lnl(tr, tip_partials(tr, codes, lut, names), model.p, model.freqs, rate.rates, rate.weights)
    assert_lnl_close(tm.lnl(), float(np.dot(want, sw)))


def test_rerooting_on_any_edge_gives_the_same_lnl_only_with_valid_partials():
    g, tm = make_tm("cfg1_gtr_g4")
    tm.initialise()
    with pytest.raises(ValueError):
        tm.compute_likelihood_at_edge(0, 1)                  # no such edge (tree_model.py:184-187)


def test_errors_cross_the_abi_as_python_exceptions():
    eng = phy.LikelihoodEngine(4, 10, 4, 4)
    with pytest.raises(RuntimeError):
        eng.compute_partials()                               # nothing uploaded yet
    with pytest.raises(ValueError):
        eng.set_tips(np.full((4, 10), 99, dtype=np.uint8), np.eye(4), np.arange(4))     # code >= n_codes
    eng.set_tips(np.zeros((4, 10), dtype=np.uint8), np.eye(4), np.array([0, 1, 3, 4]))
    with pytest.raises(ValueError):
        eng.set_schedule(np.array([[5, 2, 0], [2, 0, 1]]))   # child used before it is computed
    with pytest.raises(ValueError):
        eng.set_schedule(np.array([[2, 0, 1], [5, 2, 3]]), level_offsets=np.array([0, 2]))   # same-level dependency
    eng.close()


def test_two_tip_tree_has_no_internal_nodes():
    t = phy.tree.parse_newick("(a:0.1,b:0.2);")
    tm = phy.TreeModel()
    tm.set_tree(t)
    tm.set_alignment([SeqRecord("a", "ACGTAC"), SeqRecord("b", "ACGTTT")], 0)
    tm.set_rate_model(phy.rate_models.UniformRateModel())
    k80 = phy.substitution_models.K80(2.)
    tm.set_substitution_model(k80)
    tm.initialise()
    site = tm.compute_likelihood_at_edge(*tm.traversal.root_edge)
    # a two-leaf tree is left alone by deroot(), so the reference's root-edge rule (utils.py:205-207) takes
    # the LARGER of the two root branches, not their sum
    assert tm.traversal.brlens[tm.traversal.root_edge] == 0.2
    p = k80.p(0.2)
    want = np.log(0.25 * np.array([p[0, 0], p[1, 1], p[2, 2], p[3, 3], p[0, 3], p[1, 3]]))
    assert_lnl_close(site, want, rtol=1e-12)


# ---- operand-resident kernel (4-state models): lnL-only and streaming-store variants ------------------------
DNA_CASES = ["cfg1_gtr_g4", "cfg1_jc_g4", "cfg1_gtr_uniform", "ambig_hky_ig", "ambig_tn93_inv", "deep300_gtr_g4",
             "ladder120_k80_g4"]


@pytest.mark.parametrize("name", [n for n in DNA_CASES if n != "ambig_hky_ig"])      # K=5 is not a resident shape
def test_resident_lnl_only_matches_reference(name):
    g, tr, codes, lut, sw, ii, names, model, rate = problem(name)
    tm = phy.TreeModel(store_partials=False)
    tm.set_tree(tree(g))
    tm.set_alignment(records(g), int(g["alphabet"]))
    tm.set_rate_model(rate)
    tm.set_substitution_model(model)
    tm.initialise()
    site = tm.compute_likelihood_at_edge(*tm.traversal.root_edge)
    assert_lnl_close(site, g["site_lnl"], what=name + " per-site lnL (resident)")
    assert_lnl_close(tm.lnl(), g["total_lnl"], what=name + " total (resident)")
    assert tm.engine.workspace_bytes < 2e8          # no per-node storage: only the fixed L2 parking area
    with pytest.raises(RuntimeError):
        tm.engine.compute_partials()


@pytest.mark.parametrize("name", ["cfg1_gtr_g4", "deep300_gtr_g4", "ladder120_k80_g4"])
def test_resident_store_mode_writes_the_same_partials(name):
    g, tm = make_tm(name, "resident")
    tm.initialise()
    site = tm.compute_likelihood_at_edge(*tm.traversal.root_edge)
    assert_lnl_close(site, g["site_lnl"], what=name)
    g2, tm2 = make_tm(name, "level")
    tm2.initialise()
    tm2.compute_likelihood_at_edge(*tm2.traversal.root_edge)
    assert np.array_equal(tm.partials, tm2.partials) and np.array_equal(tm.scale, tm2.scale)   # bitwise


def test_resident_rejects_unsupported_shapes():
    g, tr, codes, lut, sw, ii, names, model, rate = problem("prot12_lg_g4")
    tm = phy.TreeModel(store_partials=False)
    tm.set_tree(tree(g))
    tm.set_alignment(records(g), 1)
    tm.set_rate_model(rate)
    tm.set_substitution_model(model)
    tm.initialise()
    with pytest.raises(ValueError):
        tm.lnl()


# ---- regressions for state that must not go stale ----------------------------------------------------------------
def _engine_lnl(eng, tr, rows, lengths):
    a, b = tr.root_edge
    eng.set_edge_lengths(lengths)
    eng.build_pmatrices()
    eng.compute_partials()
    return eng.root_lnl(a, b, tr.brlens[(a, b)])[0]


@pytest.mark.parametrize("mode_patterns", [200, 20000])      # level / tile kernels, and the operand-resident pair kernels
def test_set_tips_twice_rebuilds_the_tip_tables(mode_patterns):
    """phb_set_tips after phb_build_pmatrices: the P.lut tip tables depend on the look-up table and on n_codes (8 vs 16
    rows per category) - a second phb_set_tips must invalidate them, or the next pass reads the old tables."""
    rng = np.random.default_rng(11)
    n_taxa, n_pat = 9, mode_patterns
    t = phy.tree.random_tree(n_taxa, 11)
    tr = phy.traversal.Traversal(phy.utils.deepcopy_tree(t))
    labels = [lf.taxon.label for lf in t.leaf_node_iter()]
    tip_nodes = np.asarray([tr.names[n] for n in labels], dtype=np.int32)
    model = phy.substitution_models.GTR([6., 5., 4., 3., 2., 1.], [0.1, 0.2, 0.3, 0.4])
    rate = phy.rate_models.GammaRateModel(4, 0.5)
    rows = tr.locality_order()
    lengths = np.asarray([[tr.brlens[(int(p), int(c1))], tr.brlens[(int(p), int(c2))]] for p, c1, c2 in rows])
    lut5 = np.vstack([np.eye(4)[::-1], np.ones((1, 4))])                                   # 5 codes: 8-row tip tables
    lut12 = np.vstack([lut5, rng.integers(0, 2, size=(7, 4)).astype(float) + np.eye(4)[rng.integers(0, 4, 7)]]).clip(0, 1)
    codes5 = rng.integers(0, 5, size=(n_taxa, n_pat)).astype(np.uint8)
    codes12 = rng.integers(0, 12, size=(n_taxa, n_pat)).astype(np.uint8)                   # 12 codes: 16-row tip tables

    def want(codes, lut):
        tips = {tr.names[n]: np.ascontiguousarray(lut[codes[i]]) for i, n in enumerate(labels)}
        return float(oracle.tree_lnl(tr, tips, model.p, model.freqs, rate.rates, rate.weights).sum())

    eng = phy.LikelihoodEngine(n_taxa, n_pat, 4, 4)
    e = model.eigen
    eng.set_model(e.evecs, e.evals, np.ascontiguousarray(e.ivecs), model.freqs, rate.rates, rate.weights)
    eng.set_tips(codes5, lut5, tip_nodes)
    eng.set_schedule(rows)
    assert_lnl_close(_engine_lnl(eng, tr, rows, lengths), want(codes5, lut5))
    eng.set_tips(codes12, lut12, tip_nodes)                  # other table, other geometry, SAME matrices
    with pytest.raises(RuntimeError):
        eng.compute_partials()                               # the matrices (and their tip tables) are stale now
    assert_lnl_close(_engine_lnl(eng, tr, rows, lengths), want(codes12, lut12))
    eng.set_tips(codes5, lut5[[4, 3, 2, 1, 0]], tip_nodes)   # back to 5 codes, rows permuted
    assert_lnl_close(_engine_lnl(eng, tr, rows, lengths), want(codes5, lut5[[4, 3, 2, 1, 0]]))
    eng.close()


@pytest.mark.parametrize("store", [True, False])
def test_mutating_the_rate_model_in_place_takes_effect(store):
    """The reference reads rate_model.rates on every compute_partials (tree_model.py:166-169): `tm.rate_model.alpha = x`
    followed by compute_partials() must evaluate the new categories here too."""
    g, tr, codes, lut, sw, ii, names, model, _ = problem("cfg1_gtr_g4")
    rate = phy.rate_models.GammaRateModel(4, 0.5)
    tm = phy.TreeModel(store_partials=store)
    tm.set_tree(tree(g))
    tm.set_alignment(records(g), 0)
    tm.set_rate_model(rate)
    tm.set_substitution_model(model)
    tm.initialise()
    assert_lnl_close(tm.lnl(), g["total_lnl"])
    rate.alpha = 2.0
    tm.compute_partials()
    fresh = phy.rate_models.GammaRateModel(4, 2.0)
    want = oracle.tree_lnl(tr, tip_partials(tr, codes, lut, names), model.p, model.freqs, fresh.rates, fresh.weights)
    assert_lnl_close(tm.lnl(), float(np.dot(want, sw)))
    rate.alpha = 0.5
    tm.compute_partials()
    assert_lnl_close(tm.lnl(), g["total_lnl"])
    # +I+G: pinvar changes rates AND weights
    ig = phy.rate_models.InvariantGammaModel(0.1, 3, 0.7)
    tm.set_rate_model(ig)
    tm.compute_partials()
    ig.pinvar = 0.35
    tm.compute_partials()
    fresh = phy.rate_models.InvariantGammaModel(0.35, 3, 0.7)
    want = oracle.tree_lnl(tr, tip_partials(tr, codes, lut, names), model.p, model.freqs, fresh.rates, fresh.weights)
    assert_lnl_close(tm.lnl(), float(np.dot(want, sw)))


def test_device_codon_matrices_match_the_textbook_definition():
    """GY94 has no counterpart in the reference: pin the device-built P (csrc/pmatrix.cu, A = 61) against
    scipy's expm of the rate matrix written out from the published definition (tests/test_substitution_models.py)."""
    from scipy.linalg import expm
    from test_substitution_models import _textbook_gy94
    from phylo_utils_b200.substitution_models.codon import f3x4, SENSE_CODONS
    pi = f3x4(np.random.default_rng(4).dirichlet(np.ones(4) * 5, size=3))
    m = phy.substitution_models.GY94(2.0, 0.2, pi)
    q = _textbook_gy94(2.0, 0.2, pi, SENSE_CODONS)
    times = np.array([0.003, 0.11, 0.9, 2.7])
    got = transition_matrices(m.eigen, times, 0)
    want = np.stack([expm(q * t) for t in times])
    assert np.allclose(got, want, rtol=1e-9, atol=1e-13)
    assert np.allclose(transition_matrices(m.eigen, times, 1), np.stack([q.dot(expm(q * t)) for t in times]), rtol=1e-8, atol=1e-12)
