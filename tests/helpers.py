"""Shared test helpers: load golden fixtures and rebuild the same problem with phylo_utils_b200 / the oracle."""
import os

import numpy as np

import phylo_utils_b200 as phy
from phylo_utils_b200.alignment.alignment import SeqRecord, alignment_to_codes
from phylo_utils_b200.tree import parse_newick

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RTOL = 1e-10   # BASELINE.json north_star: total and per-site lnL within 1e-10 relative in fp64


def load(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


def records(g):
    return [SeqRecord(str(n), bytes(row).decode("ascii")) for n, row in zip(g["names"], g["seqs"])]


def tree(g):
    return parse_newick(str(g["newick"]))


sm = phy.substitution_models
rm = phy.rate_models
_GTR = lambda: sm.GTR([6., 5., 4., 3., 2., 1.], [0.1, 0.2, 0.3, 0.4])
_UNREST = lambda: sm.Unrest(rates=[[0., 1., 2., 3.], [4., 0., 5., 6.], [7., 8., 0., 9.], [10., 11., 12., 0.]])

# name -> (substitution model factory, rate model factory); must match oracle/make_golden.py
CASES = {
    "cfg1_gtr_g4": (_GTR, lambda: rm.GammaRateModel(4, 0.5)),
    "cfg1_jc_g4": (lambda: sm.GTR(), lambda: rm.GammaRateModel(4, 0.5)),
    "cfg1_gtr_uniform": (_GTR, lambda: rm.UniformRateModel()),
    "ambig_hky_ig": (lambda: sm.HKY85(2.5, [0.3, 0.2, 0.15, 0.35]), lambda: rm.InvariantGammaModel(0.2, 4, 0.7)),
    "ambig_tn93_inv": (lambda: sm.TN93(2.0, 3.0, 1.0, [0.25, 0.2, 0.3, 0.25]), lambda: rm.InvariantSitesModel(0.3)),
    "deep300_gtr_g4": (_GTR, lambda: rm.GammaRateModel(4, 0.5)),
    "ladder120_k80_g4": (lambda: sm.K80(2.0), lambda: rm.GammaRateModel(4, 1.3)),
    "prot12_lg_g4": (lambda: sm.LG(), lambda: rm.GammaRateModel(4, 0.8)),
    "prot12_wag_g4": (lambda: sm.WAG(), lambda: rm.GammaRateModel(4, 0.8)),
    "prot150_jtt_g4": (lambda: sm.JTT(), lambda: rm.GammaRateModel(4, 0.6)),
    "nonrev_unrest_g4": (_UNREST, lambda: rm.GammaRateModel(4, 0.5)),
}
ASC_CASES = {
    "ascbias_gtr_uniform": (_GTR, lambda: rm.UniformRateModel()),
    "ascbias_gtr_g4": (_GTR, lambda: rm.GammaRateModel(4, 2.0)),
}


def problem(name):
    """-> (golden dict, traversal, codes, lut, siteweights, inverse_index, names, model, rate_model)"""
    g = load(name)
    factories = CASES.get(name) or ASC_CASES[name]
    model, rate = factories[0](), factories[1]()
    tr = phy.traversal.Traversal(phy.utils.deepcopy_tree(tree(g)))
    codes, lut, sw, ii, names = alignment_to_codes(records(g), int(g["alphabet"]))
    return g, tr, codes, lut, sw, ii, names, model, rate


def tip_partials(tr, codes, lut, names):
    return {tr.names[nm]: np.ascontiguousarray(lut[codes[names[nm]]]) for nm in tr.names}


def assert_lnl_close(got, want, rtol=RTOL, what="lnL"):
    got, want = np.asarray(got, dtype=float), np.asarray(want, dtype=float)
    assert got.shape == want.shape, "{}: shape {} vs {}".format(what, got.shape, want.shape)
    finite = np.isfinite(want)
    assert np.array_equal(np.isfinite(got), finite), "{}: finiteness pattern differs".format(what)
    if finite.any():
        err = np.abs(got[finite] - want[finite]) / np.maximum(np.abs(want[finite]), 1e-300)
        assert err.max() <= rtol, "{}: max relative error {:.3e} > {:.1e}".format(what, err.max(), rtol)
    if (~finite).any():
        assert np.array_equal(got[~finite], want[~finite]) or np.all(np.isnan(want[~finite]) == np.isnan(got[~finite]))


# ---- derivative oracle by composition (SURVEY.md 8(c) "not in the reference at all") ---------------------------
def oracle_up_partials(tr, ot, model, rates):
    """
    Pre-order partials composed from the oracle's clv: up[c] = clv(P(p,sib), P(p,gpa), down[sib], X)
    following the reference's re-rooting rows [PAR,SIB,GPA,NOD,PAR] (utils.py:169).  Returns
    {node: (partials (S,K,A), scale (S,K))}.
    """
    from oracle import oracle
    a, b = tr.root_edge
    up = {}
    for par, c1, c2 in tr.postorder_traversal[::-1]:
        par, c1, c2 = int(par), int(c1), int(c2)
        if par in (a, b):
            other = b if par == a else a
            x_part, x_scale, x_len = ot.partials[other], ot.scale[other], tr.brlens[(a, b)]
        else:
            gpa = [int(r[0]) for r in tr.postorder_traversal if par in (int(r[1]), int(r[2]))][0]
            x_part, x_scale = up[par]
            x_len = tr.brlens[(par, gpa)]
        for child, sib in ((c1, c2), (c2, c1)):
            sc = np.zeros_like(ot.scale[0])
            part = oracle.clv(model.p(tr.brlens[(par, sib)], rates), model.p(x_len, rates), ot.partials[sib], x_part,
                              ot.scale[sib], x_scale, sc)
            up[child] = (part, sc)
    return up


def oracle_edge_derivatives(tr, ot, up, model, rate, node, t, siteweights, chain_rule=True):
    """(lnL, dlnL/dt, d2lnL/dt2) for the edge above ``node`` from per-category lnl_branch_derivs outputs."""
    from oracle import oracle
    from scipy.special import logsumexp
    a, b = tr.root_edge
    if node in (a, b):
        other = b if node == a else a
        pb, sb = ot.partials[other], ot.scale[other]
    else:
        pb, sb = up[node]
    pa, sa = ot.partials[node], ot.scale[node]
    S, K = sa.shape
    lnf = np.empty((S, K))
    g1 = np.empty((S, K))
    g2 = np.empty((S, K))
    for k, r in enumerate(rate.rates):
        c1, c2 = (r, r * r) if chain_rule else (1.0, 1.0)
        probs = np.stack([model.p(t * r), model.dp_dt(t * r) * c1, model.d2p_dt2(t * r) * c2])
        d = oracle.lnl_branch_derivs(probs, model.freqs, pa[:, k], pb[:, k], sa[:, k], sb[:, k])
        lnf[:, k], g1[:, k], g2[:, k] = d[:, 0], d[:, 1], d[:, 2]
    # a rate-0 (invariant) category has f_k = 0 at variable patterns: log f = -inf, the ratios are 0/0.
    # Its f'_k and f''_k vanish too (dP/dt carries the factor r_k = 0), so it contributes nothing.
    dead = ~np.isfinite(lnf)
    lnf = np.where(dead, -np.inf, lnf)
    g1 = np.where(dead | ~np.isfinite(g1), 0.0, g1)
    g2 = np.where(dead | ~np.isfinite(g2), 0.0, g2)
    logw = np.log(rate.weights)
    lnL = logsumexp(lnf + logw, axis=1)
    post = np.exp(lnf + logw - lnL[:, None])              # w_k f_k / L
    d1 = (post * g1).sum(axis=1)                           # L'/L
    d2 = (post * (g2 + g1 * g1)).sum(axis=1) - d1 * d1     # L''/L - (L'/L)^2
    w = np.asarray(siteweights, dtype=float)
    return np.array([np.dot(w, lnL), np.dot(w, d1), np.dot(w, d2)])


# ---- a tiny sequence simulator for the optimiser tests (sites evolved down the tree under the model) --------------------
def simulate_codes(tr_tree, model, rate, n_sites, seed):
    """-> (codes uint8 (ntax, nsites) in state order, lut = identity, names {label: row}); one Gamma category per site."""
    rng = np.random.default_rng(seed)
    A = model.size
    cats = rng.integers(0, rate.ncat, n_sites)
    states = {}
    for nd in tr_tree.preorder_node_iter():
        if nd.parent_node is None:
            states[nd] = rng.choice(A, size=n_sites, p=np.asarray(model.freqs) / np.sum(model.freqs))
            continue
        pm = np.stack([model.p(nd.edge_length * r) for r in rate.rates])          # (K, A, A)
        rows = np.clip(pm[cats, states[nd.parent_node]], 0, None)                  # (n_sites, A) transition rows
        cdf = np.cumsum(rows / rows.sum(1, keepdims=True), axis=1)
        states[nd] = np.minimum((rng.random(n_sites)[:, None] > cdf).sum(1), A - 1)
    leaves = list(tr_tree.leaf_node_iter())
    codes = np.stack([states[lf] for lf in leaves]).astype(np.uint8)
    return codes, np.eye(A), {lf.taxon.label: i for i, lf in enumerate(leaves)}
