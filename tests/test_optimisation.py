"""
Branch-length optimisers: the scalar minimisers under the reference's names
(src/optimisation.pyx), the batched bracketing search, and the two sweep drivers.  CPU tests drive the sweeps through an
oracle-backed stand-in for the engine (host logic only); `gpu` tests use the real engine.
"""
import numpy as np
import pytest
from scipy.optimize import minimize_scalar
from scipy.special import logsumexp

import phylo_utils_b200 as phy
from phylo_utils_b200 import optimisation as opt
from phylo_utils_b200.optimise import optimise_by_rerooting, optimise_branch_lengths, _maximise_edge, edge_nodes
from phylo_utils_b200.tree import random_tree
from helpers import problem, records, tree, tip_partials, simulate_codes
from oracle import oracle


# ---- scalar routines --------------------------------------------------------------------------------------------
def test_simplex_transforms_round_trip():
    rng = np.random.default_rng(0)
    for n in (2, 4, 20, 61):
        p = rng.dirichlet(np.ones(n))
        theta = opt.simplex_encode(p)
        assert theta.shape == (n - 1,) and np.all((theta > 0) & (theta < 1))
        assert np.allclose(opt.simplex_decode(theta), p, rtol=1e-12, atol=1e-15)
        q = opt.transform_params(p)
        assert np.all(np.isfinite(q)) and np.allclose(opt.decode_params(q), p, rtol=1e-10, atol=1e-14)
    assert np.allclose(opt.simplex_decode(np.array([0.1, 0.5])), [0.1, 0.45, 0.45])       # stick breaking, by hand


def test_quad_interp_finds_the_vertex():
    f = lambda x: 3.0 * (x - 1.7) ** 2 - 4.0          # noqa: E731
    assert abs(opt.quad_interp(0.0, 1.0, 3.0, f(0.0), f(1.0), f(3.0)) - 1.7) < 1e-12
    assert np.isfinite(opt.quad_interp(0.0, 1.0, 2.0, 1.0, 1.0, 1.0))        # flat: guarded division


def test_brent_and_dbrent_agree_with_scipy():
    cases = [(lambda x: (x - 0.3) ** 2, lambda x: 2 * (x - 0.3), 0.0, 1.0, 0.9),
             (lambda x: np.cosh(x - 2.0), lambda x: np.sinh(x - 2.0), -1.0, 6.0, 0.0),
             (lambda x: -(5 * np.log(x) - 12 * x), lambda x: -(5 / x - 12), 1e-5, 10.0, 5.0)]
    for f, df, lo, hi, guess in cases:
        want = minimize_scalar(f, bounds=(lo, hi), method="bounded", options={"xatol": 1e-12}).x
        x, fx, it = opt.brent_wrap(guess, lo, hi, f)
        assert abs(x - want) < 1e-6 * max(1.0, abs(want)) and abs(fx - f(want)) < 1e-10 and it <= opt.ITMAX
        x, fx, it = opt.dbrent_wrap(guess, lo, hi, f, df)
        assert abs(x - want) < 1e-6 * max(1.0, abs(want)) and it <= opt.ITMAX


def test_batched_bracketing_search():
    a = np.array([1.0, 2.0, 0.5, 3.0, 1e-4, 0.0])
    b = np.array([10.0, 1.0, 100.0, 0.1, 5.0, 1.0])
    calls = []

    def fn(t, idx):                                   # lnL-like curves a log t - b t: maximum at a / b
        calls.append(len(idx))
        with np.errstate(divide="ignore"):
            return np.stack([a[idx] * np.log(t) - b[idx] * t, a[idx] / t - b[idx]], axis=1)
    x, fx, dx, evals = opt.maximise_bracketed(fn, 1e-6, 20.0, np.full(6, 0.1), tol=1e-9)
    assert np.allclose(x[[0, 1, 2, 4]], (a / b)[[0, 1, 2, 4]], rtol=1e-6)
    assert x[3] == 20.0 and x[5] == 1e-6              # monotone on the interval: the boundary the derivative points to
    assert evals == len(calls) and calls[0] == 18 and calls[-1] < 6      # both ends + start in one launch; finished curves drop out


# ---- the sweeps on an oracle-backed engine (host logic, no GPU) ------------------------------------------------------
class OracleEngine(object):
    """update_node / branch_derivatives / root_lnl of LikelihoodEngine, computed by the CPU oracle."""

    def __init__(self, tr, codes, lut, names, model, rate, weights):
        self.tr, self.model, self.rate, self.w = tr, model, rate, np.asarray(weights, dtype=float)
        self.ot = oracle.OracleTree(2 * len(tr.names) - 2, tip_partials(tr, codes, lut, names), rate.ncat)
        self.updates = 0

    def compute_partials(self):
        rows = np.asarray(self.tr.postorder_traversal, dtype=np.int64)
        pm = np.stack([np.stack([self.model.p(self.tr.brlens[(int(p), int(c))], self.rate.rates) for c in (c1, c2)])
                       for p, c1, c2 in rows])
        self.ot.partials[[int(r[0]) for r in rows]] = 0
        self.ot.scale[:] = 0
        self.ot.compute_partials(rows, pm)

    def update_node(self, node, a, la, b, lb):
        sc = np.zeros_like(self.ot.scale[0])
        self.ot.partials[node] = oracle.clv(self.model.p(la, self.rate.rates), self.model.p(lb, self.rate.rates),
                                            self.ot.partials[a], self.ot.partials[b], self.ot.scale[a], self.ot.scale[b], sc)
        self.ot.scale[node] = sc
        self.updates += 1

    def branch_derivatives(self, x, y, lengths, chain_rule=True):
        out = []
        for t in np.atleast_1d(lengths):
            cols = []
            for k, r in enumerate(self.rate.rates):
                probs = np.stack([self.model.p(t * r), self.model.dp_dt(t * r) * r, self.model.d2p_dt2(t * r) * r * r])
                cols.append(oracle.lnl_branch_derivs(probs, self.model.freqs, self.ot.partials[x][:, k], self.ot.partials[y][:, k],
                                                     self.ot.scale[x][:, k], self.ot.scale[y][:, k]))
            d = np.stack(cols, axis=1)                                            # (S, K, 3)
            logw = np.log(self.rate.weights)
            lnl = logsumexp(d[:, :, 0] + logw, axis=1)
            post = np.exp(d[:, :, 0] + logw - lnl[:, None])
            d1 = (post * d[:, :, 1]).sum(1)
            d2 = (post * (d[:, :, 2] + d[:, :, 1] ** 2)).sum(1) - d1 ** 2
            out.append([np.dot(self.w, lnl), np.dot(self.w, d1), np.dot(self.w, d2)])
        return np.asarray(out)

    def root_lnl(self, a, b, length):
        return (float(self.branch_derivatives(a, b, [length])[0, 0]),)


def simulated_problem(n_taxa=9, n_sites=1500, seed=5):
    """A data set that HAS an interior optimum: sites evolved on the tree under GTR+G4."""
    t = random_tree(n_taxa, seed)
    model = phy.substitution_models.GTR([6., 5., 4., 3., 2., 1.], [0.1, 0.2, 0.3, 0.4])
    rate = phy.rate_models.GammaRateModel(4, 0.8)
    codes, lut, names = simulate_codes(t, model, rate, n_sites, seed)
    return t, codes, lut, names, model, rate


class OracleTreeModel(object):
    def __init__(self, name=None):
        if name is None:
            t, codes, lut, names, model, rate = simulated_problem()
            tr = phy.traversal.Traversal(phy.utils.deepcopy_tree(t))
            sw = np.ones(codes.shape[1])
        else:
            g, tr, codes, lut, sw, ii, names, model, rate = problem(name)
        self.traversal, self.engine = tr, OracleEngine(tr, codes, lut, names, model, rate, sw)

    def compute_partials(self):
        self.engine.compute_partials()

    def lnl(self):
        a, b = self.traversal.root_edge
        return self.engine.root_lnl(a, b, self.traversal.brlens[(a, b)])[0]


def test_rerooting_sweep_host_logic_on_the_oracle():
    tm = OracleTreeModel()
    start = tm.traversal.brlens.copy()
    tm.compute_partials()
    at_truth = tm.lnl()
    for key in list(tm.traversal.brlens.keys())[::2]:
        tm.traversal.brlens[key] = tm.traversal.brlens[key] * 3.0            # knock the tree off its optimum
    tm.compute_partials()
    before = tm.lnl()
    res = optimise_by_rerooting(tm, max_sweeps=6, tol=1e-6)
    n = len(tm.traversal.names)
    assert np.all(np.diff(res["trace"]) >= -1e-9) and res["lnl"] > before + 1.0 and res["lnl"] >= at_truth
    for key, t in tm.traversal.brlens.items():                                # back in the neighbourhood of the simulating lengths
        assert abs(t - start[key]) < 0.1 + 0.5 * start[key], (key, t, start[key])
    assert res["node_updates"] == res["sweeps"] * (3 * n - 6)               # every row but the root-edge row rebuilds a node
    # at the end every partial points at the root again: a from-scratch evaluation at the new lengths gives the same lnL
    tm.compute_partials()
    assert abs(tm.lnl() - res["lnl"]) <= 1e-9 * abs(res["lnl"])
    # and every edge sits at a stationary point of its own curve
    a, b = tm.traversal.root_edge
    d = tm.engine.branch_derivatives(a, b, [tm.traversal.brlens[(a, b)]])[0]
    assert d[2] < 0 and abs(d[1] / d[2]) < 2e-3                              # a Newton step from here would move it by < 0.002
    assert set(start.keys()) == set(tm.traversal.brlens.keys())


def test_maximise_edge_newton_and_fallback():
    curve = lambda t: np.stack([7 * np.log(t) - 20 * t, 7 / t - 20, -7 / t ** 2], axis=1)        # noqa: E731
    t, f, n = _maximise_edge(curve, 0.2)
    assert abs(t - 0.35) < 1e-6 and n <= 8
    bumpy = lambda t: np.stack([-np.cos(3 * t) - 0.1 * t, 3 * np.sin(3 * t) - 0.1, 9 * np.cos(3 * t)], axis=1)   # noqa: E731
    t, f, n = _maximise_edge(bumpy, 0.1, lo=1e-5, hi=2.0)                                       # convex at the start
    assert abs(3 * np.sin(3 * t) - 0.1) < 1e-4 and 9 * np.cos(3 * t) < 0


# ---- the real engine ---------------------------------------------------------------------------------------------
def _gpu_model(name, up=False):
    tm = phy.TreeModel(up_partials=up)
    if name is None:
        t, codes, lut, names, model, rate = simulated_problem()
        tm.set_tree(t)
        tm.set_tip_codes(codes, lut, names)
    else:
        g, tr, codes, lut, sw, ii, names, model, rate = problem(name)
        tm.set_tree(tree(g))
        tm.set_alignment(records(g), int(g["alphabet"]))
    tm.set_rate_model(rate)
    tm.set_substitution_model(model)
    tm.initialise()
    return tm


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["cfg1_gtr_g4", "prot12_lg_g4"])
def test_update_node_and_branch_derivatives_match_the_oracle(name):
    tm, ref = _gpu_model(name), OracleTreeModel(name)
    ref.compute_partials()
    table = np.asarray(tm.traversal.optimising_traversal).tolist()
    br = tm.traversal.brlens
    for par, c1, c2, x, y in table[:9]:
        if par >= 0:
            tm.engine.update_node(par, c1, br[(par, c1)], c2, br[(par, c2)])
            ref.engine.update_node(par, c1, br[(par, c1)], c2, br[(par, c2)])
            got = tm.engine.get_partials(par) * np.exp(tm.engine.get_scalers(par))[..., None]
            want = ref.engine.ot.partials[par] * np.exp(ref.engine.ot.scale[par])[..., None]
            assert np.allclose(got, want, rtol=1e-10, atol=1e-300)
        if x >= 0:
            trial = np.array([br[(x, y)], br[(x, y)] * 0.4, 1.3])
            got = tm.engine.branch_derivatives(x, y, trial)
            want = ref.engine.branch_derivatives(x, y, trial)
            assert np.allclose(got[:, 0], want[:, 0], rtol=1e-10)
            assert np.allclose(got[:, 1:], want[:, 1:], rtol=1e-8, atol=1e-7)
    with pytest.raises(ValueError):
        tm.engine.update_node(0, 1, 0.1, 2, 0.1)           # a tip cannot be rebuilt


@pytest.mark.gpu
def test_rerooting_sweep_and_batched_newton_reach_the_same_optimum():
    results = {}
    for which in ("reroot", "newton"):
        tm = _gpu_model(None, up=(which == "newton"))
        for key in list(tm.traversal.brlens.keys())[::2]:
            tm.traversal.brlens[key] = tm.traversal.brlens[key] * 3.0
        tm.compute_partials()
        before = tm.lnl()
        if which == "reroot":
            res = optimise_by_rerooting(tm, max_sweeps=8, tol=1e-7)
            tm.compute_partials()
            assert abs(tm.lnl() - res["lnl"]) <= 1e-10 * abs(res["lnl"])   # partials were restored exactly
        else:
            res = optimise_branch_lengths(tm, max_sweeps=40, inner_iterations=3, tol=1e-7)
        assert np.all(np.diff(res["trace"]) >= -1e-9) and res["lnl"] > before
        results[which] = (res["lnl"], tm.traversal.brlens.copy())
    assert abs(results["reroot"][0] - results["newton"][0]) < 1e-3
    for key, t in results["reroot"][1].items():
        assert abs(t - results["newton"][1][key]) < 1e-2 * max(t, 0.01)     # two routes to a flat maximum


@pytest.mark.gpu
def test_newton_sweep_falls_back_to_the_bracket_on_non_concave_edges():
    tm = _gpu_model(None, up=True)
    for key in tm.traversal.brlens.keys():
        tm.traversal.brlens[key] = 6.0                      # saturated branches: curves are flat / convex out here
    tm.compute_partials()
    before = tm.lnl()
    res = optimise_branch_lengths(tm, max_sweeps=25, inner_iterations=2, tol=1e-6)
    assert res["fallback_edges"] > 0 and np.all(np.diff(res["trace"]) >= 0) and res["lnl"] > before + 100
    tm2 = _gpu_model(None, up=True)
    best = optimise_branch_lengths(tm2, max_sweeps=40, inner_iterations=3, tol=1e-7)["lnl"]
    assert abs(res["lnl"] - best) < 0.05                     # the same optimum as from the true starting lengths
