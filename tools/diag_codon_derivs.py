"""Diagnostic: 61-state edge derivatives - first pass (DMMA kernel) vs the sum-table pass vs the composed oracle."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import phylo_utils_b200 as phy  # noqa: E402
from phylo_utils_b200.optimise import edge_nodes  # noqa: E402
from phylo_utils_b200.substitution_models.codon import f3x4  # noqa: E402
from phylo_utils_b200.tree import random_tree  # noqa: E402
import helpers  # noqa: E402
from oracle import oracle  # noqa: E402


def main(n_taxa=14, n_pat=300):
    rng = np.random.default_rng(4)
    model = phy.substitution_models.GY94(2.0, 0.2, f3x4(rng.dirichlet(np.ones(4) * 5, size=3)))
    rate = phy.rate_models.GammaRateModel(4, 0.5)
    tree = random_tree(n_taxa, 5)
    names = [l.taxon.label for l in tree.leaf_node_iter()]
    codes = rng.integers(0, 62, size=(n_taxa, n_pat)).astype(np.uint8)
    lut = np.vstack([np.eye(61)[::-1], np.ones((1, 61))])
    tm = phy.TreeModel(up_partials=True)
    tm.set_tree(tree)
    tm.set_tip_codes(codes, lut, {n: i for i, n in enumerate(names)})
    tm.set_rate_model(rate)
    tm.set_substitution_model(model)
    tm.initialise()
    tm.compute_up_partials()
    nodes = edge_nodes(tm.traversal)
    lengths = tm.lengths_above(nodes)
    first = tm.edge_derivatives(nodes, lengths)
    again = tm.edge_derivatives(nodes, lengths)
    third = tm.edge_derivatives(nodes, lengths)
    rel = lambda a, b: np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300), axis=0)   # noqa: E731
    print("first vs again (max rel per column):", rel(first, again))
    print("again vs third:", rel(again, third))
    tr = tm.traversal
    tips = {tr.names[n]: np.ascontiguousarray(lut[codes[i]]) for i, n in enumerate(names)}
    _, ot = oracle.tree_lnl(tr, tips, model.p, model.freqs, rate.rates, rate.weights, return_tree=True)
    up = helpers.oracle_up_partials(tr, ot, model, rate.rates)
    want = np.array([helpers.oracle_edge_derivatives(tr, ot, up, model, rate, int(n), float(t), np.ones(n_pat)) for n, t in zip(nodes, lengths)])
    print("first vs oracle:", rel(first, want))
    print("again vs oracle:", rel(again, want))
    worst = np.argmax(np.abs(first - again)[:, 1] / np.abs(want[:, 1]))
    print("worst edge", worst, "node", nodes[worst], "first", first[worst], "again", again[worst], "oracle", want[worst])


if __name__ == "__main__":
    main(*[int(a) for a in sys.argv[1:]])
