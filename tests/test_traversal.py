"""Tree container, Newick I/O and the schedule tables (reference: traversal.py, utils.py:114-213)."""
import copy

import numpy as np
import pytest

import phylo_utils_b200 as phy
from phylo_utils_b200.tree import parse_newick, random_tree, caterpillar_tree, balanced_tree
from phylo_utils_b200.traversal import Traversal
from phylo_utils_b200.utils import deepcopy_tree, BranchLengths
from helpers import load


def build(newick):
    return Traversal(deepcopy_tree(parse_newick(newick)))


@pytest.mark.parametrize("key", [str(k) for k in load("traversal")["keys"]])
def test_tables_equal_reference(key):
    g = load("traversal")
    t = build(str(g[key + "_newick"]))
    assert np.array_equal(t.postorder_traversal, g[key + "_postorder"]) and t.postorder_traversal.dtype == np.int64
    assert np.array_equal(t.optimising_traversal, g[key + "_optimising"])
    assert tuple(t.root_edge) == tuple(g[key + "_root_edge"])
    keys = sorted(t.brlens.keys())
    assert np.array_equal(np.asarray(keys), g[key + "_brlen_keys"])
    assert np.array_equal(np.asarray([t.brlens[k] for k in keys]), g[key + "_brlen_vals"])
    assert sorted(t.names, key=t.names.get) == [str(n) for n in g[key + "_tip_names"]]
    assert sorted(t.names.values()) == g[key + "_tip_nodes"].tolist()


def test_table_shapes_and_sweep_structure():
    for n in (3, 4, 9, 57):
        t = Traversal(deepcopy_tree(random_tree(n, n)))
        assert t.postorder_traversal.shape == (n - 2, 3)
        assert t.optimising_traversal.shape == (3 * n - 5, 5)
        assert len(t.brlens) == 2 * n - 3
        sweep = t.optimising_traversal
        assert sweep[0, :3].tolist() == [-1, -1, -1] and tuple(sweep[0, 3:]) == t.root_edge
        optimised = {tuple(sorted(r[3:])) for r in sweep if r[3] >= 0}
        assert optimised == set(t.brlens.keys())          # every edge is visited exactly once
        assert sum(r[3] >= 0 for r in sweep) == 2 * n - 3


def test_branch_lengths_lookup_is_symmetric():
    b = BranchLengths()
    b[(1, 2)] = 0.5
    assert b[(2, 1)] == 0.5 and b[1, 2] == 0.5 and (2, 1) in b
    with pytest.raises(KeyError):
        b[(1, 3)]


def _check_order(t, rows):
    done = set(t.names.values())
    assert sorted(map(tuple, rows)) == sorted(map(tuple, t.postorder_traversal))
    for par, c1, c2 in rows:
        assert c1 in done and c2 in done
        done.add(int(par))


@pytest.mark.parametrize("maker,n", [(random_tree, 40), (caterpillar_tree, 25), (balanced_tree, 33), (random_tree, 3)])
def test_level_and_locality_orders_are_valid_schedules(maker, n):
    t = Traversal(deepcopy_tree(maker(n, 7)))
    rows, off = t.level_order()
    _check_order(t, rows)
    level = t.node_levels()
    for l in range(len(off) - 1):
        assert set(level[rows[off[l]:off[l + 1], 0]]) == {l + 1}
    assert off[0] == 0 and off[-1] == len(rows)
    loc = t.locality_order()
    _check_order(t, loc)
    # in the locality order most rows with an internal child use the row just before them
    hits = sum(1 for i in range(1, len(loc)) if loc[i - 1, 0] in (loc[i, 1], loc[i, 2]))
    with_internal = sum(1 for r in loc if not (t.is_leaf(r[1]) and t.is_leaf(r[2])))
    if with_internal:
        assert hits >= with_internal - 2


def test_caterpillar_depth_and_balanced_depth():
    lad = Traversal(deepcopy_tree(caterpillar_tree(30, 1)))
    assert len(lad.level_order()[1]) - 1 >= 27
    bal = Traversal(deepcopy_tree(balanced_tree(32, 1)))
    assert len(bal.level_order()[1]) - 1 <= 6


def test_newick_round_trip_and_quoting():
    src = "((a:0.1,'b c':0.25)x:0.5,(d:1e-3,e:2)[comment]:0.125,f:3);"
    t = parse_newick(src)
    assert [l.taxon.label for l in t.leaf_node_iter()] == ["a", "b c", "d", "e", "f"]
    again = parse_newick(t.as_newick())
    assert again.as_newick() == t.as_newick()
    lens = [n.edge_length for n in again.preorder_node_iter() if n.edge_length is not None]
    assert 0.125 in lens and 1e-3 in lens
    for bad in ("((a,b);", "(a,b));", ""):
        with pytest.raises(ValueError):
            parse_newick(bad)


def test_deepcopy_does_not_touch_the_original_and_survives_deep_trees():
    tree = caterpillar_tree(3000, 2)
    before = tree.as_newick()
    clone = deepcopy_tree(tree)
    assert tree.as_newick() == before
    assert len(clone.seed_node.child_nodes()) == 2
    t = Traversal(clone)
    assert t.postorder_traversal.shape == (2998, 3)
    assert copy.deepcopy(tree).as_newick() == before


def test_root_edge_length_is_the_sum_of_the_two_root_branches():
    t = build("((a:0.1,b:0.2):0.05,(c:0.3,d:0.4):0.07);")
    a, b = t.root_edge
    assert abs(t.brlens[(a, b)] - 0.12) < 1e-15


def test_traversal_needs_a_binary_root():
    with pytest.raises(ValueError):
        Traversal(parse_newick("(a:1,b:1,c:1);"))


def test_edge_keys_and_vectorised_length_gathers_match_the_dictionary():
    """Host logic of the Newton driver: the brlens keys of the edges above a list of nodes are resolved once
    (TreeModel.edge_keys) and lengths move through them; both must agree with the plain dictionary look-ups."""
    import phylo_utils_b200 as phy
    from phylo_utils_b200.optimise import edge_nodes, newton_step, _get_lengths, _set_lengths
    tm = phy.TreeModel()
    tm.set_tree(random_tree(37, 3))
    tr = tm.traversal
    nodes = edge_nodes(tr)
    a, b = tr.root_edge
    assert b not in nodes and a in nodes and len(nodes) == 2 * 37 - 3
    keys = tm.edge_keys(nodes)
    assert len(set(keys)) == len(keys) == len(tr.brlens)            # one key per edge of the unrooted tree, all covered
    want = np.array([tm.branch_length_above(int(n)) for n in nodes])
    assert np.array_equal(tm.lengths_above(nodes), want)
    assert np.array_equal(_get_lengths(tr.brlens, keys), want)
    _set_lengths(tr.brlens, keys, want * 2.0)
    assert np.array_equal(tm.lengths_above(nodes), want * 2.0)
    assert tm.branch_length_above(b) == tm.branch_length_above(a)  # either root child means the root edge
    with pytest.raises(ValueError):
        tm.edge_keys([2 * 37])                                      # no such node
    # safeguarded Newton step: concave -> Newton, convex -> doubling / halving, always inside the bounds
    t = np.array([0.1, 0.1, 0.1, 1e-6, 19.0])
    out = newton_step(t, np.array([1.0, 1.0, -1.0, -5.0, 3.0]), np.array([-20.0, 5.0, 5.0, -1.0, 1.0]))
    assert np.allclose(out[:3], [0.15, 0.2, 0.05])
    assert out[3] >= 1.0 / 2 ** 16 and out[4] <= 20.0
