"""
Tree -> integer schedule tables.

``Traversal`` keeps the reference's attributes and numbering
(``/root/reference/phylo_utils/traversal.py:6-35``): nodes are numbered in postorder with
the seed node skipped, ``names`` maps leaf labels to node ids, ``root_edge`` is the pair of
root children, ``brlens`` the symmetric edge-length lookup, ``postorder_traversal`` the
``(N-2, 3)`` ``[PAR, CH1, CH2]`` table and ``optimising_traversal`` the ``(3N-5, 5)``
re-rooting sweep.

On top of that it derives what the device wants:

* ``level_order()``  - the same rows grouped by depth (all rows of one level are
  independent, one launch per level);
* ``locality_order()`` - a valid postorder that visits the larger child subtree first, so the
  second operand of almost every row is the row just before it (site-tile resident kernels
  re-read it from L2 / registers rather than HBM);
* ``preorder_edges()`` - edges in root-to-tip order for the up-partial / derivative sweep.
"""
import numpy as np

from .utils import get_postorder_traversal, get_optimising_traversal, get_branch_lengths

__all__ = ["Traversal"]


class Traversal(object):
    def __init__(self, tree):
        self.node_dict = {}
        self.names = {}
        idx = 0
        for node in tree.postorder_node_iter():
            if node is tree.seed_node:
                continue
            self.node_dict[node] = idx
            if node.is_leaf():
                self.names[node.taxon.label] = idx
            idx += 1

        root_children = tree.seed_node.child_nodes()
        if len(root_children) != 2:
            raise ValueError("Traversal needs a tree whose seed node has exactly two children "
                             "(use utils.deepcopy_tree first); found {}".format(len(root_children)))
        self.root_edge = tuple(self.node_dict[n] for n in root_children)
        self.brlens = get_branch_lengths(self.node_dict)

        nleaves = len(self.names)
        self.n_leaves = nleaves
        self.n_nodes = idx
        self.postorder_traversal = get_postorder_traversal(
            tree, np.zeros((max(nleaves - 2, 0), 3), dtype=np.int64), self.node_dict)
        self.optimising_traversal = get_optimising_traversal(
            tree, np.zeros((max(3 * nleaves - 5, 1), 5), dtype=np.int64), self.node_dict)

        self._is_leaf = np.zeros(self.n_nodes, dtype=bool)
        for i in self.names.values():
            self._is_leaf[i] = True

    # ----------------------------------------------------------------------------------
    # derived schedules
    # ----------------------------------------------------------------------------------
    def is_leaf(self, node_id):
        return bool(self._is_leaf[node_id])

    def node_levels(self):
        """level 0 = leaves; an internal node sits one above its deeper child."""
        level = np.zeros(self.n_nodes, dtype=np.int64)
        for par, c1, c2 in self.postorder_traversal:
            level[par] = max(level[c1], level[c2]) + 1
        return level

    def level_order(self):
        """
        -> (rows, offsets): ``rows`` are the postorder rows stably re-sorted by level,
        ``offsets[l]:offsets[l+1]`` is the slice holding level ``l+1``.
        """
        rows = self.postorder_traversal
        if len(rows) == 0:
            return rows.copy(), np.zeros(1, dtype=np.int64)
        level = self.node_levels()[rows[:, 0]]
        order = np.argsort(level, kind="stable")
        counts = np.bincount(level[order] - 1)
        offsets = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
        return rows[order].copy(), offsets

    def locality_order(self):
        """A postorder of the same rows in which the heavier child subtree is finished first."""
        rows = self.postorder_traversal
        n = len(rows)
        if n == 0:
            return rows.copy()
        row_of = {int(r[0]): i for i, r in enumerate(rows)}
        weight = np.zeros(self.n_nodes, dtype=np.int64)       # internal nodes in the subtree
        for par, c1, c2 in rows:                              # rows are already a postorder
            weight[par] = 1 + weight[c1] + weight[c2]
        has_parent = np.zeros(self.n_nodes, dtype=bool)
        has_parent[rows[:, 1]] = True
        has_parent[rows[:, 2]] = True
        tops = [int(r[0]) for r in rows if not has_parent[r[0]]]
        tops.sort(key=lambda v: -weight[v])
        out = []
        for top in tops:
            stack = [(top, False)]
            while stack:
                node, done = stack.pop()
                i = row_of.get(node)
                if i is None:
                    continue
                if done:
                    out.append(i)
                    continue
                _, c1, c2 = rows[i]
                first, second = (c1, c2) if weight[c1] >= weight[c2] else (c2, c1)
                stack.append((node, True))
                stack.append((int(second), False))
                stack.append((int(first), False))
        return rows[np.asarray(out, dtype=np.int64)].copy()

    def edges(self):
        """
        All 2N-3 edges of the unrooted tree as ``(lower_id, upper_id)`` pairs where
        ``lower`` is the node whose down-partial faces away from the root edge; the root
        edge itself comes first.
        """
        out = [tuple(self.root_edge)]
        for par, c1, c2 in self.postorder_traversal[::-1]:
            out.append((int(c1), int(par)))
            out.append((int(c2), int(par)))
        return out

    def preorder_rows(self):
        """
        Rows ``[NOD, PAR, SIB]`` for every non-root-child node, parents before children:
        the up-partial of NOD (everything outside NOD's subtree, seen from PAR's end of the
        edge) is built from SIB's down-partial and PAR's own up-partial.  For children of a
        root child the 'parent up-partial' is the other root child's down-partial.
        """
        out = []
        for par, c1, c2 in self.postorder_traversal[::-1]:
            out.append((int(c1), int(par), int(c2)))
            out.append((int(c2), int(par), int(c1)))
        return np.asarray(out, dtype=np.int64).reshape(-1, 3)
