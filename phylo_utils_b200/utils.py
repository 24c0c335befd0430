"""
Tree bookkeeping helpers with the reference's names and table semantics
(``/root/reference/phylo_utils/utils.py:47-213``), re-written iteratively so that ladder
trees with thousands of taxa work, and typed ``int64`` (the reference asks for the
long-removed ``np.int``).

Everything here is host-side, one-off work per topology; the tables it produces are
what :mod:`phylo_utils_b200.traversal` turns into the device schedule.
"""
import copy
import logging

import numpy as np

__all__ = [
    "setup_logger", "deepcopy_tree", "get_sibling", "get_grandparent", "get_node_dict",
    "get_postorder_traversal", "get_optimising_traversal", "get_branch_lengths", "BranchLengths",
]


def setup_logger(name="phylo_utils_b200"):
    """Module logger (reference: utils.py:7-17); handlers are only attached once."""
    logger = logging.getLogger(name)
    if not logger.handlers:
        handler = logging.StreamHandler()
        handler.setFormatter(logging.Formatter("%(asctime)s - %(name)s - %(levelname)s - %(message)s"))
        logger.addHandler(handler)
        logger.setLevel(logging.INFO)
    return logger


def deepcopy_tree(tree):
    """Private, unrooted-at-a-trifurcation-then-binarised copy (reference: utils.py:114-118)."""
    clone = copy.deepcopy(tree)
    clone.deroot()
    clone.resolve_polytomies()
    return clone


def get_sibling(node):
    """The other child of ``node``'s parent (reference: utils.py:121-124)."""
    for other in node.parent_node.child_node_iter():
        if other is not node:
            return other
    return None


def get_grandparent(tree, node):
    """Reference: utils.py:50-54 - across the seed node the 'grandparent' is the parent's sibling."""
    above = node.parent_node.parent_node
    if above is tree.seed_node:
        return get_sibling(node.parent_node)
    return above


def get_node_dict(tree):
    """leaves first, then internal nodes in postorder, seed excluded (reference: utils.py:81-84)."""
    order = list(tree.leaf_node_iter()) + list(tree.postorder_internal_node_iter(exclude_seed_node=True))
    return {nd: i for i, nd in enumerate(order)}


def get_postorder_traversal(tree, descriptor, node_dict):
    """
    Fill ``descriptor`` (N-2, 3) with one ``[PAR, CH1, CH2]`` row per internal node, in
    postorder, the seed node left out (reference: utils.py:127-134).
    """
    row = 0
    for node in tree.postorder_internal_node_iter(exclude_seed_node=True):
        kids = node.child_nodes()
        if len(kids) != 2:
            raise ValueError("tree is not strictly bifurcating below the seed node")
        descriptor[row, 0] = node_dict[node]
        descriptor[row, 1] = node_dict[kids[0]]
        descriptor[row, 2] = node_dict[kids[1]]
        row += 1
    return descriptor


def get_optimising_traversal(tree, descriptor, node_dict):
    """
    Re-rooting sweep table (3N-5, 5) (reference: utils.py:137-188).

    Row 0 ``[-1,-1,-1,LEFT,RIGHT]``: optimise the root edge.  On first reaching a node NOD
    that is not a root child: ``[PAR,SIB,GPA,NOD,PAR]`` = rebuild PAR's partial from SIB
    and GPA so that it faces NOD, then optimise edge NOD-PAR.  After both children of an
    internal NOD have been handled: ``[NOD,CH1,CH2,-1,-1]`` = point NOD's partial back at
    the root.  Written as an explicit-stack DFS instead of the reference's recursion.
    """
    left, right = tree.seed_node.child_nodes()
    descriptor[0] = (-1, -1, -1, node_dict[left], node_dict[right])
    row = 1
    for top in (left, right):
        stack = [(top, False)]
        while stack:
            node, done = stack.pop()
            if done:
                c1, c2 = node.child_nodes()
                descriptor[row] = (node_dict[node], node_dict[c1], node_dict[c2], -1, -1)
                row += 1
                continue
            if node is not left and node is not right:
                par = node.parent_node
                if par is left:
                    gpa = right
                elif par is right:
                    gpa = left
                else:
                    gpa = par.parent_node
                sib = get_sibling(node)
                descriptor[row] = (node_dict[par], node_dict[sib], node_dict[gpa],
                                   node_dict[node], node_dict[par])
                row += 1
            if not node.is_leaf():
                c1, c2 = node.child_nodes()
                stack.append((node, True))
                stack.append((c2, False))
                stack.append((c1, False))
    return descriptor


class BranchLengths(dict):
    """
    ``{(i, j): length}`` with order-insensitive lookup (reference: utils.py:191-199).

    The lengths themselves live in one dense float64 array (``dict`` maps a key to its slot), because the
    device path moves ALL branch lengths of a tree on every evaluation: ``slots(keys)`` resolves keys once
    per schedule, ``gather(slots)`` / ``scatter(slots, values)`` are single numpy indexing operations where a
    dictionary walk of 2 000 keys would cost as much as a tenth of an evaluation on eight GPUs.  Seen through
    ``[]``, ``in``, ``get``, ``items``, ``values``, ``keys``, ``len`` it is the reference's dictionary.
    """

    def __init__(self, *args, **kwargs):
        dict.__init__(self)
        self._dense = np.empty(16, dtype=np.double)
        for key, value in dict(*args, **kwargs).items():
            self[key] = value

    # -- dictionary face --------------------------------------------------------------------
    def canonical_key(self, key):
        if dict.__contains__(self, key):
            return key
        flipped = tuple(key)[::-1]
        if dict.__contains__(self, flipped):
            return flipped
        raise KeyError(key)

    def __getitem__(self, key):
        return float(self._dense[dict.__getitem__(self, self.canonical_key(key))])

    def __setitem__(self, key, value):
        try:
            slot = dict.__getitem__(self, self.canonical_key(key))
        except KeyError:
            slot = dict.__len__(self)
            if slot == self._dense.shape[0]:
                self._dense = np.concatenate([self._dense, np.empty(slot, dtype=np.double)])
            dict.__setitem__(self, tuple(key), slot)
        self._dense[slot] = value

    def __contains__(self, key):
        return dict.__contains__(self, key) or dict.__contains__(self, tuple(key)[::-1])

    def get(self, key, default=None):
        try:
            return self[key]
        except KeyError:
            return default

    def values(self):
        return [float(v) for v in self._dense[:dict.__len__(self)]]     # slots are handed out in insertion order

    def items(self):
        return list(zip(dict.keys(self), self.values()))

    def __eq__(self, other):
        return dict(self.items()) == (dict(other.items()) if isinstance(other, BranchLengths) else other)

    def __ne__(self, other):
        return not self == other

    __hash__ = None

    def __repr__(self):
        return "BranchLengths({!r})".format(dict(self.items()))

    def __delitem__(self, key):
        raise TypeError("edges cannot be removed from a BranchLengths table")

    pop = popitem = clear = setdefault = lambda self, *a, **k: BranchLengths.__delitem__(self, None)

    def update(self, *args, **kwargs):
        for key, value in dict(*args, **kwargs).items():
            self[key] = value

    def copy(self):
        return BranchLengths(self.items())

    def __deepcopy__(self, memo):
        return self.copy()

    def __reduce__(self):
        return (BranchLengths, (self.items(),))

    # -- array face (what TreeModel and the optimiser use) ---------------------------------------
    def slots(self, keys):
        """Slot index of every key (either orientation) - resolve once, then ``gather`` / ``scatter``."""
        return np.fromiter((dict.__getitem__(self, self.canonical_key(k)) for k in keys), dtype=np.intp, count=len(keys))

    def gather(self, slots):
        return self._dense[slots]

    def scatter(self, slots, values):
        self._dense[slots] = values


def get_branch_lengths(node_dict):
    """
    One entry per edge of the unrooted tree, keyed by the sorted node-id pair
    (reference: utils.py:202-213).  The two root children share a single entry whose
    length is the larger of their two edge lengths - after ``deepcopy_tree`` one of the
    two is the zero-length edge introduced by resolving the root trifurcation.
    """
    brlens = BranchLengths()
    for node, idx in node_dict.items():
        parent = node.parent_node
        if parent.parent_node is None:          # parent is the seed node
            other = get_sibling(node)
            length = max(ch.edge_length for ch in parent.child_nodes())
        else:
            other = parent
            length = node.edge_length
        key = (idx, node_dict[other])
        brlens[key if key[0] <= key[1] else key[::-1]] = length
    return brlens
