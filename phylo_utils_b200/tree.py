"""
Minimal rooted-tree container with the subset of the dendropy interface that the
likelihood path touches.

The reference never constructs trees itself: it receives a dendropy ``Tree`` and
reads it through a handful of attributes (``/root/reference/phylo_utils/traversal.py:16-24``,
``/root/reference/phylo_utils/utils.py:114-134,202-213``):

    tree.seed_node, tree.postorder_node_iter(), tree.postorder_internal_node_iter(exclude_seed_node=),
    tree.deroot(), tree.resolve_polytomies(),
    node.child_nodes(), node.child_node_iter(), node.is_leaf(), node.parent_node,
    node.edge_length, node.taxon.label, node.preorder_iter()

dendropy is not part of this image, so this module supplies objects with exactly that
surface.  A real dendropy tree can be handed to :class:`phylo_utils_b200.traversal.Traversal`
just as well - nothing in the package type-checks the tree.

Also here: a Newick reader/writer and the seeded random-topology generator used by
the benchmarks ("repeatedly joining two uniformly chosen live subtrees").
"""
from __future__ import annotations

import numpy as np

__all__ = ["Taxon", "Node", "Tree", "parse_newick", "random_tree", "balanced_tree", "caterpillar_tree"]


class Taxon(object):
    __slots__ = ("label",)

    def __init__(self, label):
        self.label = label

    def __repr__(self):
        return "Taxon({!r})".format(self.label)


class Node(object):
    """A tree node; ``edge_length`` is the length of the edge to the parent."""

    def __init__(self, label=None, edge_length=None):
        self.taxon = Taxon(label) if label is not None else None
        self.edge_length = edge_length
        self.parent_node = None
        self._children = []

    # ---- dendropy-compatible surface -------------------------------------------------
    def child_nodes(self):
        return list(self._children)

    def child_node_iter(self):
        return iter(list(self._children))

    def is_leaf(self):
        return not self._children

    def is_internal(self):
        return bool(self._children)

    def add_child(self, node, pos=None):
        node.parent_node = self
        if pos is None:
            self._children.append(node)
        else:
            self._children.insert(pos, node)
        return node

    def remove_child(self, node):
        self._children.remove(node)
        node.parent_node = None
        return node

    def preorder_iter(self):
        stack = [self]
        while stack:
            nd = stack.pop()
            yield nd
            stack.extend(reversed(nd._children))

    def postorder_iter(self):
        # iterative so that caterpillar trees with thousands of taxa do not hit the recursion limit
        stack = [(self, 0)]
        while stack:
            nd, i = stack.pop()
            if i < len(nd._children):
                stack.append((nd, i + 1))
                stack.append((nd._children[i], 0))
            else:
                yield nd

    def leaf_iter(self):
        for nd in self.preorder_iter():
            if not nd._children:
                yield nd

    @property
    def label(self):
        return self.taxon.label if self.taxon is not None else None

    def __repr__(self):
        return "<Node {} len={}>".format(self.label, self.edge_length)


class Tree(object):
    def __init__(self, seed_node=None):
        self.seed_node = seed_node if seed_node is not None else Node()
        self.is_rooted = True

    # ---- iteration -----------------------------------------------------------------
    def postorder_node_iter(self):
        return self.seed_node.postorder_iter()

    def preorder_node_iter(self):
        return self.seed_node.preorder_iter()

    def postorder_internal_node_iter(self, exclude_seed_node=False):
        for nd in self.seed_node.postorder_iter():
            if nd._children and not (exclude_seed_node and nd is self.seed_node):
                yield nd

    def leaf_node_iter(self):
        return self.seed_node.leaf_iter()

    def leaf_nodes(self):
        return list(self.leaf_node_iter())

    def nodes(self):
        return list(self.preorder_node_iter())

    def __len__(self):
        return sum(1 for _ in self.leaf_node_iter())

    # ---- copying ----------------------------------------------------------------------
    def __deepcopy__(self, memo):
        # iterative: copy.deepcopy's default recursion overflows on ladder-like trees
        clone_of = {}
        for nd in self.seed_node.preorder_iter():
            dup = Node(label=nd.label, edge_length=nd.edge_length)
            clone_of[nd] = dup
            if nd.parent_node is not None and nd is not self.seed_node:
                clone_of[nd.parent_node].add_child(dup)
        twin = Tree(clone_of[self.seed_node])
        twin.is_rooted = self.is_rooted
        return twin

    # ---- structural edits used by utils.deepcopy_tree --------------------------------
    def deroot(self):
        """
        Turn a bifurcating root into a trifurcation by dissolving one root child.

        Same convention as dendropy: if the second root child is internal it is the one
        dissolved (its edge length is added to the first child's edge and its children
        are spliced into the root at its position); otherwise the first one is.  A root
        with two leaf children, or with != 2 children, is left alone.
        """
        root = self.seed_node
        kids = root._children
        if len(kids) != 2:
            self.is_rooted = False
            return root
        if len(kids[1]._children) >= 2:
            keep, gone = kids[0], kids[1]
        elif len(kids[0]._children) >= 2:
            gone, keep = kids[0], kids[1]
        else:
            return root
        if gone.edge_length is not None:
            keep.edge_length = (keep.edge_length or 0.0) + gone.edge_length
        pos = kids.index(gone)
        root.remove_child(gone)
        for off, ch in enumerate(list(gone._children)):
            gone.remove_child(ch)
            root.add_child(ch, pos + off)
        self.is_rooted = False
        return root

    def resolve_polytomies(self, limit=2):
        """
        Resolve every node with more than ``limit`` children into a ladder of
        zero-length edges, deterministically: the first ``limit - 1`` children stay put,
        the remaining ones are pushed below a new zero-length internal child which is
        resolved in turn.
        """
        pending = [nd for nd in self.postorder_node_iter() if len(nd._children) > limit]
        while pending:
            nd = pending.pop()
            extra = nd._children[limit - 1:]
            for ch in extra:
                nd.remove_child(ch)
            joint = Node(edge_length=0.0)
            nd.add_child(joint)
            for ch in extra:
                joint.add_child(ch)
            if len(joint._children) > limit:
                pending.append(joint)

    # ---- I/O --------------------------------------------------------------------------
    def as_newick(self, precision=17):
        fmt = "{:." + str(precision) + "g}"

        out = []
        stack = [(self.seed_node, 0)]
        while stack:
            nd, i = stack.pop()
            if not nd._children:
                out.append(_quote(nd.label))
                if nd.edge_length is not None:
                    out.append(":" + fmt.format(nd.edge_length))
                continue
            if i == 0:
                out.append("(")
            elif i < len(nd._children):
                out.append(",")
            if i < len(nd._children):
                stack.append((nd, i + 1))
                stack.append((nd._children[i], 0))
            else:
                out.append(")")
                if nd.label is not None:
                    out.append(_quote(nd.label))
                if nd.edge_length is not None and nd is not self.seed_node:
                    out.append(":" + fmt.format(nd.edge_length))
        return "".join(out) + ";"

    def __str__(self):
        return self.as_newick(6)

    @classmethod
    def get_from_string(cls, text, schema="newick", **_):
        if schema != "newick":
            raise ValueError("only the newick schema is supported")
        return parse_newick(text)

    @classmethod
    def get_from_path(cls, path, schema="newick", **_):
        with open(path) as fh:
            return cls.get_from_string(fh.read(), schema)


def _quote(label):
    if label is None:
        return ""
    if any(c in label for c in " ()[]':;,"):
        return "'" + label.replace("'", "''") + "'"
    return label


def parse_newick(text):
    """Parse one Newick string (quoted labels, [comments], branch lengths, internal labels)."""
    s = text.strip()
    n = len(s)
    i = 0
    root = Node()
    cur = root
    expect_label_for = root   # node that a following label / length belongs to
    depth = 0
    seen_any = False

    def skip_ws_comments(i):
        while i < n:
            c = s[i]
            if c.isspace():
                i += 1
            elif c == "[":
                j = s.find("]", i)
                if j < 0:
                    raise ValueError("unterminated [comment] in newick string")
                i = j + 1
            else:
                break
        return i

    while True:
        i = skip_ws_comments(i)
        if i >= n:
            break
        c = s[i]
        if c == "(":
            child = Node()
            if seen_any and cur is root and depth == 0 and root._children:
                raise ValueError("unexpected '(' after the root clade")
            cur.add_child(child)
            cur = child
            expect_label_for = child
            depth += 1
            seen_any = True
            i += 1
        elif c == ",":
            if depth == 0:
                raise ValueError("',' outside of any clade")
            sib = Node()
            cur.parent_node.add_child(sib)
            cur = sib
            expect_label_for = sib
            i += 1
        elif c == ")":
            if depth == 0:
                raise ValueError("unbalanced ')' in newick string")
            cur = cur.parent_node
            expect_label_for = cur
            depth -= 1
            i += 1
        elif c == ";":
            i += 1
            break
        elif c == ":":
            i = skip_ws_comments(i + 1)
            j = i
            while j < n and s[j] not in ",();[ \t\r\n":
                j += 1
            expect_label_for.edge_length = float(s[i:j])
            i = j
        else:
            if c == "'":
                j = i + 1
                buf = []
                while True:
                    if j >= n:
                        raise ValueError("unterminated quoted label")
                    if s[j] == "'":
                        if j + 1 < n and s[j + 1] == "'":
                            buf.append("'")
                            j += 2
                            continue
                        break
                    buf.append(s[j])
                    j += 1
                label = "".join(buf)
                i = j + 1
            else:
                j = i
                while j < n and s[j] not in ":,();[" and not s[j].isspace():
                    j += 1
                label = s[i:j]
                i = j
            expect_label_for.taxon = Taxon(label)
            seen_any = True
    if depth != 0:
        raise ValueError("unbalanced parentheses in newick string")
    if not root._children and root.taxon is None:
        raise ValueError("empty newick string")
    for nd in root.preorder_iter():
        if nd._children and nd.taxon is not None and nd.taxon.label == "":
            nd.taxon = None
    return Tree(root)


# --------------------------------------------------------------------------------------
# synthetic topologies for tests and benchmarks
# --------------------------------------------------------------------------------------
def _taxon_names(n):
    width = len(str(n - 1))
    return ["t{:0{w}d}".format(i, w=width) for i in range(n)]


def random_tree(n_taxa, rng=None, min_len=0.01, max_len=0.3):
    """
    Random rooted binary topology: start from ``n_taxa`` single-leaf subtrees and keep
    joining two uniformly chosen live subtrees.  Branch lengths ~ U(min_len, max_len).
    """
    if n_taxa < 2:
        raise ValueError("need at least two taxa")
    rng = np.random.default_rng(rng)
    live = [Node(label=name) for name in _taxon_names(n_taxa)]
    while len(live) > 1:
        i, j = rng.choice(len(live), size=2, replace=False)
        a, b = live[i], live[j]
        par = Node()
        par.add_child(a)
        par.add_child(b)
        for k in sorted((int(i), int(j)), reverse=True):
            live.pop(k)
        live.append(par)
    tree = Tree(live[0])
    for nd in tree.preorder_node_iter():
        if nd is not tree.seed_node:
            nd.edge_length = float(rng.uniform(min_len, max_len))
    return tree


def balanced_tree(n_taxa, rng=None, min_len=0.01, max_len=0.3):
    """As balanced as the taxon count allows (pairs neighbours level by level)."""
    rng = np.random.default_rng(rng)
    level = [Node(label=name) for name in _taxon_names(n_taxa)]
    while len(level) > 1:
        nxt = []
        for k in range(0, len(level) - 1, 2):
            par = Node()
            par.add_child(level[k])
            par.add_child(level[k + 1])
            nxt.append(par)
        if len(level) % 2:
            nxt.append(level[-1])
        level = nxt
    tree = Tree(level[0])
    for nd in tree.preorder_node_iter():
        if nd is not tree.seed_node:
            nd.edge_length = float(rng.uniform(min_len, max_len))
    return tree


def caterpillar_tree(n_taxa, rng=None, min_len=0.01, max_len=0.3):
    """Fully unbalanced ladder: depth n_taxa - 1, one internal node per level."""
    rng = np.random.default_rng(rng)
    names = _taxon_names(n_taxa)
    cur = Node(label=names[0])
    for name in names[1:]:
        par = Node()
        par.add_child(cur)
        par.add_child(Node(label=name))
        cur = par
    tree = Tree(cur)
    for nd in tree.preorder_node_iter():
        if nd is not tree.seed_node:
            nd.edge_length = float(rng.uniform(min_len, max_len))
    return tree
