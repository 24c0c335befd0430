"""
The reference's legacy object API for pairwise / small-tree likelihoods - ``Leaf``, ``LnlNode``, ``LnlModel``,
``GammaMixture``, ``OptWrapper``, ``BranchLengthOptimiser``, ``optimise``, ``brent_optimise`` - with the names, call
signatures and return shapes of /root/reference/phylo_utils/likelihood.py:9-299.

In the reference that module is unimportable as shipped: it calls a ``likcalc`` extension whose import is commented
out (likelihood.py:2) and needs dendropy.  Here the four ``likcalc`` entry points it uses are expressed through the
operators of this package (``clv``, ``lnl_branch``, ``lnl_branch_derivs`` - CUDA kernels behind the C ABI), one rate
category per object exactly as the legacy code has it:

    likcalc.likvec_2desc[_scaled](p1, p2, a, b)               -> clv   (numba_likelihood_engine.py:10-46)
    likcalc.sitewise_lik(p, pi, a, b)                         -> lnl_branch         (:60-79)   column 0 = ln f
    likcalc.sitewise_lik_derivs(p, dp, d2p, pi, a, b)         -> lnl_branch_derivs  (:49-57)   [ln f, f'/f, (f''f - f'^2)/f^2]

``STATE_ORDER``: the known answers in the reference's tests/test_likelihood.py (0.0764 / 0.0378 / 0.0011 / 0.0011 and
-3.5371 for partials [1,0,0,0] x [0,1,0,0] under K80) were written when the nucleotide order of the engine was
T, C, A, G, so that states 0 and 1 form a transition pair; today's model objects are A, C, G, T.
Setting ``STATE_ORDER = "TCAG"`` makes the objects below read and return 4-state vectors in that legacy order; the
default (None) uses the model's own order.
"""
import numpy as np

from . import cuda_likelihood_engine as _engine
from ..utils import setup_logger

logger = setup_logger()
STATE_ORDER = None
_operators = _engine


def use_operators(module):
    """Swap the operator module (``clv``, ``lnl_branch``, ``lnl_branch_derivs``); tests use it to run the same objects on
    the CPU oracle when no GPU is present.  Returns the previous module."""
    global _operators
    previous, _operators = _operators, module
    return previous


def _perm(n_states):
    if STATE_ORDER is None or n_states != 4:
        return None
    return np.array(["ACGT".index(ch) for ch in STATE_ORDER])


def _to_model(v):
    perm = _perm(v.shape[-1])
    if perm is None:
        return v
    out = np.empty_like(v)
    out[..., perm] = v
    return out


def _from_model(v):
    perm = _perm(v.shape[-1])
    return v if perm is None else np.ascontiguousarray(v[..., perm])


def likvec_2desc_scaled(probs1, probs2, partials1, partials2):
    """-> (partials (S, A), per-site natural-log scale buffer (S,))"""
    a, b = _to_model(np.asarray(partials1, dtype=np.double)), _to_model(np.asarray(partials2, dtype=np.double))
    n = a.shape[0]
    scale = np.zeros((n, 1))
    out = _operators.clv(probs1[None], probs2[None], np.ascontiguousarray(a[:, None, :]), np.ascontiguousarray(b[:, None, :]),
                         np.zeros((n, 1)), np.zeros((n, 1)), scale)
    return _from_model(out[:, 0, :]), scale[:, 0]


def likvec_2desc(probs1, probs2, partials1, partials2):
    """Unscaled product: the scaled result multiplied back by exp(scale)."""
    out, scale = likvec_2desc_scaled(probs1, probs2, partials1, partials2)
    return out * np.exp(scale)[:, None]


def sitewise_lik(probs, freqs, partials_a, partials_b):
    a, b = _to_model(np.asarray(partials_a, dtype=np.double)), _to_model(np.asarray(partials_b, dtype=np.double))
    zeros = np.zeros(a.shape[0])
    return np.asarray(_operators.lnl_branch(probs, freqs, a, b, zeros, zeros)).reshape(-1, 1)


def sitewise_lik_derivs(probs, dprobs, d2probs, freqs, partials_a, partials_b):
    a, b = _to_model(np.asarray(partials_a, dtype=np.double)), _to_model(np.asarray(partials_b, dtype=np.double))
    zeros = np.zeros(a.shape[0])
    return np.asarray(_operators.lnl_branch_derivs(np.stack([probs, dprobs, d2probs]), freqs, a, b, zeros, zeros)).reshape(-1, 3)


def _edge_length(node):
    """Branch length above ``node``: this package's trees keep it on the node, dendropy on ``node.edge``."""
    length = getattr(node, "edge_length", None)
    return length if length is not None else node.edge.length


class Leaf(object):
    """Object to store partials at a leaf (likelihood.py:9-18)."""

    def __init__(self, partials):
        self.set_partials(partials)

    def set_partials(self, partials):
        self.partials = np.ascontiguousarray(partials, dtype=np.double)


class LnlNode(object):
    """Partials and transition probabilities of one node, one rate category (likelihood.py:21-83)."""

    def __init__(self, subst_model):
        self.subst_model = subst_model
        self.partials = None
        self.sitewise = None
        self.scale_buffer = None

    def update_transition_probabilities(self, len1, len2):
        self.probs1 = self.subst_model.p(len1)
        self.probs2 = self.subst_model.p(len2)

    def set_partials(self, partials):
        partials = np.asarray(partials, dtype=np.double)
        self.partials = np.ascontiguousarray(partials[np.newaxis] if partials.ndim == 1 else partials, dtype=np.double)

    def compute_partials(self, lnlmodel1, lnlmodel2, scale=True):
        if scale:
            self.partials, self.scale_buffer = likvec_2desc_scaled(self.probs1, self.probs2, lnlmodel1.partials, lnlmodel2.partials)
        else:
            self.partials = likvec_2desc(self.probs1, self.probs2, lnlmodel1.partials, lnlmodel2.partials)

    def compute_edge_sitewise_likelihood(self, lnlmodel, brlen, derivatives=False):
        """Per-site [ln f] or [ln f, f'/f, (f''f - f'^2)/f^2] across the edge to ``lnlmodel`` (likelihood.py:53-68)."""
        probs = self.subst_model.p(brlen)
        if derivatives:
            self.sitewise = sitewise_lik_derivs(probs, self.subst_model.dp_dt(brlen), self.subst_model.d2p_dt2(brlen),
                                                self.subst_model.freqs, self.partials, lnlmodel.partials)
        else:
            self.sitewise = sitewise_lik(probs, self.subst_model.freqs, self.partials, lnlmodel.partials)

    def compute_likelihood(self, lnlmodel, brlen, derivatives=False, accumulated_scale_buffer=None):
        self.compute_edge_sitewise_likelihood(lnlmodel, brlen, derivatives)
        swlnl = self.sitewise[:, 0]
        lnl = (swlnl + accumulated_scale_buffer).sum() if accumulated_scale_buffer is not None else swlnl.sum()
        if derivatives:
            return lnl, self.sitewise[:, 1].sum(), self.sitewise[:, 2].sum()
        return lnl


class LnlModel(object):
    """One rate category over a whole tree (likelihood.py:86-137); ``tree`` is any object with the dendropy surface
    used there (``phylo_utils_b200.tree.Tree`` has it)."""

    def __init__(self, subst_model, partials_dict):
        self.leaf_models = {}
        for leafname, partials in partials_dict.items():
            model = LnlNode(subst_model)
            model.set_partials(partials)
            self.leaf_models[leafname] = model
        self.nsites = next(iter(partials_dict.values())).shape[0]
        self.subst_model = subst_model
        self.accumulated_scale_buffer = None

    def set_tree(self, tree):
        self.tree = tree
        for leaf in self.tree.leaf_nodes():
            leaf.model = self.leaf_models[leaf.taxon.label]

    def update_subst_model(self, subst_model):
        self.subst_model = subst_model
        for leaf in self.tree.leaf_nodes():
            leaf.model.subst_model = subst_model

    def run(self, derivatives=False):
        self.accumulated_scale_buffer = np.zeros(self.nsites, dtype=np.double)
        for node in self.tree.postorder_internal_node_iter():
            children = node.child_nodes()
            if node is self.tree.seed_node and len(children) == 2:
                break
            node.model = LnlNode(self.subst_model)
            node.model.update_transition_probabilities(*[_edge_length(ch) for ch in children[:2]])
            node.model.compute_partials(children[0].model, children[1].model, True)
            self.accumulated_scale_buffer += node.model.scale_buffer
        ch1, ch2 = self.tree.seed_node.child_nodes()[:2]
        return ch1.model.compute_likelihood(ch2.model, _edge_length(ch1) + _edge_length(ch2), derivatives,
                                            self.accumulated_scale_buffer)

    def get_sitewise_likelihoods(self):
        ch = self.tree.seed_node.child_nodes()[0]
        scaler = np.zeros_like(ch.model.sitewise)
        scaler[:, 0] = self.accumulated_scale_buffer
        return ch.model.sitewise + scaler


class Mixture(object):
    def mix_likelihoods(self, sw_lnls):
        ma = sw_lnls.max(1)[:, np.newaxis]
        wa = sw_lnls + self.logweights
        return np.log(np.exp(wa - ma).sum(1))[:, np.newaxis] + ma


class GammaMixture(Mixture):
    """``ncat`` LnlModel runners on rate-scaled copies of one tree (likelihood.py:147-193)."""

    def __init__(self, alpha, ncat):
        from ..discrete_gamma import discrete_gamma
        self._discrete_gamma = discrete_gamma
        self.ncat = ncat
        self.rates = discrete_gamma(alpha, ncat)
        self.weights = np.array([1.0 / ncat] * ncat)
        self.logweights = np.log(self.weights)

    def update_alpha(self, alpha):
        self.rates = self._discrete_gamma(alpha, self.ncat)
        self.set_tree(self.tree)

    def update_substitution_model(self, tm):
        for runner in self.runners:
            runner.update_subst_model(tm)

    def init_models(self, tm, partials_dict):
        self.runners = [LnlModel(tm, partials_dict) for _ in range(self.ncat)]

    def set_tree(self, tree):
        """``tree``: a Newick string, as in the reference (likelihood.py:169-175)."""
        from ..tree import parse_newick
        self.tree = tree
        for cat in range(self.ncat):
            t = parse_newick(tree)
            t.resolve_polytomies()
            for nd in t.preorder_node_iter():
                if nd.edge_length is not None:
                    nd.edge_length *= self.rates[cat]
            self.runners[cat].set_tree(t)

    def run(self, derivatives=False):
        for runner in self.runners:
            runner.run(derivatives)

    def get_sitewise_likelihoods(self):
        swlnls = np.empty((self.runners[0].nsites, self.ncat))
        for cat in range(self.ncat):
            swlnls[:, cat] = self.runners[cat].get_sitewise_likelihoods()[:, 0]
        return swlnls

    def get_scale_bufs(self):
        return np.array([model.accumulated_scale_buffer for model in self.runners]).T

    def get_likelihood(self):
        return self.mix_likelihoods(self.get_sitewise_likelihoods()).sum()


class OptWrapper(object):
    """For scipy root finders on dlnL/dt (likelihood.py:196-223)."""

    def __init__(self, tm, partials1, partials2, initial_brlen=1.0):
        self.root = LnlNode(tm)
        self.leaf = Leaf(partials2)
        self.root.set_partials(partials1)
        self.updated = None
        self.update(initial_brlen)

    def update(self, brlen):
        if self.updated != brlen:
            self.updated = brlen
            self.lnl, self.dlnl, self.d2lnl = self.root.compute_likelihood(self.leaf, brlen, derivatives=True)

    def get_dlnl(self, brlen):
        self.update(brlen)
        return self.dlnl

    def get_d2lnl(self, brlen):
        self.update(brlen)
        return self.d2lnl

    def __str__(self):
        return 'Branch length={}, Variance={}, Likelihood+derivatives = {} {} {}'.format(
            self.updated, -1 / self.d2lnl, self.lnl, self.dlnl, self.d2lnl)


def optimise(likelihood, partials_a, partials_b, min_brlen=0.00001, max_brlen=10, verbose=True):
    """ML distance between two sets of partials: root of dlnL/dt by Brent's method (likelihood.py:226-237)."""
    from scipy.optimize import brenth
    wrapper = OptWrapper(likelihood, partials_a, partials_b, (min_brlen + max_brlen) / 2.)
    n = brenth(wrapper.get_dlnl, min_brlen, max_brlen)
    if verbose:
        logger.info(wrapper)
    return n, -1 / wrapper.get_d2lnl(n)


class BranchLengthOptimiser(object):
    """likelihood.py:240-281"""

    def __init__(self, node1, node2, initial_brlen=1.0):
        self.root = node1
        self.desc = node2
        self.updated = None
        self.__call__(initial_brlen)

    def __call__(self, brlen):
        if self.updated != brlen:
            self.updated = brlen
            self.lnl, self.dlnl, self.d2lnl = self.root.compute_likelihood(self.desc, brlen, derivatives=True)
        return self.lnl, self.dlnl, self.d2lnl

    def get_lnl(self, brlen):
        return self.__call__(brlen)[0]

    def get_dlnl(self, brlen):
        return np.array([self.__call__(brlen)[1]])

    def get_d2lnl(self, brlen):
        return np.array([self.__call__(brlen)[2]])

    def get_negative_lnl(self, brlen):
        return -self.__call__(max(0, brlen))[0]

    def get_negative_dlnl(self, brlen):
        return -self.__call__(max(0, brlen))[1]

    def get_negative_d2lnl(self, brlen):
        return -self.__call__(max(0, brlen))[2]

    def __str__(self):
        return 'Branch length={}, Variance={}, Likelihood+derivatives = {} {} {}'.format(
            self.updated, -1 / self.d2lnl, self.lnl, self.dlnl, self.d2lnl)


def brent_optimise(node1, node2, min_brlen=0.00001, max_brlen=10, verbose=True):
    """likelihood.py:283-293"""
    from scipy.optimize import minimize_scalar
    wrapper = BranchLengthOptimiser(node1, node2, (min_brlen + max_brlen) / 2.)
    n = minimize_scalar(lambda x: -wrapper(x)[0], method='brent', bracket=(min_brlen, max_brlen))['x']
    if verbose:
        logger.info(wrapper)
    return n, -1 / wrapper.get_d2lnl(n)
