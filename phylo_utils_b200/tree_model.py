"""
TreeModel - the reference's orchestration object, re-backed by the GPU engine.

Same methods and attributes as /root/reference/phylo_utils/tree_model.py:12-217:

    set_alignment, get_empirical_freqs, set_substitution_model, set_rate_model, set_tree,
    set_ascertainment_bias_correction, initialise, compute_partials, compute_partials_at_edge,
    compute_likelihood_at_edge;  attributes alignment, siteweights, inverse_index, names,
    traversal, partials, scale, root_partials, root_scale.

What differs underneath:

* tips live on the device as one uint8 state-set code per (taxon, pattern); ``alignment`` /
  ``partials`` / ``scale`` are materialised as numpy arrays only when somebody reads them
* all 2(N-2) transition matrices x K categories come from one kernel launch, not from
  2(N-2) ``model.p`` calls
* the post-order loop is one launch (pattern-tile resident) or one launch per tree level
* scalers are one cumulative binary exponent per pattern; ``scale[node]`` reports it as a
  natural log replicated over the categories, so ``partials * exp(scale)`` is comparable with the
  reference's ``partials * exp(scale)``
* ``compute_likelihood_at_edge`` runs root combine + lnl_node + mixture + log on the device

Extras for the derivative path (no counterpart in TreeModel, composed from lnl_branch_derivs):
``compute_up_partials`` and ``edge_derivatives``.
"""
import numpy as np
from scipy.special import logsumexp

from . import _lib
from .alignment.alignment import alignment_to_codes, invariant_sites
from .engine import LikelihoodEngine
from .traversal import Traversal
from .utils import deepcopy_tree, setup_logger

logger = setup_logger()

# below this many patterns the pattern axis alone cannot occupy the GPU: schedule level by level
_TILE_MODE_MIN_PATTERNS = 16384


class PendingLnl(object):
    """One pipelined host-fed evaluation in flight (``TreeModel.lnl_from_host_submit``)."""

    def __init__(self, model, slot, reduce=None):
        self.model, self.slot, self._reduce, self._posted, self._value = model, slot, reduce, False, None
        self.check = None     # optional: validates / converts the value when it arrives

    def post(self):
        """Enqueue (not wait for) what follows the walk: the optional reduction over ranks, then the copy to the host."""
        if not self._posted:
            if self._reduce is not None:
                self._reduce(self.model.engine.result_tensor(1, offset=self.slot))
            self.model.engine.result_post(self.slot)
            self._posted = True

    def result(self):
        if self._value is None:
            self.post()
            self._value = self.model.engine.result_wait(self.slot)
            if self.check is not None:
                self._value = self.check(self._value)
        return self._value


class TreeModel(object):
    alignment_codes = None
    ascbias = False

    def __init__(self, device=0, up_partials=False, mode="auto", store_partials=True):
        """
        mode: "auto" | "tile" | "level" | "resident" - how the post-order rows are walked on the device.
        store_partials=False builds an lnL-only model: no per-node partials are kept in HBM (``partials`` /
        ``scale`` / derivatives are unavailable), every likelihood call runs the operand-resident kernel
        (4-state models only).
        """
        self.device = device
        self.want_up_partials = up_partials
        self.mode = mode
        self.store_partials = store_partials
        self.engine = None
        self.substitution_model = None
        self.rate_model = None
        self.traversal = None
        self.tree = None
        self._lut = None
        self.siteweights = None
        self.inverse_index = None
        self.names = None
        self._rows = None
        self._n_dummy = 0

    # ------------------------------------------------------------------------------------------
    # inputs (reference: tree_model.py:42-98)
    # ------------------------------------------------------------------------------------------
    def set_alignment(self, alignment, alphabet, compress=True, compress_on_gpu=True):
        """``alignment``: iterable of records with ``.name`` and ``.seq`` (Biopython alignment or alignment.SeqRecord list).
        Character look-up and site-pattern compression run on this model's GPU (csrc/compress.cu, bit-identical to the
        reference's np.unique); ``compress_on_gpu=False`` keeps them on the host."""
        codes, lut, sw, ii, names = alignment_to_codes(alignment, alphabet, compress,
                                                       device=self.device if (compress and compress_on_gpu) else None)
        self.set_tip_codes(codes, lut, names, sw, ii)

    def set_tip_codes(self, codes, lut, names, siteweights=None, inverse_index=None):
        """Direct entry for already-encoded data: ``codes`` uint8 (ntax, npat) indexing ``lut`` (ncodes, A)."""
        self.alignment_codes = codes
        self._lut = np.ascontiguousarray(lut, dtype=np.double)
        npat = codes.shape[1]
        self.siteweights = np.ones(npat, dtype=np.int64) if siteweights is None else np.asarray(siteweights, dtype=np.int64)
        self.inverse_index = np.arange(npat, dtype=np.int64) if inverse_index is None else np.asarray(inverse_index, dtype=np.int64)
        self.names = dict(names)
        self.engine = None

    @property
    def alignment(self):
        """(ntax, npat, A) float array, as the reference stores it (materialised on demand)."""
        if self.alignment_codes is None:
            return None
        codes = self.alignment_codes
        if hasattr(codes, "cpu"):
            codes = codes.cpu().numpy()
        return np.ascontiguousarray(self._lut[codes])

    def get_empirical_freqs(self, pseudocount=None, include_ambiguous=False):
        if self.alignment_codes is None:
            logger.error("No alignment has been set")
            return 0
        if include_ambiguous:
            logger.warning("Not implemented")
        counts = (self.alignment * self.siteweights[np.newaxis, :, np.newaxis]).sum((0, 1))
        if pseudocount is not None:
            try:
                counts += np.array(pseudocount)
            except TypeError:
                logger.warning("Pseudocount {} caused Type error. Carrying on without pseudocount.".format(pseudocount))
            except ValueError:
                logger.warning("Pseudocount {} caused Value error (probably the wrong length). "
                               "Carrying on without pseudocount.".format(pseudocount))
        return counts / counts.sum()

    def set_substitution_model(self, model):
        self.substitution_model = model
        if self.engine is not None:
            self._upload_model()

    def set_rate_model(self, rate_model):
        if self.engine is not None and self.rate_model is not None and rate_model.ncat != self.rate_model.ncat:
            self.engine = None          # category count is baked into the device layout
        self.rate_model = rate_model
        if self.engine is not None:
            self._upload_model()

    def set_tree(self, dpytree):
        self.tree = deepcopy_tree(dpytree)
        self.traversal = Traversal(self.tree)
        self.engine = None

    def set_ascertainment_bias_correction(self):
        """Lewis (2001) correction with one dummy constant pattern per state (reference: tree_model.py:92-98)."""
        if np.any(invariant_sites(self.alignment)):
            logger.warning("Using Lewis ascertainment bias correction on an alignment with invariant sites!")
        self.ascbias = True
        self.engine = None

    # ------------------------------------------------------------------------------------------
    # device set-up (reference: initialise, tree_model.py:101-158)
    # ------------------------------------------------------------------------------------------
    def _tip_rows(self):
        missing = [name for name in self.traversal.names if name not in self.names]
        if missing:
            raise ValueError("taxa in the tree but not in the alignment: {}".format(missing[:5]))
        order = sorted(self.traversal.names, key=lambda nm: self.names[nm])
        rows = [self.names[nm] for nm in order]
        nodes = [self.traversal.names[nm] for nm in order]
        return np.asarray(rows), np.asarray(nodes, dtype=np.int32)

    @property
    def tip_row_order(self):
        """Alignment row (index into ``alignment_codes``) of every device tip row: the row order of the host codes
        ``lnl_from_host_codes`` expects."""
        return self._tip_rows()[0]

    def _choose_mode(self, n_patterns):
        if self.mode == "tile":
            return _lib.PHB_MODE_TILE
        if self.mode == "level":
            return _lib.PHB_MODE_LEVEL
        if self.mode == "resident" or not self.store_partials:
            return _lib.PHB_MODE_RESIDENT
        if n_patterns < _TILE_MODE_MIN_PATTERNS:
            return _lib.PHB_MODE_LEVEL
        return _lib.PHB_MODE_AUTO       # library picks: operand-resident walk for 4-state models, tile walk otherwise

    def initialise(self):
        """Allocate device storage, upload tips / model / schedule, run one post-order pass."""
        if self.alignment_codes is None or self.traversal is None:
            raise ValueError("alignment and tree must be set before initialise()")
        if self.substitution_model is None or self.rate_model is None:
            raise ValueError("substitution model and rate model must be set before initialise()")
        codes, lut = self.alignment_codes, self._lut
        weights = self.siteweights
        n_states = lut.shape[1]
        rows_in_aln, tip_nodes = self._tip_rows()
        on_device = hasattr(codes, "data_ptr")
        if on_device:
            if len(rows_in_aln) != codes.shape[0] or np.any(rows_in_aln != np.arange(codes.shape[0])):
                codes = codes[rows_in_aln.tolist()].contiguous()
        else:
            if len(rows_in_aln) != codes.shape[0] or np.any(rows_in_aln != np.arange(codes.shape[0])):
                codes = np.ascontiguousarray(codes[rows_in_aln])
        self._n_dummy = 0
        if self.ascbias:
            if on_device:
                raise ValueError("ascertainment-bias correction needs host-resident codes")
            # one constant pattern per state, appended after the real patterns (tree_model.py:151-156);
            # weight 0 keeps them out of the total
            lut = lut.copy()
            single = []
            for st in range(n_states):
                onehot = np.zeros(n_states)
                onehot[st] = 1.0
                hit = np.flatnonzero((lut == onehot).all(axis=1))
                if hit.size:
                    single.append(int(hit[0]))
                else:
                    lut = np.vstack([lut, onehot])
                    single.append(lut.shape[0] - 1)
            dummy = np.repeat(np.asarray(single, dtype=np.uint8)[None, :], codes.shape[0], axis=0)
            codes = np.ascontiguousarray(np.hstack([codes, dummy]))
            weights = np.concatenate([weights, np.zeros(n_states, dtype=np.int64)])
            self._n_dummy = n_states
        n_tips, n_patterns = codes.shape
        self._mode = self._choose_mode(n_patterns)
        if not self.store_partials and (self.ascbias or self.want_up_partials):
            raise ValueError("store_partials=False supports plain likelihood evaluation only")
        self.engine = LikelihoodEngine(n_tips, n_patterns, self.rate_model.ncat, n_states, device=self.device,
                                       up_partials=self.want_up_partials, store_partials=self.store_partials)
        self.engine.set_tips(codes, lut, tip_nodes)
        self.engine.set_pattern_weights(weights)
        if self._mode == _lib.PHB_MODE_LEVEL:
            rows, offsets = self.traversal.level_order()
            self.engine.set_schedule(rows, offsets)
        else:
            rows = self.traversal.locality_order()
            self.engine.set_schedule(rows)
        self._rows = rows
        self._upload_model()
        self.compute_partials()

    def _model_signature(self):
        """Everything the device holds of the two model objects, as bytes: a few hundred bytes for a nucleotide
        model, 60 kB for a codon model - cheap enough to compare on every evaluation."""
        m, r = self.substitution_model, self.rate_model
        parts = [np.asarray(r.rates, dtype=np.double).tobytes(), np.asarray(r.weights, dtype=np.double).tobytes(),
                 np.asarray(m.freqs, dtype=np.double).tobytes()]
        if getattr(m, "has_real_eigensystem", True):
            e = m.eigen
            parts += [np.asarray(e.evals).tobytes(), np.asarray(e.evecs).tobytes(), np.asarray(e.ivecs).tobytes()]
        return b"".join(parts)

    def _upload_model(self, only_if_changed=False):
        m, r = self.substitution_model, self.rate_model
        if m is None or r is None or self.engine is None:
            return
        signature = self._model_signature()
        if only_if_changed and signature == getattr(self, "_uploaded_signature", None) and \
                getattr(self, "_uploaded_to", None) is self.engine:
            return
        self._uploaded_signature, self._uploaded_to = signature, self.engine
        if r.ncat != self.engine.n_cat:
            raise ValueError("rate model has {} categories, device layout was built for {}".format(r.ncat, self.engine.n_cat))
        if m.size != self.engine.n_states:
            raise ValueError("substitution model has {} states, alignment has {}".format(m.size, self.engine.n_states))
        if getattr(m, "has_real_eigensystem", True):
            e = m.eigen
            self.engine.set_model(e.evecs, e.evals, np.ascontiguousarray(e.ivecs), m.freqs, r.rates, r.weights)
        else:
            self.engine.set_mixture(m.freqs, r.rates, r.weights)

    def _row_lengths(self):
        """(n_rows, 2) branch lengths in schedule order.  The dictionary keys of the rows' edges are resolved once
        per schedule: this runs on every evaluation, next to kernels that take a few milliseconds."""
        br = self.traversal.brlens
        cache = getattr(self, "_row_slots", None)
        if cache is None or cache[0] is not self._rows or cache[1] is not br:
            keys = []
            for par, c1, c2 in self._rows:
                keys.append((int(par), int(c1)))
                keys.append((int(par), int(c2)))
            cache = self._row_slots = (self._rows, br, br.slots(keys))
        return br.gather(cache[2]).reshape(-1, 2)

    # ------------------------------------------------------------------------------------------
    # the hot path
    # ------------------------------------------------------------------------------------------
    def compute_partials(self):
        """One post-order traversal over all internal nodes (reference: tree_model.py:160-176)."""
        if self.engine is None:
            raise ValueError("call initialise() first")
        # the reference reads rate_model.rates and calls substitution_model.p on every pass (tree_model.py:166-169), so a
        # model object mutated in place (rate_model.alpha = x) takes effect there; re-upload when anything changed
        self._upload_model(only_if_changed=True)
        lengths = self._row_lengths()
        self.engine.set_edge_lengths(lengths)
        if not self.store_partials:
            return                      # lnL-only: the walk happens inside every likelihood call
        m = self.substitution_model
        if getattr(m, "has_real_eigensystem", True):
            self.engine.build_pmatrices()
        else:
            rates = self.rate_model.rates
            pm = np.empty((len(self._rows), 2, self.engine.n_cat, m.size, m.size))
            for i in range(len(self._rows)):
                pm[i, 0] = m.p(lengths[i, 0], rates)
                pm[i, 1] = m.p(lengths[i, 1], rates)
            self.engine.set_pmatrices(pm)
        self.engine.compute_partials(self._mode)

    def _edge_length(self, node_a, node_b):
        try:
            return self.traversal.brlens[node_a, node_b]
        except KeyError:
            raise ValueError('There is no edge connecting nodes {} and {}'.format(node_a, node_b))

    def _root_pmats(self, length):
        m = self.substitution_model
        if getattr(m, "has_real_eigensystem", True):
            return None
        rates = self.rate_model.rates
        return np.stack([m.p(0, rates), m.p(length, rates)])

    def compute_partials_at_edge(self, node_a, node_b):
        """Root the tree on edge (a, b); -> (root_partials (S,K,A), root_scale (S,K)) (reference: tree_model.py:178-198)."""
        length = self._edge_length(node_a, node_b)
        self.engine.root_lnl(node_a, node_b, length, root_pmats=self._root_pmats(length))
        return self.engine.get_root_partials()

    def _pattern_lnl(self, node_a, node_b, want_pattern=True):
        length = self._edge_length(node_a, node_b)
        if not self.store_partials:
            if not getattr(self.substitution_model, "has_real_eigensystem", True):
                raise ValueError("store_partials=False needs a model with a real eigen-system")
            return self.engine.lnl_resident(node_a, node_b, length, want_pattern=want_pattern)
        rp = self._root_pmats(length)
        if not self.ascbias:
            total, pattern, _ = self.engine.root_lnl(node_a, node_b, length, want_pattern=want_pattern, root_pmats=rp)
            return total, pattern
        # Lewis correction exactly as the reference composes it (tree_model.py:209-216): subtract
        # log(1 - sum over dummy patterns and categories of exp(lnl)) from every per-category value
        _, _, cat = self.engine.root_lnl(node_a, node_b, length, want_cat=True, root_pmats=rp)
        n = self._n_dummy
        correction = np.log(1 - np.exp(logsumexp(cat[-n:])))
        cat[:-n] -= correction
        pattern = logsumexp(cat + np.log(self.rate_model.weights), axis=1)
        total = float(np.dot(pattern[:-n], self.siteweights))
        return total, pattern

    def compute_likelihood_at_edge(self, node_a, node_b):
        """Per ORIGINAL site log-likelihoods, gathered through ``inverse_index`` (reference: tree_model.py:200-217)."""
        _, pattern = self._pattern_lnl(node_a, node_b)
        return pattern[self.inverse_index]

    def lnl(self, node_a=None, node_b=None):
        """Total log-likelihood = sum_p siteweights_p * lnl_p (what bin/phy.py:146 prints), reduced on the device."""
        if node_a is None:
            node_a, node_b = self.traversal.root_edge
        total, _ = self._pattern_lnl(node_a, node_b, want_pattern=False)   # the per-pattern vector stays on the device
        return total

    def lnl_enqueue(self, node_a=None, node_b=None, peer_sum=False):
        """Stream-ordered ``lnl``: the evaluation is only enqueued and its sum stays on the device; returns a
        one-element torch tensor VIEW of it (``engine.result_tensor``) that a collective on the same stream can
        reduce in place.  No host synchronisation.  Not available with the ascertainment-bias correction or a
        model without a real eigen-system (their last step runs on the host)."""
        if node_a is None:
            node_a, node_b = self.traversal.root_edge
        if self.ascbias or not getattr(self.substitution_model, "has_real_eigensystem", True):
            raise ValueError("lnl_enqueue: this model composes its likelihood on the host; use lnl()")
        length = self._edge_length(node_a, node_b)
        if peer_sum:                                  # arms exactly the next call (phb_peer_sum_next)
            self.engine.peer_sum_next()
        if self.store_partials:
            self.engine.root_lnl_async(node_a, node_b, length)
        else:
            self.engine.lnl_resident_async(node_a, node_b, length)
        return self.engine.result_tensor(1)

    def lnl_from_host_codes(self, packed_codes, node_a=None, node_b=None, n_chunks=0, enqueue_only=False, peer_sum=False):
        """lnL of a NEW alignment over the same taxa, tree and models, starting from pinned HOST codes: two 4-bit codes
        per byte (``LikelihoodEngine.pack_codes``) or, as a ``(low, high)`` tuple, the 3-bit planes of
        ``LikelihoodEngine.split_codes`` (look-up tables of at most 8 rows); rows in ``tip_row_order``.  The
        host-to-device copy is pipelined with the pruning (``phb_lnl_from_host_packed`` / ``_split``).  lnL-only models
        (``store_partials=False``).  ``enqueue_only``: as ``lnl_enqueue``."""
        if self.store_partials or self.ascbias:
            raise ValueError("lnl_from_host_codes needs an lnL-only model (store_partials=False) without asc-bias correction")
        if node_a is None:
            node_a, node_b = self.traversal.root_edge
        length = self._edge_length(node_a, node_b)
        split = isinstance(packed_codes, (tuple, list))
        if enqueue_only:
            if peer_sum:
                self.engine.peer_sum_next()
            if split:
                self.engine.lnl_from_host_split_async(packed_codes, node_a, node_b, length, n_chunks)
            else:
                self.engine.lnl_from_host_packed_async(packed_codes, node_a, node_b, length, n_chunks)
            return self.engine.result_tensor(1)
        if split:
            return self.engine.lnl_from_host_split(packed_codes, node_a, node_b, length, n_chunks=n_chunks)[0]
        return self.engine.lnl_from_host(packed_codes, node_a, node_b, length, n_chunks=n_chunks, packed=True)[0]

    def lnl_from_host_submit(self, packed_codes, node_a=None, node_b=None, n_chunks=0, reduce=None, peer_sum=False):
        """Pipelined ``lnl_from_host_codes``: enqueue the evaluation of one more alignment and return a ``PendingLnl``
        whose ``result()`` delivers its lnL.  Up to two are in flight: the host-to-device copy of this one runs under the
        walk of the one submitted before (many alignments over one tree - bootstrap replicates, simulated data).
        ``reduce``: called with the device tensor view of the sum before it is copied back (``ShardedTreeModel`` passes
        its all-reduce)."""
        if self.store_partials or self.ascbias:
            raise ValueError("lnl_from_host_submit needs an lnL-only model (store_partials=False) without asc-bias correction")
        if node_a is None:
            node_a, node_b = self.traversal.root_edge
        length = self._edge_length(node_a, node_b)
        pending = getattr(self, "_pending_lnl", None) or {}
        for other in pending.values():
            other.post()                              # their copies to the host go in front of the new walk
        if peer_sum:
            self.engine.peer_sum_next()
        slot = self.engine.host_fed_submit(packed_codes, node_a, node_b, length, n_chunks)
        if slot in pending:
            pending[slot].result()                    # the slot's previous tenant: its value must be read before the word is reused
        handle = PendingLnl(self, slot, reduce)
        pending[slot] = handle
        self._pending_lnl = pending
        return handle

    # ------------------------------------------------------------------------------------------
    # attribute views of device state (reference attributes partials / scale / root_partials / root_scale)
    # ------------------------------------------------------------------------------------------
    @property
    def partials(self):
        n_nodes = 2 * self.engine.n_tips - 2
        return np.stack([self.engine.get_partials(i) for i in range(n_nodes)])

    @property
    def scale(self):
        n_nodes = 2 * self.engine.n_tips - 2
        return np.stack([self.engine.get_scalers(i) for i in range(n_nodes)])

    @property
    def root_partials(self):
        return self.engine.get_root_partials()[0]

    @property
    def root_scale(self):
        return self.engine.get_root_partials()[1]

    # ------------------------------------------------------------------------------------------
    # derivatives (composition of lnl_branch_derivs over the Gamma mixture)
    # ------------------------------------------------------------------------------------------
    def compute_up_partials(self):
        """Pre-order pass: partial of everything outside each node's subtree (needs ``up_partials=True``)."""
        a, b = self.traversal.root_edge
        self.engine.compute_up_partials(a, b, self.traversal.brlens[(a, b)])

    def edge_derivatives(self, nodes, lengths=None, chain_rule=True):
        """
        For each node id (meaning the edge above it), (lnL, dlnL/dt, d2lnL/dt2) at ``lengths`` (default:
        the current branch lengths).  ``chain_rule=False`` reproduces Model.dp_dt's convention.
        """
        nodes = np.asarray(nodes, dtype=np.int32)
        if lengths is None:
            lengths = self.lengths_above(nodes)
        return self.engine.edge_derivatives(nodes, lengths, chain_rule)

    def edge_derivatives_enqueue(self, nodes, lengths=None, chain_rule=True):
        """Stream-ordered ``edge_derivatives``: returns a torch tensor VIEW (n_edges, 3) of the device sums."""
        nodes = np.asarray(nodes, dtype=np.int32)
        if lengths is None:
            lengths = self.lengths_above(nodes)
        n = self.engine.edge_derivatives_async(nodes, lengths, chain_rule)
        return self.engine.result_tensor(3 * n).view(n, 3)

    def edge_keys(self, nodes):
        """``brlens`` key of the edge above each node (the root edge for either root child)."""
        a, b = self.traversal.root_edge
        br = self.traversal.brlens
        parents = self._parents()
        keys = []
        for n in nodes:
            n = int(n)
            if n == a or n == b:
                keys.append(br.canonical_key((a, b)))
            elif n in parents:
                keys.append(br.canonical_key((n, parents[n])))
            else:
                raise ValueError("node {} has no edge above it".format(n))
        return keys

    def lengths_above(self, nodes):
        br = self.traversal.brlens
        return br.gather(br.slots(self.edge_keys(nodes)))

    def _parents(self):
        parents = getattr(self, "_parent_map", None)
        if parents is None or getattr(self, "_parent_map_for", None) is not self.traversal:
            parents = {}
            for par, c1, c2 in self.traversal.postorder_traversal:
                parents[int(c1)] = int(par)
                parents[int(c2)] = int(par)
            self._parent_map, self._parent_map_for = parents, self.traversal
        return parents

    def branch_length_above(self, node):
        """Length of the edge above ``node`` (the root edge for either root child)."""
        a, b = self.traversal.root_edge
        if node == a or node == b:
            return self.traversal.brlens[(a, b)]
        parents = self._parents()
        if int(node) not in parents:
            raise ValueError("node {} has no edge above it".format(node))
        return self.traversal.brlens[(int(node), parents[int(node)])]
