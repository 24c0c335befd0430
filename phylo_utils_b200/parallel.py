"""
Site-pattern sharding across the GPUs of one box (BASELINE north_star subsystem 5).

Patterns are independent through the whole pruning recursion (the reference's `clv` simply broadcasts
over sites), so each rank owns a contiguous block of patterns - tip codes, partials, scalers and weights
are sliced on the pattern axis; tree schedule, eigen-system and P matrices are replicated (KBs).  No data
moves between GPUs inside an evaluation: the only exchange is an all-reduce of the scalar lnL (or of the
3 x n_edges derivative sums), issued through torch.distributed (NCCL over NVLink on GPUs, gloo in the CPU
tests).  Per-site output, when asked for, is an all-gather of the per-pattern vector.

The sums never visit the host on their way into the collective: the engine's stream-ordered entry points
(phb_lnl_resident_async, phb_root_lnl_async, phb_edge_derivatives_async, phb_lnl_from_host_packed_async) leave
them in the context's device result buffer, the collective reduces a tensor VIEW of that buffer in place, ordered
behind the kernels on the same stream, and one 8-byte (or 3 x n_edges x 8-byte) copy brings the global value
back: one host synchronisation per evaluation, the same as on one GPU.

One process per GPU; launch with ``python -m torch.distributed.run --nproc-per-node N ...``.
"""
import numpy as np

from .tree_model import TreeModel

__all__ = ["shard_bounds", "shard_slices", "allreduce_sum", "allgather_concat", "ShardedTreeModel"]


def shard_bounds(n_patterns, rank, world):
    """Contiguous, balanced block [lo, hi) of the pattern axis owned by ``rank`` (sizes differ by at most one)."""
    if not 0 <= rank < world:
        raise ValueError("rank {} outside world of size {}".format(rank, world))
    base, extra = divmod(int(n_patterns), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_slices(n_patterns, world):
    return [shard_bounds(n_patterns, r, world) for r in range(world)]


def _dist():
    import torch.distributed as dist
    return dist if dist.is_available() and dist.is_initialized() else None


def allreduce_sum(values, device=None):
    """Sum a small float64 array over all ranks (identity when torch.distributed is not initialised)."""
    arr = np.atleast_1d(np.asarray(values, dtype=np.double))
    dist = _dist()
    if dist is None or dist.get_world_size() == 1:
        return arr.copy()
    import torch
    t = torch.from_numpy(arr.copy())
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def allgather_concat(local, sizes, device=None):
    """Concatenate per-rank 1-D float64 arrays of known ``sizes`` in rank order."""
    local = np.ascontiguousarray(local, dtype=np.double)
    dist = _dist()
    if dist is None or dist.get_world_size() == 1:
        return local.copy()
    import torch
    width = max(sizes)
    pad = np.zeros(width)
    pad[:local.shape[0]] = local
    mine = torch.from_numpy(pad)
    if device is not None:
        mine = mine.to(device)
    parts = [torch.empty_like(mine) for _ in sizes]
    dist.all_gather(parts, mine)
    return np.concatenate([p.cpu().numpy()[:n] for p, n in zip(parts, sizes)])


class ShardedTreeModel(object):
    """
    TreeModel whose pattern axis is split over the ranks of the current torch.distributed group.
    Every rank calls every method (SPMD); scalar results are identical on all ranks.
    """

    def __init__(self, device=None, up_partials=False, mode="auto", store_partials=True, local_model=None):
        """``local_model``: the per-rank model object (default: a ``TreeModel`` on this rank's GPU).  Anything with
        TreeModel's methods will do - the CPU tests of the sharding logic plug in a stand-in."""
        dist = _dist()
        self.rank = dist.get_rank() if dist else 0
        self.world = dist.get_world_size() if dist else 1
        if device is None:
            import os
            device = int(os.environ.get("LOCAL_RANK", "0"))
        self.device = device
        self.local = local_model if local_model is not None else TreeModel(
            device=device, up_partials=up_partials, mode=mode, store_partials=store_partials)
        self._torch_device = None
        self.n_patterns = None
        self.collectives = 0          # all-reduces / all-gathers issued so far (bench and tests read it)
        self.peer_sums = False        # scalar lnL sums run inside the reduction kernel over peer memory (initialise())
        self.peer_exchanges = 0       # ... how many of the `collectives` were such exchanges

    def _comm_device(self):
        dist = _dist()
        if dist is not None and dist.get_backend() == "nccl":
            import torch
            return torch.device("cuda", self.device)
        return None

    # -- inputs: same calls as TreeModel; the alignment is given in full and sliced here ------------------
    def set_tree(self, tree):
        self.local.set_tree(tree)

    def set_substitution_model(self, model):
        self.local.set_substitution_model(model)

    def set_rate_model(self, rate_model):
        self.local.set_rate_model(rate_model)

    def set_tip_codes(self, codes, lut, names, siteweights=None, inverse_index=None):
        """``codes`` (ntax, npat) is the FULL compressed alignment; this rank keeps patterns [lo, hi)."""
        npat = codes.shape[1]
        if npat < self.world:        # the same test on every rank: all of them raise, none is left waiting in a collective
            raise ValueError("more ranks ({}) than site patterns ({})".format(self.world, npat))
        self.n_patterns = npat
        self.sizes = [hi - lo for lo, hi in shard_slices(npat, self.world)]
        lo, hi = shard_bounds(npat, self.rank, self.world)
        self.lo, self.hi = lo, hi
        self.inverse_index = np.arange(npat) if inverse_index is None else np.asarray(inverse_index)
        w = None if siteweights is None else np.asarray(siteweights)[lo:hi]
        self.local.set_tip_codes(np.ascontiguousarray(codes[:, lo:hi]), lut, names, w)

    def set_local_tip_codes(self, codes, lut, names, n_patterns, siteweights=None):
        """This rank's shard only: ``codes`` (ntax, hi - lo) are patterns [lo, hi) = ``shard_bounds(n_patterns, rank,
        world)`` of an alignment of ``n_patterns`` patterns that no rank needs to hold in full (synthetic or
        pre-sharded data).  Per-site output (``compute_likelihood_at_edge``) is then in pattern order."""
        if int(n_patterns) < self.world:
            raise ValueError("more ranks ({}) than site patterns ({})".format(self.world, n_patterns))
        self.n_patterns = int(n_patterns)
        self.sizes = [hi - lo for lo, hi in shard_slices(self.n_patterns, self.world)]
        self.lo, self.hi = shard_bounds(self.n_patterns, self.rank, self.world)
        if codes.shape[1] != self.hi - self.lo:
            raise ValueError("rank {} owns {} patterns, got {}".format(self.rank, self.hi - self.lo, codes.shape[1]))
        self.inverse_index = np.arange(self.n_patterns)
        self.local.set_tip_codes(codes, lut, names, siteweights)

    def set_alignment(self, alignment, alphabet, compress=True):
        from .alignment.alignment import alignment_to_codes
        codes, lut, sw, ii, names = alignment_to_codes(alignment, alphabet, compress)
        self.set_tip_codes(codes, lut, names, sw, ii)

    def set_ascertainment_bias_correction(self):
        """Lewis correction under sharding (SURVEY.md 8(e) "exchange steps"): EVERY rank appends the reference's
        one-constant-pattern-per-state dummy block (tree_model.py:151-156) to its own shard with weight 0, so each
        rank derives the identical correction (tree_model.py:209-214) from its own evaluation - a pattern's value
        does not depend on where in a shard it sits - and nothing has to be broadcast.  The dummy patterns never
        reach the totals (weight 0) nor the gathered per-site vector (stripped before the all-gather)."""
        self.local.set_ascertainment_bias_correction()

    def initialise(self):
        self.local.initialise()
        self._connect_peers()

    def _connect_peers(self):
        """Map every rank's exchange buffer into every other rank (CUDA IPC; the ranks of ONE box), so that the scalar lnL
        sum runs inside the kernel that reduces the per-CTA sums (``phb_peer_sum_next``): no collective-library call on
        the evaluation path.  All ranks agree on the outcome; if any of them cannot map a peer (ranks on different
        nodes, IPC unavailable, ``PHB_PEER_SUM=0``) all of them stay with the in-place ``all_reduce``."""
        import os
        self.peer_sums = False
        dist = _dist()
        engine = getattr(self.local, "engine", None)
        if dist is None or self.world < 2 or self.world > 16 or engine is None or not hasattr(engine, "peer_buffer"):
            return
        import torch
        dev = self._comm_device() or torch.device("cpu")
        n = engine.PEER_HANDLE_BYTES
        ok, handle = 1, bytes(n)
        if os.environ.get("PHB_PEER_SUM", "1") == "0":
            ok = 0
        else:
            try:
                handle = engine.peer_buffer()
            except RuntimeError:
                ok = 0
        mine = torch.tensor(list(handle) + [ok], dtype=torch.uint8, device=dev)
        got = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(got, mine)
        got = [bytes(t.cpu().tolist()) for t in got]
        if all(g[n] == 1 for g in got):
            try:
                engine.peer_connect(self.rank, self.world, b"".join(g[:n] for g in got))
            except RuntimeError:
                ok = 0
        else:
            ok = 0
        agreed = torch.tensor([ok], dtype=torch.int32, device=dev)
        dist.all_reduce(agreed, op=dist.ReduceOp.MIN)
        self.peer_sums = bool(int(agreed.item()))

    def _peer_value(self, value):
        if value != value:
            raise RuntimeError("peer sum: a rank never delivered its value (ranks out of step, or one of them died)")
        return float(value)

    def compute_partials(self):
        self.local.compute_partials()

    @property
    def traversal(self):
        return self.local.traversal

    # -- results ----------------------------------------------------------------------------------------------
    def _device_sums(self):
        """True when the local sums can stay on the device on their way through the collective."""
        m = self.local
        return _dist() is not None and self.world > 1 and not m.ascbias and hasattr(m, "lnl_enqueue") and \
            getattr(m.substitution_model, "has_real_eigensystem", True)

    def _reduce_in_place(self, view):
        """all-reduce (sum) of a device tensor view of the engine's result buffer, then one copy to the host."""
        _dist().all_reduce(view)
        self.collectives += 1
        return self.local.engine.result_fetch(view.numel())

    def lnl(self, node_a=None, node_b=None):
        if self._device_sums() and self.peer_sums:
            self.local.lnl_enqueue(node_a, node_b, peer_sum=True)      # the global sum is formed by the reduction kernel
            self.collectives += 1
            self.peer_exchanges += 1
            return self._peer_value(self.local.engine.result_fetch(1)[0])
        if self._device_sums():
            return float(self._reduce_in_place(self.local.lnl_enqueue(node_a, node_b))[0])
        if self.world > 1:
            self.collectives += 1
        return float(allreduce_sum([self.local.lnl(node_a, node_b)], self._comm_device())[0])

    def lnl_from_host_codes(self, packed_codes, node_a=None, node_b=None, n_chunks=0):
        """``TreeModel.lnl_from_host_codes`` on this rank's shard of a new alignment (pinned host memory, two codes
        per byte), summed over the ranks on the device."""
        if self._device_sums() and self.peer_sums:
            self.local.lnl_from_host_codes(packed_codes, node_a, node_b, n_chunks, enqueue_only=True, peer_sum=True)
            self.collectives += 1
            self.peer_exchanges += 1
            return self._peer_value(self.local.engine.result_fetch(1)[0])
        if self._device_sums():
            view = self.local.lnl_from_host_codes(packed_codes, node_a, node_b, n_chunks, enqueue_only=True)
            return float(self._reduce_in_place(view)[0])
        return float(allreduce_sum([self.local.lnl_from_host_codes(packed_codes, node_a, node_b, n_chunks)],
                                   self._comm_device())[0])

    def lnl_from_host_submit(self, packed_codes, node_a=None, node_b=None, n_chunks=0):
        """Pipelined form of ``lnl_from_host_codes`` (``TreeModel.lnl_from_host_submit``): returns a handle whose
        ``result()`` is the global lnL; the all-reduce of an evaluation is enqueued behind its walk when the next one is
        submitted or its result is asked for."""
        reduce = None
        if self._device_sums() and self.peer_sums:
            self.collectives += 1
            self.peer_exchanges += 1
            handle = self.local.lnl_from_host_submit(packed_codes, node_a, node_b, n_chunks, None, peer_sum=True)
            handle.check = self._peer_value
            return handle
        if self._device_sums():
            def reduce(view):
                _dist().all_reduce(view)
                self.collectives += 1
        elif self.world > 1:
            raise ValueError("pipelined evaluations need the device-side reduction (a CUDA process group)")
        return self.local.lnl_from_host_submit(packed_codes, node_a, node_b, n_chunks, reduce)

    def compute_likelihood_at_edge(self, node_a, node_b):
        _, pattern = self.local._pattern_lnl(node_a, node_b)
        pattern = pattern[:self.hi - self.lo]              # without the ascertainment-bias dummy patterns
        if self.world > 1:
            self.collectives += 1
        full = allgather_concat(pattern, self.sizes, self._comm_device())
        return full[self.inverse_index]

    def compute_up_partials(self):
        self.local.compute_up_partials()

    def edge_derivatives(self, nodes, lengths=None, chain_rule=True):
        if self._device_sums():
            nodes = np.asarray(nodes, dtype=np.int32)
            if lengths is None:
                lengths = self.local.lengths_above(nodes)
            lengths = np.asarray(lengths, dtype=np.double)
            # the device result buffer holds 3 sums for every edge of the tree; longer lists (the same edge at several
            # trial lengths) go out in pieces
            cap = max(1, self.local.engine.result_capacity // 3)
            out = np.empty((nodes.shape[0], 3))
            for lo in range(0, nodes.shape[0], cap):
                view = self.local.edge_derivatives_enqueue(nodes[lo:lo + cap], lengths[lo:lo + cap], chain_rule)
                out[lo:lo + cap] = self._reduce_in_place(view.view(-1)).reshape(-1, 3)
            return out
        part = self.local.edge_derivatives(nodes, lengths, chain_rule)
        if self.world > 1:
            self.collectives += 1
        return allreduce_sum(part.ravel(), self._comm_device()).reshape(part.shape)
