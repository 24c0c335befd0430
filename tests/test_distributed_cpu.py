"""
Multi-rank host logic on CPU: world_size-2 gloo group, pattern sharding + the scalar all-reduce /
per-pattern all-gather of phylo_utils_b200.parallel.  The per-shard evaluator is the CPU oracle here
(no GPU in this container); on the GPU box the same functions run over NCCL (tests/test_gpu_distributed.py).
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from phylo_utils_b200 import parallel

HERE = os.path.dirname(os.path.abspath(__file__))


def test_shard_bounds_cover_the_axis_exactly():
    for n in (1, 2, 7, 8, 1000, 1000003):
        for world in (1, 2, 3, 8):
            if world > n:
                continue
            blocks = parallel.shard_slices(n, world)
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        parallel.shard_bounds(10, 2, 2)


def test_collectives_are_identity_without_a_process_group():
    assert parallel.allreduce_sum([1.5, 2.5]).tolist() == [1.5, 2.5]
    assert parallel.allgather_concat(np.arange(3.0), [3]).tolist() == [0.0, 1.0, 2.0]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from helpers import problem, tip_partials
    from oracle import oracle
    g, tr, codes, lut, sw, ii, names, model, rate = problem("cfg1_gtr_g4")
    npat = codes.shape[1]
    lo, hi = parallel.shard_bounds(npat, rank, world)
    local_codes = np.ascontiguousarray(codes[:, lo:hi])
    tips = tip_partials(tr, local_codes, lut, names)
    pattern = oracle.tree_lnl(tr, tips, model.p, model.freqs, rate.rates, rate.weights, n_threads=1)
    total = parallel.allreduce_sum([np.dot(pattern, sw[lo:hi])])[0]
    sizes = [b - a for a, b in parallel.shard_slices(npat, world)]
    full = parallel.allgather_concat(pattern, sizes)
    sums = parallel.allreduce_sum(np.array([[1.0, 2.0, 3.0]]) * (rank + 1))
    np.savez(os.path.join(out_dir, "rank{}.npz".format(rank)), total=total, site=full[ii], sums=sums)
    dist.destroy_process_group()


class OracleLocalModel(object):
    """The slice of TreeModel that ShardedTreeModel drives, evaluated by the CPU oracle (no GPU here): what lets the
    sharding logic itself - slicing, dummy patterns of the ascertainment-bias correction, which sums go through which
    collective - run under a two-rank gloo group."""
    ascbias = False

    def __init__(self):
        self.substitution_model = self.rate_model = self.traversal = None

    def set_tree(self, tree):
        import phylo_utils_b200 as phy
        self.traversal = phy.traversal.Traversal(phy.utils.deepcopy_tree(tree))

    def set_substitution_model(self, model):
        self.substitution_model = model

    def set_rate_model(self, rate):
        self.rate_model = rate

    def set_tip_codes(self, codes, lut, names, siteweights=None):
        self.codes, self.lut, self.names = codes, lut, names
        self.siteweights = np.ones(codes.shape[1]) if siteweights is None else np.asarray(siteweights, dtype=float)

    def set_ascertainment_bias_correction(self):
        self.ascbias = True

    def initialise(self):
        pass

    def compute_partials(self):
        pass

    def _pattern_lnl(self, node_a, node_b, want_pattern=True):
        from scipy.special import logsumexp
        from helpers import tip_partials
        from oracle import oracle
        m, r, tr = self.substitution_model, self.rate_model, self.traversal
        codes, lut = self.codes, self.lut
        n_states = lut.shape[1]
        if self.ascbias:            # one constant dummy pattern per state behind the real ones (tree_model.py:151-156)
            single = [int(np.flatnonzero((lut == np.eye(n_states)[s]).all(axis=1))[0]) for s in range(n_states)]
            codes = np.hstack([codes, np.repeat(np.asarray(single, dtype=np.uint8)[None, :], codes.shape[0], axis=0)])
        pattern, ot = oracle.tree_lnl(tr, tip_partials(tr, codes, lut, self.names), m.p, m.freqs, r.rates, r.weights,
                                      n_threads=1, return_tree=True)
        if self.ascbias:            # tree_model.py:209-216
            a, b = tr.root_edge
            length = tr.brlens[(a, b)]
            _, cat = ot.likelihood_at_edge(a, b, np.stack([m.p(0, r.rates), m.p(length, r.rates)]), m.freqs, r.weights, want_cat=True)
            cat[:-n_states] -= np.log(1 - np.exp(logsumexp(cat[-n_states:])))
            pattern = logsumexp(cat + np.log(r.weights), axis=1)
            return float(np.dot(pattern[:-n_states], self.siteweights)), pattern
        return float(np.dot(pattern, self.siteweights)), pattern

    def lnl(self, node_a=None, node_b=None):
        return self._pattern_lnl(node_a, node_b)[0]


def _sharded_model_worker(rank, world, port, out_dir):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from helpers import problem, tree
    out = {}
    for name, asc in (("cfg1_gtr_g4", False), ("ascbias_gtr_g4", True)):
        g, tr, codes, lut, sw, ii, names, model, rate = problem(name)
        tm = parallel.ShardedTreeModel(device=0, local_model=OracleLocalModel())
        tm.set_tree(tree(g))
        tm.set_tip_codes(codes, lut, names, sw, ii)
        tm.set_rate_model(rate)
        tm.set_substitution_model(model)
        if asc:
            tm.set_ascertainment_bias_correction()
        tm.initialise()
        assert not tm._device_sums() and tm.hi - tm.lo == tm.sizes[rank] and sum(tm.sizes) == codes.shape[1]
        out[name + "_total"] = tm.lnl()
        out[name + "_site"] = tm.compute_likelihood_at_edge(*tm.traversal.root_edge)
        out[name + "_collectives"] = tm.collectives
    raised = False
    try:
        parallel.ShardedTreeModel(device=0, local_model=OracleLocalModel()).set_tip_codes(np.zeros((4, 1), dtype=np.uint8), np.eye(4), {})
    except ValueError:
        raised = True                      # more ranks than patterns: every rank raises, none is left in a collective
    out["raised"] = raised
    np.savez(os.path.join(out_dir, "sharded{}.npz".format(rank)), **out)
    dist.destroy_process_group()


def test_two_rank_gloo_sharded_tree_model_logic(tmp_path):
    """ShardedTreeModel itself under a two-rank gloo group, its per-rank model replaced by an oracle-backed stand-in."""
    from helpers import load
    world = 2
    mp.spawn(_sharded_model_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    for rank in range(world):
        with np.load(os.path.join(str(tmp_path), "sharded{}.npz".format(rank))) as z:
            for name in ("cfg1_gtr_g4", "ascbias_gtr_g4"):
                g = load(name)
                assert abs(float(z[name + "_total"]) - float(g["total_lnl"])) <= 1e-10 * abs(float(g["total_lnl"]))
                assert np.allclose(z[name + "_site"], g["site_lnl"], rtol=1e-10, atol=0)
                assert int(z[name + "_collectives"]) == 2            # one all-reduce, one all-gather
            assert bool(z["raised"])


def test_two_rank_gloo_sharded_lnl_matches_reference(tmp_path):
    from helpers import load
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    g = load("cfg1_gtr_g4")
    for rank in range(world):
        with np.load(os.path.join(str(tmp_path), "rank{}.npz".format(rank))) as z:
            assert abs(float(z["total"]) - float(g["total_lnl"])) <= 1e-10 * abs(float(g["total_lnl"]))
            assert np.allclose(z["site"], g["site_lnl"], rtol=1e-10, atol=0)
            assert z["sums"].ravel().tolist() == [3.0, 6.0, 9.0]
