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
