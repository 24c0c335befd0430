"""
Two-rank runs of the sharded TreeModel on the GPU box.

With two or more GPUs the group is NCCL, one rank per GPU (the product configuration).  On a one-GPU box - the
driver's test box - both ranks share cuda:0 and the group is gloo, which accepts CUDA tensors: ShardedTreeModel's
device-side reduction path (stream-ordered evaluation -> in-place all-reduce of a tensor view of the engine's result
buffer -> one fetch) runs unchanged, only the transport differs.  The scalar lnL sums do not go through the group at
all where the ranks can map each other's memory: they are formed inside the reduction kernel (phb_peer_*).  Either way the checks are the reference's golden
outputs (tests/golden/, written by the unmodified reference).
"""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _init(rank, world, port, n_gpus):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    import torch
    import torch.distributed as dist
    device = rank if n_gpus >= world else 0
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), LOCAL_RANK=str(device))
    torch.cuda.set_device(device)
    if n_gpus >= world:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", device))
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    return dist, device


def _worker(rank, world, port, out_dir, n_gpus):
    dist, device = _init(rank, world, port, n_gpus)
    from helpers import problem, records, tree
    import phylo_utils_b200 as phy
    from phylo_utils_b200.parallel import ShardedTreeModel, shard_bounds
    g, tr, codes, lut, sw, ii, names, model, rate = problem("cfg1_gtr_g4")
    tm = ShardedTreeModel(device=device, up_partials=True)
    tm.set_tree(tree(g))
    tm.set_alignment(records(g), 0)
    tm.set_rate_model(rate)
    tm.set_substitution_model(model)
    tm.initialise()
    assert tm._device_sums()                      # the sums stay on the device on their way through the collective
    total = tm.lnl()
    site = tm.compute_likelihood_at_edge(*tm.traversal.root_edge)
    tm.compute_up_partials()
    nodes = phy.optimise.edge_nodes(tm.traversal)
    d = tm.edge_derivatives(nodes)
    d_host = tm.local.edge_derivatives(nodes)     # this rank's share only
    n_coll = tm.collectives
    # more (edge, trial length) pairs than the device result buffer holds: goes out in pieces
    reps = tm.local.engine.result_capacity // (3 * len(nodes)) + 2
    d_long = tm.edge_derivatives(np.tile(nodes, reps))
    assert np.array_equal(d_long, np.tile(d, (reps, 1))) and tm.collectives > n_coll + 1

    # lnL-only model: resident walk and the host-fed (pipelined, packed) evaluation of this rank's shard
    lo, hi = shard_bounds(codes.shape[1], rank, world)
    used = np.unique(codes)                       # the look-up table restricted to the rows this alignment uses (4 or 5 of 15)
    codes, lut = np.searchsorted(used, codes).astype(np.uint8), lut[used]
    tl = ShardedTreeModel(device=device, store_partials=False)
    tl.set_tree(tree(g))
    tl.set_tip_codes(codes, lut, names, sw, ii)
    tl.set_rate_model(rate)
    tl.set_substitution_model(model)
    tl.initialise()
    total_lnl_only = tl.lnl()
    order = tl.local.tip_row_order
    packed = phy.LikelihoodEngine.pack_codes(np.ascontiguousarray(codes[order][:, lo:hi]))
    total_from_host = tl.lnl_from_host_codes(packed, n_chunks=4)
    planes = phy.LikelihoodEngine.split_codes(np.ascontiguousarray(codes[order][:, lo:hi]))
    assert lut.shape[0] <= 8
    assert tl.lnl_from_host_codes(planes, n_chunks=3) == total_from_host      # 3 bits per code in two planes: same bits out
    h0 = tl.lnl_from_host_submit(planes)                                       # two in flight, all-reduced on the device
    h1 = tl.lnl_from_host_submit(packed)
    assert h0.result() == total_from_host and h1.result() == total_from_host
    # Where the ranks can map each other's exchange buffers (CUDA IPC: the ranks of one box) the scalar sums above were
    # formed INSIDE the reduction kernel (phb_peer_sum_next), not by the process group.  The same model over the
    # group's all_reduce must give the same bits (two addends: the order cannot matter).
    assert tl.peer_exchanges == (5 if tl.peer_sums else 0) and tm.peer_sums == tl.peer_sums
    print("peer_sums rank {}: {} ({} exchanges)".format(rank, tl.peer_sums, tl.peer_exchanges), flush=True)
    os.environ["PHB_PEER_SUM"] = "0"
    tn = ShardedTreeModel(device=device, store_partials=False)
    tn.set_tree(tree(g))
    tn.set_tip_codes(codes, lut, names, sw, ii)
    tn.set_rate_model(rate)
    tn.set_substitution_model(model)
    tn.initialise()
    del os.environ["PHB_PEER_SUM"]
    assert not tn.peer_sums and tn.lnl() == total_lnl_only and tn.peer_exchanges == 0
    assert tn.lnl_from_host_submit(packed).result() == total_from_host

    # Lewis ascertainment-bias correction under sharding: dummy patterns on every rank, no broadcast
    ga, _, codes_a, lut_a, sw_a, ii_a, names_a, model_a, rate_a = problem("ascbias_gtr_g4")
    ta = ShardedTreeModel(device=device)
    ta.set_tree(tree(ga))
    ta.set_tip_codes(codes_a, lut_a, names_a, sw_a, ii_a)
    ta.set_rate_model(rate_a)
    ta.set_substitution_model(model_a)
    ta.set_ascertainment_bias_correction()
    ta.initialise()
    asc_total = ta.lnl()
    asc_site = ta.compute_likelihood_at_edge(*ta.traversal.root_edge)

    np.savez(os.path.join(out_dir, "rank{}.npz".format(rank)), total=total, site=site, d=d, d_host=d_host, n_coll=n_coll,
             peer=int(tl.peer_sums),
             total_lnl_only=total_lnl_only, total_from_host=total_from_host, asc_total=asc_total, asc_site=asc_site)
    dist.destroy_process_group()


def test_two_rank_sharded_tree_model(tmp_path):
    import torch.multiprocessing as mp
    from helpers import load
    n_gpus = _n_gpus()
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path), n_gpus), nprocs=2, join=True)
    g, ga = load("cfg1_gtr_g4"), load("ascbias_gtr_g4")
    want = float(g["total_lnl"])
    outs = [np.load(os.path.join(str(tmp_path), "rank{}.npz".format(r))) for r in range(2)]
    for z in outs:
        assert abs(float(z["total"]) - want) <= 1e-10 * abs(want)
        assert abs(float(z["total_lnl_only"]) - want) <= 1e-10 * abs(want)
        assert abs(float(z["total_from_host"]) - want) <= 1e-10 * abs(want)
        assert np.allclose(z["site"], g["site_lnl"], rtol=1e-10, atol=0)
        assert np.all(np.abs(z["d"][:, 0] - want) <= 1e-10 * abs(want))
        assert int(z["n_coll"]) == 3                                   # lnl, per-site gather, derivatives
        assert abs(float(z["asc_total"]) - float(ga["total_lnl"])) <= 1e-10 * abs(float(ga["total_lnl"]))
        assert np.allclose(z["asc_site"], ga["site_lnl"], rtol=1e-10, atol=0)
    assert int(outs[0]["peer"]) == int(outs[1]["peer"])                # the ranks agreed on how the scalar sums are formed
    assert np.array_equal(outs[0]["d"], outs[1]["d"])                  # every rank holds the same global sums
    assert np.allclose(outs[0]["d_host"] + outs[1]["d_host"], outs[0]["d"], rtol=1e-12, atol=1e-9)
    assert not np.allclose(outs[0]["d_host"][:, 0], outs[0]["d"][:, 0])   # a shard alone is not the total


def _raise_worker(rank, world, port, out_dir, n_gpus):
    dist, device = _init(rank, world, port, n_gpus)
    from phylo_utils_b200.parallel import ShardedTreeModel
    tm = ShardedTreeModel(device=device)
    try:
        tm.set_tip_codes(np.zeros((4, 1), dtype=np.uint8), np.eye(4), {"a": 0, "b": 1, "c": 2, "d": 3})
        raised = False
    except ValueError:
        raised = True
    open(os.path.join(out_dir, "raised{}".format(rank)), "w").write(str(raised))
    dist.destroy_process_group()


def test_more_ranks_than_patterns_raises_on_every_rank(tmp_path):
    import torch.multiprocessing as mp
    mp.spawn(_raise_worker, args=(2, _free_port(), str(tmp_path), _n_gpus()), nprocs=2, join=True)
    assert [open(os.path.join(str(tmp_path), "raised{}".format(r))).read() for r in range(2)] == ["True", "True"]
