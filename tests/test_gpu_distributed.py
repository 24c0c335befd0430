"""Two-rank NCCL run of the sharded TreeModel (needs >= 2 GPUs; skipped otherwise)."""
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


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from helpers import problem, records, tree
    from phylo_utils_b200.parallel import ShardedTreeModel
    g, tr, codes, lut, sw, ii, names, model, rate = problem("cfg1_gtr_g4")
    tm = ShardedTreeModel(up_partials=True)
    tm.set_tree(tree(g))
    tm.set_alignment(records(g), 0)
    tm.set_rate_model(rate)
    tm.set_substitution_model(model)
    tm.initialise()
    total = tm.lnl()
    site = tm.compute_likelihood_at_edge(*tm.traversal.root_edge)
    tm.compute_up_partials()
    d = tm.edge_derivatives(np.arange(6))
    np.savez(os.path.join(out_dir, "rank{}.npz".format(rank)), total=total, site=site, d=d)
    dist.destroy_process_group()


@pytest.mark.skipif(_n_gpus() < 2, reason="needs two GPUs")
def test_two_rank_nccl_sharded_tree_model(tmp_path):
    import torch.multiprocessing as mp
    from helpers import load
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    g = load("cfg1_gtr_g4")
    outs = [np.load(os.path.join(str(tmp_path), "rank{}.npz".format(r))) for r in range(2)]
    for z in outs:
        assert abs(float(z["total"]) - float(g["total_lnl"])) <= 1e-10 * abs(float(g["total_lnl"]))
        assert np.allclose(z["site"], g["site_lnl"], rtol=1e-10, atol=0)
        assert np.all(np.abs(z["d"][:, 0] - float(g["total_lnl"])) <= 1e-10 * abs(float(g["total_lnl"])))
    assert np.array_equal(outs[0]["d"], outs[1]["d"])
