"""
BASELINE config 5 as it is named: branch-length Newton sweeps, GTR+G4, 2000 taxa x 500k patterns sharded over the
GPUs of one box (pattern shards, scalar / 3-per-edge sums combined by an NCCL all-reduce).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 \
        tools/bench_cfg5_sharded.py [--taxa 2000] [--patterns 500000] [--sweeps 3] [--inner 2]

Each rank synthesises only its own shard (same tree everywhere, rank-seeded tip codes).  Phases are timed with CUDA
events between barriers; the line printed by rank 0 carries the MAX over ranks.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import phylo_utils_b200 as phy  # noqa: E402
from phylo_utils_b200.parallel import ShardedTreeModel, shard_bounds  # noqa: E402
from phylo_utils_b200.tree import random_tree  # noqa: E402
from phylo_utils_b200.optimise import optimise_branch_lengths, edge_nodes  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--taxa", type=int, default=2000)
    ap.add_argument("--patterns", type=int, default=500000)
    ap.add_argument("--sweeps", type=int, default=3)
    ap.add_argument("--inner", type=int, default=2)
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    tree = random_tree(args.taxa, 5)
    names = [l.taxon.label for l in tree.leaf_node_iter()]
    lut = np.vstack([np.eye(4)[::-1], np.ones((1, 4))])
    lo, hi = shard_bounds(args.patterns, rank, world)
    rng = np.random.default_rng(5000 + rank)
    codes = rng.integers(0, 4, size=(args.taxa, hi - lo)).astype(np.uint8)
    codes[rng.random(codes.shape) < 0.01] = 4
    tm = ShardedTreeModel(device=local_rank, up_partials=True)
    tm.set_tree(tree)
    tm.set_local_tip_codes(codes, lut, {n: i for i, n in enumerate(names)}, args.patterns)
    tm.set_rate_model(phy.rate_models.GammaRateModel(4, 0.5))
    tm.set_substitution_model(phy.substitution_models.GTR([6., 5., 4., 3., 2., 1.], [0.1, 0.2, 0.3, 0.4]))
    tm.initialise()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, reps):
        fn()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            out = fn()
        b.record()
        barrier()
        ms = torch.tensor([a.elapsed_time(b) / reps], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms[0]), out

    nodes = edge_nodes(tm.traversal)
    lengths = tm.local.lengths_above(nodes)
    lnl_ms, lnl = timed(lambda: (tm.compute_partials(), tm.lnl())[1], args.reps)
    up_ms, _ = timed(lambda: tm.compute_up_partials(), args.reps)
    d_ms, d = timed(lambda: tm.edge_derivatives(nodes, lengths), args.reps)
    barrier()
    t0 = time.perf_counter()
    res = optimise_branch_lengths(tm, max_sweeps=args.sweeps, inner_iterations=args.inner, tol=0.0)
    barrier()
    wall = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(wall, op=dist.ReduceOp.MAX)
    if rank == 0:
        n_int = args.taxa - 2
        print(json.dumps({
            "config": "cfg5 GTR+G4 {} taxa x {} patterns over {} GPUs".format(args.taxa, args.patterns, world),
            "n_gpus": world, "patterns_per_gpu": hi - lo, "lnl": lnl,
            "lnl_eval_ms": lnl_ms, "up_pass_ms": up_ms, "all_edge_derivatives_ms": d_ms, "n_edges": int(len(nodes)),
            "sweep_ms_one_derivative_pass": lnl_ms + up_ms + d_ms,
            "site_node_updates_per_s_lnl": n_int * args.patterns / lnl_ms * 1e3,
            "newton_sweeps": res["sweeps"], "inner_iterations": args.inner, "newton_wall_s": float(wall[0]),
            "newton_trace": res["trace"], "max_abs_dlnl": float(np.abs(d[:, 1]).max()),
            "timing": "CUDA events between barriers, max over ranks; Newton: host wall clock, max over ranks"}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
