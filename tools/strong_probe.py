"""
Strong-scaling probe on ONE GPU: the lnL-only evaluation of a 1000-taxon GTR+G4 problem at the per-GPU shard sizes of
the headline alignment (1M / N patterns), through TreeModel (the path ShardedTreeModel runs on every rank).
Prints per shard size: the kernel sequence alone (stream-ordered, no host round trips) and the full step
(compute_partials + lnl with one synchronisation).  Tuning switches are read once per process, so A/B runs are
separate invocations:

    PHB_PAIR_GRID=2 python tools/strong_probe.py --sizes 125000,250000,500000
    PHB_PAIR_PPT=4  python tools/strong_probe.py --sizes 125000
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import phylo_utils_b200 as phy  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--taxa", type=int, default=1000)
    ap.add_argument("--sizes", default="125000,250000,500000,1000000")
    ap.add_argument("--evals", type=int, default=20)
    ap.add_argument("--tag", default="")
    args = ap.parse_args()
    tree = phy.tree.random_tree(args.taxa, 2)
    names = {lf.taxon.label: i for i, lf in enumerate(tree.leaf_node_iter())}
    lut = np.vstack([np.eye(4)[::-1], np.ones((1, 4))])
    knobs = {k: v for k, v in os.environ.items() if k.startswith("PHB_")}
    for size in (int(s) for s in args.sizes.split(",")):
        rng = np.random.default_rng(size)
        codes = rng.integers(0, 5, size=(args.taxa, size), dtype=np.uint8)
        tm = phy.TreeModel(store_partials=False)
        tm.set_tree(tree)
        tm.set_tip_codes(codes, lut, names)
        tm.set_rate_model(phy.rate_models.GammaRateModel(4, 0.5))
        tm.set_substitution_model(phy.substitution_models.GTR([6., 5., 4., 3., 2., 1.], [0.1, 0.2, 0.3, 0.4]))
        tm.initialise()
        a, b = tm.traversal.root_edge
        length = tm.traversal.brlens[(a, b)]

        def step():
            tm.compute_partials()
            return tm.lnl()
        for _ in range(3):
            lnl = step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.evals):
            tm.engine.lnl_resident_async(a, b, length)
        e1.record()
        torch.cuda.synchronize()
        kernel_ms = e0.elapsed_time(e1) / args.evals
        e0.record()
        for _ in range(args.evals):
            lnl = step()
        e1.record()
        torch.cuda.synchronize()
        step_ms = e0.elapsed_time(e1) / args.evals
        ideal = 13.78 * size / 1e6
        print(json.dumps({"tag": args.tag, "knobs": knobs, "patterns": size, "kernel_ms": round(kernel_ms, 4), "step_ms": round(step_ms, 4),
                          "tiles64": (size + 63) // 64, "vs_linear_from_1M_at_13.78ms": round(ideal / step_ms, 3), "lnl": lnl}), flush=True)
        del tm


if __name__ == "__main__":
    main()
