"""
Minimal driver for ncu: builds the headline workload at a given size and runs the evaluation a few times.

    python tools/profile_prune.py --taxa 1000 --patterns 300000 --evals 3 [--mode tile|level] [--states 4]
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import phylo_utils_b200 as phy  # noqa: E402
from phylo_utils_b200 import _lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--taxa", type=int, default=1000)
    ap.add_argument("--patterns", type=int, default=300000)
    ap.add_argument("--evals", type=int, default=3)
    ap.add_argument("--mode", default="tile")
    ap.add_argument("--seed", type=int, default=2)
    ap.add_argument("--lnl-only", action="store_true")
    args = ap.parse_args()
    import torch
    tree = phy.tree.random_tree(args.taxa, args.seed)
    names = [lf.taxon.label for lf in tree.leaf_node_iter()]
    trav = phy.traversal.Traversal(phy.utils.deepcopy_tree(tree))
    model = phy.substitution_models.GTR(bench.GTR_RATES, bench.GTR_FREQS)
    rate = phy.rate_models.GammaRateModel(4, 0.5)
    codes = torch.from_numpy(bench.make_codes(args.taxa, 0, args.patterns, args.seed)).cuda()
    eng = phy.LikelihoodEngine(args.taxa, args.patterns, 4, 4, store_partials=not args.lnl_only)
    mode = {"tile": _lib.PHB_MODE_TILE, "level": _lib.PHB_MODE_LEVEL, "resident": _lib.PHB_MODE_RESIDENT}[args.mode]
    if mode == _lib.PHB_MODE_LEVEL:
        rows, off = trav.level_order()
        eng.set_schedule(rows, off)
    else:
        rows = trav.locality_order()
        eng.set_schedule(rows)
    eng.set_tips(codes, bench.dna_lut(), np.asarray([trav.names[n] for n in names], dtype=np.int32))
    e = model.eigen
    eng.set_model(e.evecs, e.evals, np.ascontiguousarray(e.ivecs), model.freqs, rate.rates, rate.weights)
    lengths = np.asarray([[trav.brlens[(int(p), int(a))], trav.brlens[(int(p), int(b))]] for p, a, b in rows])
    a, b = trav.root_edge
    for _ in range(args.evals):
        eng.set_edge_lengths(lengths)
        if args.lnl_only:
            lnl = eng.lnl_resident(a, b, trav.brlens[(a, b)])[0]
            continue
        eng.build_pmatrices()
        eng.compute_partials(mode)
        lnl = eng.root_lnl(a, b, trav.brlens[(a, b)])[0]
    print("lnL", lnl, "launches", eng.launch_count)


if __name__ == "__main__":
    main()
