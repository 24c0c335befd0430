"""
Times the post-order walk that stores every node block (dna_pair_store_kernel, PHB_MODE_RESIDENT) at the cfg5-shard and
cfg2 shapes, CUDA events around the launch sequence.  Build variants are compared by pointing PHB_LIBRARY at them:

    PHB_LIBRARY=phylo_utils_b200/libphylo_b200_alt.so python tools/store_probe.py --tag alt
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import phylo_utils_b200 as phy  # noqa: E402
from phylo_utils_b200 import _lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shapes", default="2000x62500,1000x1000000")
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--tag", default="")
    args = ap.parse_args()
    for shape in args.shapes.split(","):
        taxa, patterns = (int(x) for x in shape.split("x"))
        tree = phy.tree.random_tree(taxa, 2)
        names = [lf.taxon.label for lf in tree.leaf_node_iter()]
        trav = phy.traversal.Traversal(phy.utils.deepcopy_tree(tree))
        model = phy.substitution_models.GTR(bench.GTR_RATES, bench.GTR_FREQS)
        rate = phy.rate_models.GammaRateModel(4, 0.5)
        codes = torch.from_numpy(bench.make_codes(taxa, 0, patterns, 2)).cuda()
        eng = phy.LikelihoodEngine(taxa, patterns, 4, 4)
        rows = trav.locality_order()
        eng.set_schedule(rows)
        eng.set_tips(codes, bench.dna_lut(), np.asarray([trav.names[n] for n in names], dtype=np.int32))
        e = model.eigen
        eng.set_model(e.evecs, e.evals, np.ascontiguousarray(e.ivecs), model.freqs, rate.rates, rate.weights)
        lengths = np.asarray([[trav.brlens[(int(p), int(a))], trav.brlens[(int(p), int(b))]] for p, a, b in rows])
        a, b = trav.root_edge
        eng.set_edge_lengths(lengths)
        eng.build_pmatrices()
        for _ in range(2):
            eng.compute_partials(_lib.PHB_MODE_RESIDENT)
        lnl = eng.root_lnl(a, b, trav.brlens[(a, b)])[0]
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.reps):
            eng.compute_partials(_lib.PHB_MODE_RESIDENT)
        e1.record()
        torch.cuda.synchronize()
        print(json.dumps({"tag": args.tag, "library": os.path.basename(_lib.LIB_PATH), "shape": shape, "store_walk_ms": round(e0.elapsed_time(e1) / args.reps, 4),
                          "lnl": lnl}), flush=True)
        del eng, codes
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
