"""
A/B timing of the lnL-only evaluation at the headline shape: the one-pattern-per-lane resident walk
(PHB_RESIDENT_V1=1) against the two-patterns-per-lane pair walk (default), tips resident and from host
(8-bit and 4-bit packed codes).

    python tools/ab_lnl.py [--taxa 1000] [--patterns 1000000] [--evals 5] [--variants v1,pair] [--chunks 16]
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import phylo_utils_b200 as phy  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--taxa", type=int, default=1000)
    ap.add_argument("--patterns", type=int, default=1000000)
    ap.add_argument("--evals", type=int, default=5)
    ap.add_argument("--seed", type=int, default=2)
    ap.add_argument("--variants", default="v1,pair")
    ap.add_argument("--chunks", type=int, default=16)
    ap.add_argument("--no-host", action="store_true")
    args = ap.parse_args()
    import torch
    tree, names = bench.make_tree(args.taxa, args.seed)
    trav = phy.traversal.Traversal(phy.utils.deepcopy_tree(tree))
    model = phy.substitution_models.GTR(bench.GTR_RATES, bench.GTR_FREQS)
    rate = phy.rate_models.GammaRateModel(4, 0.5)
    codes_host = torch.from_numpy(bench.make_codes(args.taxa, args.patterns, args.seed)).pin_memory()
    packed_host = torch.from_numpy(phy.LikelihoodEngine.pack_codes(codes_host.numpy())).pin_memory()
    eng = phy.LikelihoodEngine(args.taxa, args.patterns, 4, 4, store_partials=False)
    rows = trav.locality_order()
    eng.set_schedule(rows)
    tip_nodes = np.asarray([trav.names[n] for n in names], dtype=np.int32)
    e = model.eigen
    eng.set_model(e.evecs, e.evals, np.ascontiguousarray(e.ivecs), model.freqs, rate.rates, rate.weights)
    lengths = np.asarray([[trav.brlens[(int(p), int(a))], trav.brlens[(int(p), int(b))]] for p, a, b in rows])
    a, b = trav.root_edge
    root_len = trav.brlens[(a, b)]

    def timed(fn):
        for _ in range(2):
            out = fn()
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(args.evals):
            out = fn()
        t1.record()
        torch.cuda.synchronize()
        return t0.elapsed_time(t1) / args.evals, out

    for variant in args.variants.split(","):
        if variant == "v1":
            os.environ["PHB_RESIDENT_V1"] = "1"
        else:
            os.environ.pop("PHB_RESIDENT_V1", None)
        eng.set_tips(codes_host.cuda(), bench.dna_lut(), tip_nodes)

        def resident():
            eng.set_edge_lengths(lengths)
            return eng.lnl_resident(a, b, root_len)[0]
        ms, lnl = timed(resident)
        print("{:5s} resident         {:8.3f} ms  lnL {!r}".format(variant, ms, lnl), flush=True)
        if args.no_host:
            continue

        def from_host():
            eng.set_edge_lengths(lengths)
            return eng.lnl_from_host(codes_host.numpy(), a, b, root_len, n_chunks=args.chunks)[0]
        ms, lnl = timed(from_host)
        print("{:5s} from host u8     {:8.3f} ms  lnL {!r}".format(variant, ms, lnl), flush=True)
        if variant != "v1":
            def from_packed():
                eng.set_edge_lengths(lengths)
                return eng.lnl_from_host(packed_host.numpy(), a, b, root_len, n_chunks=args.chunks, packed=True)[0]
            ms, lnl = timed(from_packed)
            print("{:5s} from host packed {:8.3f} ms  lnL {!r}".format(variant, ms, lnl), flush=True)


if __name__ == "__main__":
    main()
