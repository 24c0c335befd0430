"""
Timings of the non-headline BASELINE configs (parity-test shapes, not bench.py lines):

  cfg3  LG+G4 protein, 500 taxa x 100k patterns: lnL and all-edge derivatives
  cfg4  GY94+G4 codon (61 states), 100 taxa x 50k patterns: lnL
  cfg5  GTR+G4, 2000 taxa x 62.5k patterns (one GPU's shard of 500k / 8): down pass + up pass + all-edge derivatives

    python tools/bench_configs.py [cfg3] [cfg4] [cfg5] [--reps 3]
Prints one JSON line per config with CUDA-event timings and algorithmic rates (SURVEY.md 8(d) figures).
"""
import gc
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import phylo_utils_b200 as phy  # noqa: E402
from phylo_utils_b200.tree import random_tree  # noqa: E402


def timed(fn, reps):
    # contexts of earlier configs may still be waiting for the cycle collector: their cudaFree must not land in a
    # timed region (it did: 10.5 vs 32-58 ms for the same derivative pass depending on what had run before)
    gc.collect()
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        out = fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps, out


def build(n_taxa, n_pat, n_states, model, seed, up=False, mode="auto"):
    rng = np.random.default_rng(seed)
    tree = random_tree(n_taxa, seed)
    names = [l.taxon.label for l in tree.leaf_node_iter()]
    lut = np.vstack([np.eye(n_states)[::-1], np.ones((1, n_states))])
    codes = rng.integers(0, n_states, size=(n_taxa, n_pat)).astype(np.uint8)
    codes[rng.random((n_taxa, n_pat)) < 0.01] = n_states
    tm = phy.TreeModel(up_partials=up, mode=mode)
    tm.set_tree(tree)
    tm.set_tip_codes(codes, lut, {n: i for i, n in enumerate(names)})
    tm.set_rate_model(phy.rate_models.GammaRateModel(4, 0.5))
    tm.set_substitution_model(model)
    tm.initialise()
    return tm


def report(name, tm, n_taxa, n_pat, A, reps, derivs=False):
    K = 4
    b_node = 2 * K * A * 8 + 16
    flops_node = K * (4 * A * A + A)
    ms, lnl = timed(lambda: (tm.compute_partials(), tm.lnl())[1], reps)
    prune_ms, _ = timed(lambda: tm.compute_partials(), reps)
    nodes = (n_taxa - 2) * n_pat
    out = {"config": name, "taxa": n_taxa, "patterns": n_pat, "states": A, "lnl": lnl, "lnl_ms": ms, "prune_ms": prune_ms,
           "site_node_updates_per_s": nodes / ms * 1e3, "algorithmic_GBs": nodes * b_node / ms / 1e6,
           "fp64_TFLOPs": nodes * flops_node / ms / 1e9, "mma_disabled": bool(os.environ.get("PHB_DISABLE_MMA"))}
    if derivs:
        up_ms, _ = timed(lambda: tm.compute_up_partials(), reps)
        all_nodes = np.arange(2 * n_taxa - 2)
        all_lengths = tm.lengths_above(all_nodes)        # resolved once, as the Newton driver does
        d_ms, d = timed(lambda: tm.edge_derivatives(all_nodes, all_lengths), reps)
        out.update(up_pass_ms=up_ms, all_edge_derivatives_ms=d_ms, n_edges=int(len(all_nodes)),
                   sweep_ms=ms + up_ms + d_ms, max_abs_dlnl=float(np.abs(d[:, 1]).max()))
        if "--newton" in sys.argv:
            from phylo_utils_b200.optimise import optimise_branch_lengths
            t0 = time.perf_counter()
            res = optimise_branch_lengths(tm, max_sweeps=3, inner_iterations=2, tol=0.0)
            out.update(newton_sweeps=res["sweeps"], newton_wall_s=time.perf_counter() - t0, newton_trace=res["trace"])
    print(json.dumps(out), flush=True)


def main():
    which = [a for a in sys.argv[1:] if a.startswith("cfg")] or ["cfg1", "cfg3", "cfg4", "cfg5"]
    reps = int(sys.argv[sys.argv.index("--reps") + 1]) if "--reps" in sys.argv else 3
    if "cfg1" in which:
        # the reference's own test-sized case: 10 taxa x 1000 patterns, launch-latency bound.  Wall clock per
        # evaluation through TreeModel (new branch lengths -> lnL on the host), 200 evaluations each way.
        model = phy.substitution_models.GTR([6., 5., 4., 3., 2., 1.], [0.1, 0.2, 0.3, 0.4])
        out = {"config": "cfg1 GTR+G4 10x1000"}
        for label, kw in (("partials_stored", {}), ("lnl_only", {"store_partials": False})):
            rng = np.random.default_rng(1)
            tree = random_tree(10, 1)
            names = [l.taxon.label for l in tree.leaf_node_iter()]
            lut = np.vstack([np.eye(4)[::-1], np.ones((1, 4))])
            codes = rng.integers(0, 5, size=(10, 1000)).astype(np.uint8)
            tm = phy.TreeModel(**kw)
            tm.set_tree(tree)
            tm.set_tip_codes(codes, lut, {n: i for i, n in enumerate(names)})
            tm.set_rate_model(phy.rate_models.GammaRateModel(4, 0.5))
            tm.set_substitution_model(model)
            tm.initialise()

            def one():
                tm.compute_partials()      # new branch lengths (and, when partials are stored, P build + post-order pass)
                return tm.lnl()
            for _ in range(20):
                lnl = one()
            t0 = time.perf_counter()
            for _ in range(200):
                lnl = one()
            out[label + "_us_per_eval"] = (time.perf_counter() - t0) / 200 * 1e6
            out["lnl"] = lnl
        print(json.dumps(out), flush=True)
    if "cfg3" in which:
        tm = build(500, 100000, 20, phy.substitution_models.LG(), 3, up=True)
        report("cfg3 LG+G4 500x100k", tm, 500, 100000, 20, reps, derivs=True)
        del tm
    if "cfg4" in which:
        from phylo_utils_b200.substitution_models.codon import f3x4
        model = phy.substitution_models.GY94(2.0, 0.2, f3x4(np.random.default_rng(4).dirichlet(np.ones(4) * 5, size=3)))
        tm = build(100, 50000, 61, model, 4)
        report("cfg4 GY94+G4 100x50k", tm, 100, 50000, 61, reps)
        del tm
    if "cfg5" in which:
        model = phy.substitution_models.GTR([6., 5., 4., 3., 2., 1.], [0.1, 0.2, 0.3, 0.4])
        tm = build(2000, 62500, 4, model, 5, up=True)
        report("cfg5 GTR+G4 2000x62.5k (1/8 shard)", tm, 2000, 62500, 4, reps, derivs=True)


def custom():
    """--shape N,S: a 4-state GTR+G4 problem of that size with the derivative passes (tuning aid)."""
    n, s = (int(v) for v in sys.argv[sys.argv.index("--shape") + 1].split(","))
    reps = int(sys.argv[sys.argv.index("--reps") + 1]) if "--reps" in sys.argv else 3
    model = phy.substitution_models.GTR([6., 5., 4., 3., 2., 1.], [0.1, 0.2, 0.3, 0.4])
    tm = build(n, s, 4, model, 5, up=True)
    report("custom GTR+G4 {}x{}".format(n, s), tm, n, s, 4, reps, derivs=True)


if __name__ == "__main__":
    if "--shape" in sys.argv:
        custom()
        sys.exit(0)
    main()
