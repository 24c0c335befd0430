"""
Bounded measurements of the non-headline BASELINE configs (parity-test shapes; bench.py embeds them as `configs`):

  cfg1  GTR+G4, 10 taxa x 1000 sites - the reference's own test size (golden case, launch-latency bound)
  cfg3  LG+G4 protein, 500 taxa x 100k patterns: lnL, pre-order pass, all-edge derivatives (FP64 tensor cores)
  cfg4  GY94+G4 codon (61 states), 100 taxa x 50k patterns: lnL (FP64 tensor cores), derivatives for the parity flag
  cfg5  GTR+G4, 2000 taxa x 500k patterns over 8 GPUs: down pass + pre-order pass + all-edge derivatives, Newton sweeps
        (world == 1: one GPU's shard of 62.5k patterns; world == 8: the config as named, through ShardedTreeModel)

    python tools/bench_configs.py [cfg1] [cfg3] [cfg4] [cfg5] [--reps 3] [--newton]
    python -m torch.distributed.run --nproc-per-node 8 ... tools/bench_configs.py cfg5

Every record carries: the config's value, the kernel times (CUDA events on the launching stream), the fraction of the
resource that bounds it (HBM GB/s over the measured copy peak, or algorithmic fp64 flops over the measured DMMA
rate) and a parity flag from a size-independent property (pulley principle: the derivative pass must reproduce the
root lnL on every edge; cfg1: the reference-generated golden value).
"""
import gc
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import phylo_utils_b200 as phy  # noqa: E402
from phylo_utils_b200.tree import random_tree  # noqa: E402

GTR = lambda: phy.substitution_models.GTR([6., 5., 4., 3., 2., 1.], [0.1, 0.2, 0.3, 0.4])   # noqa: E731


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


def synthetic_codes(n_taxa, n_pat, n_states, seed):
    rng = np.random.default_rng(seed)
    codes = rng.integers(0, n_states, size=(n_taxa, n_pat)).astype(np.uint8)
    codes[rng.random((n_taxa, n_pat)) < 0.01] = n_states
    return codes, np.vstack([np.eye(n_states)[::-1], np.ones((1, n_states))])


def build(n_taxa, n_pat, n_states, model, seed, up=False, mode="auto", device=0):
    tree = random_tree(n_taxa, seed)
    names = [l.taxon.label for l in tree.leaf_node_iter()]
    codes, lut = synthetic_codes(n_taxa, n_pat, n_states, seed)
    tm = phy.TreeModel(device=device, up_partials=up, mode=mode)
    tm.set_tree(tree)
    tm.set_tip_codes(codes, lut, {n: i for i, n in enumerate(names)})
    tm.set_rate_model(phy.rate_models.GammaRateModel(4, 0.5))
    tm.set_substitution_model(model)
    tm.initialise()
    return tm


def measure(name, tm, n_taxa, n_pat, A, reps, peaks, derivs=True, newton=False):
    """lnL evaluation (+ pre-order and derivative passes) of an initialised TreeModel; `peaks` = (hbm GB/s, dmma TF/s, dfma TF/s)."""
    K = 4
    peak_hbm, peak_dmma, peak_dfma = peaks
    b_node = 2 * K * A * 8 + 16
    flops_node = K * (4 * A * A + A)
    nodes = (n_taxa - 2) * n_pat
    ms, lnl = timed(lambda: (tm.compute_partials(), tm.lnl())[1], reps)
    prune_ms, _ = timed(lambda: tm.compute_partials(), reps)
    tensor = A in (20, 61)
    out = {"config": name, "taxa": n_taxa, "patterns": n_pat, "states": A, "lnl": lnl,
           "value": 1e3 / ms, "unit": "lnL evals/s", "lnl_ms": ms, "prune_kernel_ms": prune_ms,
           "site_node_updates_per_s": nodes / ms * 1e3,
           "hbm": {"algorithmic_gbs": nodes * b_node / prune_ms / 1e6, "frac_of_measured_peak": nodes * b_node / prune_ms / 1e6 / peak_hbm},
           "fp64": {"algorithmic_tflops": nodes * flops_node / prune_ms / 1e9,
                    "frac_of_measured_peak": nodes * flops_node / prune_ms / 1e9 / (peak_dmma if tensor else peak_dfma),
                    "pipe": "DMMA (mma.sync m8n8k4 f64)" if tensor else "DFMA",
                    "note": "algorithmic flops (SURVEY.md 8(d)); tip operands are table look-ups, so the executed share is lower"}}
    if derivs:
        up_ms, _ = timed(lambda: tm.compute_up_partials(), reps)
        all_nodes = phy.optimise.edge_nodes(tm.traversal)
        all_lengths = tm.lengths_above(all_nodes)        # resolved once, as the Newton driver does
        tm.compute_up_partials()
        t0 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0[0].record()
        d = tm.edge_derivatives(all_nodes, all_lengths)   # first pass after a pre-order pass (leaves the sum tables at A = 20 / 61)
        t0[1].record()
        torch.cuda.synchronize()
        d_ms, d = timed(lambda: tm.edge_derivatives(all_nodes, all_lengths), reps)
        st_bytes = float(len(all_nodes)) * n_pat * K * A * 8
        pulley = float(np.max(np.abs(d[:, 0] - lnl)) / abs(lnl))
        out.update(up_pass_ms=up_ms, first_derivative_pass_ms=t0[0].elapsed_time(t0[1]), derivative_pass_ms=d_ms,
                   n_edges=int(len(all_nodes)), sweep_ms=ms + up_ms + d_ms,
                   derivative_pass_hbm={"bytes": st_bytes, "gbs": st_bytes / d_ms / 1e6, "frac_of_measured_peak": st_bytes / d_ms / 1e6 / peak_hbm},
                   parity={"check": "pulley principle: lnL from the derivative pass on every edge vs the root lnL",
                           "max_rel_err": pulley, "finite": bool(np.all(np.isfinite(d))), "ok": bool(pulley <= 1e-10 and np.all(np.isfinite(d)))})
        if newton:
            from phylo_utils_b200.optimise import optimise_branch_lengths
            t0 = time.perf_counter()
            res = optimise_branch_lengths(tm, max_sweeps=3, inner_iterations=2, tol=0.0)
            out.update(newton_sweeps=res["sweeps"], newton_wall_s=time.perf_counter() - t0, newton_trace=res["trace"],
                       newton_monotone=bool(np.all(np.diff(res["trace"]) >= 0)))
    return out


def cfg1(device=0):
    """The reference's own test-sized case through TreeModel: wall clock per evaluation (new branch lengths -> lnL on
    the host), and the value against the committed output of the unmodified reference."""
    from phylo_utils_b200.alignment.alignment import SeqRecord
    with np.load(os.path.join(ROOT, "tests", "golden", "cfg1_gtr_g4.npz")) as z:
        g = {k: z[k] for k in z.files}
    records = [SeqRecord(str(n), bytes(row).decode("ascii")) for n, row in zip(g["names"], g["seqs"])]
    want = float(g["total_lnl"])
    out = {"config": "cfg1 GTR+G4 10 taxa x 1000 sites (tests/golden/cfg1_gtr_g4.npz)", "unit": "lnL evals/s"}
    for label, kw in (("partials_stored", {}), ("lnl_only", {"store_partials": False})):
        tm = phy.TreeModel(device=device, **kw)
        tm.set_tree(phy.tree.parse_newick(str(g["newick"])))
        tm.set_alignment(records, int(g["alphabet"]))
        tm.set_rate_model(phy.rate_models.GammaRateModel(4, 0.5))
        tm.set_substitution_model(GTR())
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
        out[label + "_rel_err_vs_reference"] = abs(lnl - want) / abs(want)
    out["value"] = 1e6 / out["lnl_only_us_per_eval"]
    out["bound"] = "launch latency (4 launches, 2 small copies, 1 synchronisation per evaluation)"
    out["lnl"] = lnl
    out["parity"] = {"check": "total lnL vs the unmodified reference's output", "max_rel_err": max(
        out["partials_stored_rel_err_vs_reference"], out["lnl_only_rel_err_vs_reference"]),
        "ok": bool(max(out["partials_stored_rel_err_vs_reference"], out["lnl_only_rel_err_vs_reference"]) <= 1e-10)}
    return out


def cfg5_sharded(world, rank, device, reps, peaks, n_taxa=2000, n_pat=500000, newton=True):
    """cfg5 as named, through ShardedTreeModel: every rank synthesises its own shard; times are max over ranks."""
    import torch.distributed as dist
    from phylo_utils_b200.parallel import ShardedTreeModel, shard_bounds
    from phylo_utils_b200.optimise import optimise_branch_lengths, edge_nodes
    dev = torch.device("cuda", device)
    tree = random_tree(n_taxa, 5)
    names = [l.taxon.label for l in tree.leaf_node_iter()]
    lo, hi = shard_bounds(n_pat, rank, world)
    codes, lut = synthetic_codes(n_taxa, hi - lo, 4, 5000 + rank)
    tm = ShardedTreeModel(device=device, up_partials=True)
    tm.set_tree(tree)
    tm.set_local_tip_codes(codes, lut, {n: i for i, n in enumerate(names)}, n_pat)
    tm.set_rate_model(phy.rate_models.GammaRateModel(4, 0.5))
    tm.set_substitution_model(GTR())
    tm.initialise()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_all(fn):
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
    lnl_ms, lnl = timed_all(lambda: (tm.compute_partials(), tm.lnl())[1])
    up_ms, _ = timed_all(lambda: tm.compute_up_partials())
    d_ms, d = timed_all(lambda: tm.edge_derivatives(nodes, lengths))
    pulley = float(np.max(np.abs(d[:, 0] - lnl)) / abs(lnl))
    st_bytes = float(len(nodes)) * (hi - lo) * 4 * 4 * 8
    out = {"config": "cfg5 GTR+G4 {} taxa x {} patterns over {} GPU(s) (ShardedTreeModel)".format(n_taxa, n_pat, world),
           "taxa": n_taxa, "patterns": n_pat, "patterns_per_gpu": hi - lo, "states": 4, "lnl": lnl,
           "lnl_ms": lnl_ms, "up_pass_ms": up_ms, "derivative_pass_ms": d_ms, "sweep_ms": lnl_ms + up_ms + d_ms,
           "value": 1e3 / (lnl_ms + up_ms + d_ms), "unit": "derivative sweeps/s (down + pre-order + all {} edges)".format(len(nodes)),
           "n_edges": int(len(nodes)), "collectives": tm.collectives,
           "derivative_pass_hbm": {"bytes_per_gpu": st_bytes, "gbs_per_gpu": st_bytes / d_ms / 1e6, "frac_of_measured_peak": st_bytes / d_ms / 1e6 / peaks[0]},
           "parity": {"check": "pulley principle on the all-reduced derivative sums", "max_rel_err": pulley, "ok": bool(pulley <= 1e-10)}}
    if newton:
        t0 = time.perf_counter()
        res = optimise_branch_lengths(tm, max_sweeps=3, inner_iterations=2, tol=0.0)
        out.update(newton_sweeps=res["sweeps"], newton_wall_s=time.perf_counter() - t0, newton_trace=res["trace"],
                   newton_monotone=bool(np.all(np.diff(res["trace"]) >= 0)))
    return out


def sub_records(world, rank, device, peak_hbm, peak_dmma, peak_dfma, which=("cfg1", "cfg3", "cfg4", "cfg5"), reps=3, newton=True):
    """What bench.py embeds.  world == 1: all four on this GPU (cfg5 = one 1/8 shard); world == 8: cfg5 as named."""
    peaks = (peak_hbm, peak_dmma, peak_dfma)
    out = {}

    def release():
        gc.collect()
        torch.cuda.empty_cache()
    if world > 1:
        if "cfg5" in which:
            out["cfg5"] = cfg5_sharded(world, rank, device, reps, peaks, newton=newton)
        return out
    if "cfg1" in which:
        out["cfg1"] = cfg1(device)
    if "cfg3" in which:
        tm = build(500, 100000, 20, phy.substitution_models.LG(), 3, up=True, device=device)
        out["cfg3"] = measure("cfg3 LG+G4 500 taxa x 100k patterns", tm, 500, 100000, 20, reps, peaks, newton=newton)
        del tm
        release()
    if "cfg4" in which:
        from phylo_utils_b200.substitution_models.codon import f3x4
        model = phy.substitution_models.GY94(2.0, 0.2, f3x4(np.random.default_rng(4).dirichlet(np.ones(4) * 5, size=3)))
        tm = build(100, 50000, 61, model, 4, up=True, device=device)
        out["cfg4"] = measure("cfg4 GY94+G4 (61 states) 100 taxa x 50k codons", tm, 100, 50000, 61, reps, peaks)
        del tm
        release()
    if "cfg5" in which:
        tm = build(2000, 62500, 4, GTR(), 5, up=True, device=device)
        out["cfg5"] = measure("cfg5 GTR+G4 2000 taxa x 62.5k patterns (one GPU's 1/8 shard of 500k)", tm, 2000, 62500, 4, reps, peaks, newton=newton)
        del tm
        release()
    return out


def main():
    from phylo_utils_b200.engine import fp64_peak
    which = tuple(a for a in sys.argv[1:] if a.startswith("cfg")) or ("cfg1", "cfg3", "cfg4", "cfg5")
    reps = int(sys.argv[sys.argv.index("--reps") + 1]) if "--reps" in sys.argv else 3
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    device = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(device)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", device))
    peak_hbm = 6531.6
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        peak_hbm = float(json.load(open(path))["hbm_gbs"])
    recs = sub_records(world, rank, device, peak_hbm, fp64_peak(device, tensor=True), fp64_peak(device), which, reps,
                       newton="--newton" in sys.argv)
    if rank == 0:
        for rec in recs.values():
            print(json.dumps(rec), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


def custom():
    """--shape N,S: a 4-state GTR+G4 problem of that size with the derivative passes (tuning aid)."""
    from phylo_utils_b200.engine import fp64_peak
    n, s = (int(v) for v in sys.argv[sys.argv.index("--shape") + 1].split(","))
    reps = int(sys.argv[sys.argv.index("--reps") + 1]) if "--reps" in sys.argv else 3
    tm = build(n, s, 4, GTR(), 5, up=True)
    print(json.dumps(measure("custom GTR+G4 {}x{}".format(n, s), tm, n, s, 4, reps, (6531.6, fp64_peak(0, True), fp64_peak(0)))), flush=True)


if __name__ == "__main__":
    if "--shape" in sys.argv:
        custom()
        sys.exit(0)
    main()
