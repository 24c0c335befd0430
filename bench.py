#!/usr/bin/env python
"""
Headline benchmark: full-tree lnL evaluations per second, GTR+Gamma4 nucleotide model,
1,000 taxa x 1,000,000 site patterns per GPU (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one evaluation: branch lengths -> all transition matrices (one kernel) -> post-order
pruning over all N-2 internal nodes -> root combine, Gamma mixture, log, weighted reduction -> scalar lnL
on the host.

  value   evaluations/s with the tip codes already resident in HBM (CUDA events, max over ranks)
  e2e     the same evaluation through the public TreeModel/engine API starting from HOST buffers:
          every step copies the tip codes (N x S bytes, pinned) and the branch lengths to the device and reads
          the lnL back
  roofline  for the dominant kernel (the pruning kernel), timed on its own with CUDA events
  cpu_baseline  the CPU oracle (C restatement of the reference engine, OpenMP over all host cores) on a
          bounded pattern sample of the same tree, extrapolated linearly in the pattern count

Multi-GPU: site patterns are independent, so each rank owns its own block of 1M patterns of one alignment
(weak scaling); the only exchange is an NCCL all-reduce of the scalar lnL.  value = total patterns evaluated
per second / 1e6 = "1M-pattern evaluations per second", summed over ranks.

--impl reference times the reference's CPU algorithm (the oracle port - the reference itself is Python +
numba and cannot travel to the GPU box) with all host threads on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GTR_RATES = [6., 5., 4., 3., 2., 1.]
GTR_FREQS = [0.1, 0.2, 0.3, 0.4]
ALPHA = 0.5
NCAT = 4


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--taxa", type=int, default=1000)
    ap.add_argument("--patterns", type=int, default=1000000, help="site patterns per GPU")
    ap.add_argument("--cpu-patterns", type=int, default=20000, help="pattern sample for the CPU baseline")
    ap.add_argument("--seed", type=int, default=2)
    ap.add_argument("--mode", choices=["auto", "tile", "level", "resident"], default="auto",
                    help="how the rows are walked when per-node partials are stored (--store-partials)")
    ap.add_argument("--store-partials", action="store_true",
                    help="headline path keeps every node's partials in HBM (TreeModel.partials / derivatives); "
                         "default is the pure lnL evaluation with the operand-resident kernel")
    ap.add_argument("--chunks", type=int, default=64, help="host->device pipeline depth of the e2e path (16: 68.6, 32: 68.6, 64: 69.9, 128: 69.5 evals/s)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-stored", action="store_true", help="skip the extra 'partials stored' measurements")
    return ap.parse_args()


# --------------------------------------------------------------------------------------------------
# synthetic workload: seeded random topology, branch lengths U(0.01, 0.3), iid uniform
# states with 1 % fully ambiguous codes; at these sizes every column is a distinct pattern
# --------------------------------------------------------------------------------------------------
def make_tree(n_taxa, seed):
    import phylo_utils_b200 as phy
    tree = phy.tree.random_tree(n_taxa, seed)
    names = [lf.taxon.label for lf in tree.leaf_node_iter()]
    return tree, names


def make_codes(n_taxa, n_pat, seed, rank=0):
    rng = np.random.default_rng([seed, rank])
    codes = rng.integers(0, 4, size=(n_taxa, n_pat), dtype=np.uint8)
    gaps = rng.random((n_taxa, n_pat), dtype=np.float32) < 0.01
    codes[gaps] = 4
    return codes


def dna_lut():
    # rows in lexicographic rank order of the 0/1 vectors: T, G, C, A, then the all-ones gap row
    return np.vstack([np.eye(4)[::-1], np.ones((1, 4))])


def algorithmic_bytes(n_taxa, n_pat, K=NCAT, A=4):
    """S * [(N-2) * (2*K*A*8 + 16) + N + 16]; the pruning kernel's share excludes the root's 16."""
    b_node = 2 * K * A * 8 + 16
    prune = n_pat * ((n_taxa - 2) * b_node + n_taxa)
    return prune, prune + n_pat * 16


# --------------------------------------------------------------------------------------------------
class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self._stop, self._thread = index, [], threading.Event(), None

    def _nvml_handle(self):
        """NVML handle of CUDA device ``index`` (by UUID: CUDA_VISIBLE_DEVICES may renumber), or None."""
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            uuid = str(torch.cuda.get_device_properties(self.index).uuid)
            if not uuid.startswith("GPU-"):
                uuid = "GPU-" + uuid
            try:
                return pynvml, pynvml.nvmlDeviceGetHandleByUUID(uuid)
            except TypeError:
                return pynvml, pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
        except Exception:
            return None

    def _run(self):
        nv = self._nvml_handle()
        if nv is not None:
            # in-process NVML queries: a sample every 10 ms instead of one nvidia-smi process every few hundred
            pynvml, h = nv
            bits = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))
            try:
                smax = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
                while not self._stop.is_set():
                    sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                    try:
                        mask = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                    except Exception:
                        mask = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                    try:
                        watts = pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0
                    except Exception:
                        watts = 0.0
                    self.rows.append([str(sm), str(smax), str(watts)] + ["Active" if mask & b else "Not Active" for b, _ in bits])
                    self._stop.wait(0.01)
                return
            except Exception:
                pass                                   # fall through to nvidia-smi
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                      "--format=csv,noheader,nounits"], stdout=subprocess.PIPE,
                                     stderr=subprocess.DEVNULL, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.splitlines()[0].split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._thread.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm, smax, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                smax = float(r[1])
            except (ValueError, IndexError):
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm), "sm_min_mhz": float(min(sm)) if sm else None}


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(workload):
    """dram bytes per launch of the dominant kernel from the committed ncu --set full summary, or None."""
    path = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(path):
        with open(path) as fh:
            return json.load(fh).get(workload)
    return None


# --------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference engine
# --------------------------------------------------------------------------------------------------
def cpu_eval_factory(n_taxa, n_pat, seed):
    import phylo_utils_b200 as phy
    from oracle import oracle
    tree, names = make_tree(n_taxa, seed)
    clone = phy.utils.deepcopy_tree(tree)
    trav = phy.traversal.Traversal(clone)
    codes = make_codes(n_taxa, n_pat, seed)
    lut = dna_lut()
    model = phy.substitution_models.GTR(GTR_RATES, GTR_FREQS)
    rate = phy.rate_models.GammaRateModel(NCAT, ALPHA)
    rows = np.asarray(trav.postorder_traversal, dtype=np.int64)
    tips = {trav.names[n]: np.ascontiguousarray(lut[codes[i]]) for i, n in enumerate(names)}
    ot = oracle.OracleTree(2 * n_taxa - 2, tips, NCAT)         # TreeModel.initialise (not timed)
    a, b = trav.root_edge

    def one_eval():
        # TreeModel.compute_partials + compute_likelihood_at_edge, including the 2(N-2) model.p calls
        pm = np.empty((len(rows), 2, NCAT, 4, 4))
        for i, (par, c1, c2) in enumerate(rows):
            pm[i, 0] = model.p(trav.brlens[(int(par), int(c1))], rate.rates)
            pm[i, 1] = model.p(trav.brlens[(int(par), int(c2))], rate.rates)
        ot.compute_partials(rows, pm)
        length = trav.brlens[(a, b)]
        root_pm = np.stack([model.p(0, rate.rates), model.p(length, rate.rates)])
        return float(ot.likelihood_at_edge(a, b, root_pm, model.freqs, rate.weights).sum())

    return one_eval, oracle.max_threads()


def run_cpu_baseline(args, budget_s=12.0):
    one_eval, threads = cpu_eval_factory(args.taxa, args.cpu_patterns, args.seed)
    one_eval()                                   # warm-up (page faults, thread pool)
    t0 = time.perf_counter()
    one_eval()
    t1 = time.perf_counter() - t0
    reps = int(min(50, max(1, budget_s / max(t1, 1e-3))))
    t0 = time.perf_counter()
    for _ in range(reps):
        one_eval()
    per_eval = (time.perf_counter() - t0) / reps
    scale = args.cpu_patterns / float(args.patterns)
    return {
        "value": scale / per_eval, "unit": "lnL evals/s", "cores": threads, "kind": "port",
        "sample": "{} taxa x {} of {} patterns, {} evals of {:.3f} s each, extrapolated linearly in patterns "
                  "(patterns are independent); oracle/pruning_oracle.c, OpenMP".format(
                      args.taxa, args.cpu_patterns, args.patterns, reps, per_eval),
        "site_node_updates_per_s": (args.taxa - 2) * args.cpu_patterns / per_eval,
    }


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    one_eval, threads = cpu_eval_factory(args.taxa, args.cpu_patterns, args.seed)
    for _ in range(max(args.warmup, 1)):
        one_eval()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        lnl = one_eval()
    elapsed = time.perf_counter() - t0
    per_eval = elapsed / args.steps
    scale = args.cpu_patterns / float(args.patterns)
    value = scale / per_eval
    sample = ("each step = one evaluation of {} taxa x {} patterns (of {}), value extrapolated linearly in "
              "patterns; oracle port of the reference engine (the reference is Python+numba and cannot travel), "
              "OpenMP over {} threads").format(args.taxa, args.cpu_patterns, args.patterns, threads)
    line = {
        "impl": "reference", "metric": "lnL evals/s (GTR+G4, {} taxa x {} patterns per GPU)".format(args.taxa, args.patterns),
        "value": value, "unit": "lnL evals/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": per_eval * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, max(1, args.gpus)),   # the other arm's config, key for key
        "cpu_baseline": {"value": value, "unit": "lnL evals/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "lnL evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "sample_lnl": lnl,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, world):
    return {
        "workload": "GTR+G4 nucleotide full-tree lnL, {} taxa x {} site patterns per GPU (BASELINE configs[1])".format(
            args.taxa, args.patterns),
        "taxa": args.taxa, "patterns_per_gpu": args.patterns, "global_patterns": args.patterns * world,
        "categories": NCAT, "states": 4, "tree": "random joins, seed {}".format(args.seed),
        "parallelism": "pattern shards x{}".format(world),
        "l2": "inputs (>=128 MB per node block, 128 GB per evaluation) far exceed the 126 MB L2; no flush needed",
    }


# --------------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist
    import phylo_utils_b200 as phy
    from phylo_utils_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: phylo_utils_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    n_taxa, n_pat = args.taxa, args.patterns
    tree, names = make_tree(n_taxa, args.seed)
    clone = phy.utils.deepcopy_tree(tree)
    trav = phy.traversal.Traversal(clone)
    model = phy.substitution_models.GTR(GTR_RATES, GTR_FREQS)
    rate = phy.rate_models.GammaRateModel(NCAT, ALPHA)
    lut = dna_lut()
    codes_host = torch.from_numpy(make_codes(n_taxa, n_pat, args.seed, rank)).pin_memory()
    codes_np = codes_host.numpy()
    # the e2e path ships the tip codes two per byte (4-bit state-set codes; packed once when the alignment is loaded)
    packed_np = torch.from_numpy(phy.LikelihoodEngine.pack_codes(codes_np)).pin_memory().numpy()
    tip_nodes = np.asarray([trav.names[n] for n in names], dtype=np.int32)

    args.lnl_only = not args.store_partials
    mode = {"auto": _lib.PHB_MODE_RESIDENT if n_pat >= 16384 else _lib.PHB_MODE_LEVEL, "tile": _lib.PHB_MODE_TILE,
            "level": _lib.PHB_MODE_LEVEL, "resident": _lib.PHB_MODE_RESIDENT}[args.mode]
    eng = phy.LikelihoodEngine(n_taxa, n_pat, NCAT, 4, device=local_rank, store_partials=not args.lnl_only)
    if mode == _lib.PHB_MODE_LEVEL:
        rows, offsets = trav.level_order()
        eng.set_schedule(rows, offsets)
    else:
        rows = trav.locality_order()
        eng.set_schedule(rows)
    e = model.eigen
    eng.set_model(e.evecs, e.evals, np.ascontiguousarray(e.ivecs), model.freqs, rate.rates, rate.weights)
    lengths = np.asarray([[trav.brlens[(int(p), int(c1))], trav.brlens[(int(p), int(c2))]] for p, c1, c2 in rows])
    a, b = trav.root_edge
    root_len = trav.brlens[(a, b)]

    codes_dev = codes_host.to(dev, non_blocking=False)
    eng.set_tips(codes_dev, lut, tip_nodes)

    def eval_resident():
        eng.set_edge_lengths(lengths)
        if args.lnl_only:
            return eng.lnl_resident(a, b, root_len)[0]
        eng.build_pmatrices()
        eng.compute_partials(mode)
        return eng.root_lnl(a, b, root_len)[0]

    def eval_e2e(packed=True):
        eng.set_edge_lengths(lengths)
        if args.lnl_only:
            # pinned host codes -> device in chunks, overlapped with the pruning of the previous chunk
            if packed:
                return eng.lnl_from_host(packed_np, a, b, root_len, n_chunks=args.chunks, packed=True)[0]
            return eng.lnl_from_host(codes_np, a, b, root_len, n_chunks=args.chunks)[0]
        eng.set_tips(codes_np, lut, tip_nodes)          # pinned host -> device, N x S bytes
        eng.build_pmatrices()
        eng.compute_partials(mode)
        return eng.root_lnl(a, b, root_len)[0]

    def allreduce(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t)
        return float(t.item())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps):
        barrier()
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        out = None
        for _ in range(steps):
            out = allreduce(fn())
        stop.record()
        barrier()
        ms = start.elapsed_time(stop)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, out

    for _ in range(max(args.warmup, 3)):
        lnl = allreduce(eval_resident())
    launches0 = eng.launch_count
    with ClockSampler(local_rank) as clocks:
        ms, lnl = timed(eval_resident, args.steps)
    launches = eng.launch_count - launches0
    ms_per_step = ms / args.steps
    value = world * 1e3 / ms_per_step          # every rank evaluates its own block of `patterns` patterns per step

    # dominant kernel alone (pruning), CUDA events on the launching stream
    eng.set_edge_lengths(lengths)
    if not args.lnl_only:
        eng.build_pmatrices()
    torch.cuda.synchronize(dev)
    k_start, k_stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k_start.record()
    for _ in range(args.steps):
        if args.lnl_only:
            eng.lnl_resident(a, b, root_len)        # P build (tiny) + the resident kernel + reduction
        else:
            eng.compute_partials(mode)
    k_stop.record()
    torch.cuda.synchronize(dev)
    kernel_ms = k_start.elapsed_time(k_stop) / args.steps
    prune_bytes, eval_bytes = algorithmic_bytes(n_taxa, n_pat)
    peak, peak_src = measured_peak()
    achieved = prune_bytes / (kernel_ms * 1e-3) / 1e9
    if args.lnl_only:
        kernel_name = "dna_pair_kernel<K=4,NC=8> (two patterns per lane, operands on chip, no partials stored)"
    else:
        kernel_name = {_lib.PHB_MODE_TILE: "dna_prune_kernel<K=4> (tile mode)",
                       _lib.PHB_MODE_LEVEL: "dna_prune_kernel<K=4> (level mode)",
                       _lib.PHB_MODE_RESIDENT: "dna_resident_kernel<K=4,STORE=1> (operands on chip, blocks streamed out)"}[mode]
    workload_key = "{}_{}x{}".format("lnl_only" if args.lnl_only else args.mode, n_taxa, n_pat)
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu_traffic(workload_key), "kernel": kernel_name,
                "kernel_ms": kernel_ms, "algorithmic_bytes_per_launch": prune_bytes, "peak_source": peak_src,
                "share_of_step": kernel_ms / ms_per_step}
    if args.lnl_only or mode == _lib.PHB_MODE_RESIDENT:
        # 64 fp64 FMA/clk/SM on B200: (N-2) * S * K * (2*16 + 4) FMA-pipe instructions-worth of work
        fma_ops = (n_taxa - 2) * n_pat * NCAT * 36.0
        roofline["fp64_pipe_frac_at_max_clock"] = fma_ops / (kernel_ms * 1e-3) / (64.0 * 148 * 1.965e9)
        roofline["note"] = ("operand-resident kernel: intermediate partials stay in registers / shared memory / L2, so the "
                            "kernel moves far fewer DRAM bytes than the algorithmic count (see `traffic` and "
                            "profiles/); frac > 1 is expected here (SURVEY.md 8(d): 'kernels that fuse levels on-chip "
                            "may legitimately exceed 100 %'); what binds it is the shared-memory data pipe (78 % busy, profiles/r01l_pair_v2_1000x1M.txt), then the FP64 pipe")

    # the same evaluation while ALSO keeping every node's partials in HBM (what TreeModel.partials and the
    # derivative path need): reported next to the headline so that both costs are on record
    stored = None
    if args.lnl_only and world == 1 and not args.no_stored:
        eng2 = phy.LikelihoodEngine(n_taxa, n_pat, NCAT, 4, device=local_rank)
        eng2.set_schedule(rows)
        eng2.set_model(e.evecs, e.evals, np.ascontiguousarray(e.ivecs), model.freqs, rate.rates, rate.weights)
        eng2.set_tips(codes_dev, lut, tip_nodes)
        stored = {}
        for label, m2 in (("resident_store", _lib.PHB_MODE_RESIDENT), ("tile", _lib.PHB_MODE_TILE)):
            def eval_stored():
                eng2.set_edge_lengths(lengths)
                eng2.build_pmatrices()
                eng2.compute_partials(m2)
                return eng2.root_lnl(a, b, root_len)[0]
            for _ in range(2):
                eval_stored()
            s_ms, s_lnl = timed(eval_stored, max(2, args.steps // 2))
            s_ms /= max(2, args.steps // 2)
            stored[label] = {"evals_per_s": 1e3 / s_ms, "ms_per_step": s_ms, "lnl": s_lnl,
                             "algorithmic_gbs": eval_bytes / (s_ms * 1e-3) / 1e9,
                             "frac_of_hbm_peak": eval_bytes / (s_ms * 1e-3) / 1e9 / peak}
        eng2.close()
        del eng2

    e2e = None
    if not args.no_e2e:
        for _ in range(2):
            allreduce(eval_e2e())
        e_ms, e_lnl = timed(eval_e2e, args.steps)
        e_ms /= args.steps
        code_bytes = packed_np.nbytes if args.lnl_only else n_taxa * n_pat
        e2e = {"value": world * 1e3 / e_ms, "unit": "lnL evals/s", "ms_per_step": e_ms,
               "h2d_bytes_per_step": int(code_bytes + lengths.nbytes + 16), "d2h_bytes_per_step": 8,
               "api": ("LikelihoodEngine.set_edge_lengths + lnl_from_host(packed=True) (C ABI phb_lnl_from_host_packed: "
                       "pinned host tip codes, two 4-bit codes per byte, {} chunks, copy overlapped with compute)".format(
                           args.chunks) if args.lnl_only else
                       "LikelihoodEngine.set_tips(host codes) + set_edge_lengths + build_pmatrices + "
                       "compute_partials + root_lnl (C ABI, pinned host buffers)"), "lnl": e_lnl}
        if args.lnl_only:
            # the same call with one byte per code (phb_lnl_from_host), for the record
            for _ in range(2):
                allreduce(eval_e2e(False))
            u_ms, u_lnl = timed(lambda: eval_e2e(False), args.steps)
            u_ms /= args.steps
            e2e["one_byte_codes"] = {"value": world * 1e3 / u_ms, "ms_per_step": u_ms,
                                     "h2d_bytes_per_step": int(n_taxa * n_pat + lengths.nbytes + 16), "lnl": u_lnl}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = run_cpu_baseline(args)

    if rank == 0:
        line = {
            "metric": "lnL evals/s (GTR+G4, {} taxa x {} patterns per GPU)".format(n_taxa, n_pat),
            "value": value, "unit": "lnL evals/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(args, world),
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
            "path": "lnL only (no per-node partials stored)" if args.lnl_only else "partials stored ({})".format(args.mode),
            "with_stored_partials": stored,
            "clocks": clocks.summary(), "lnl": lnl,
            "site_node_updates_per_s": world * (n_taxa - 2) * n_pat * 1e3 / ms_per_step,
            "eval_algorithmic_gbytes": eval_bytes / 1e9,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
