#!/usr/bin/env python
"""
Headline benchmark (BASELINE.json configs[1]): full-tree lnL evaluations per second, GTR+Gamma4 nucleotide
model, ONE alignment of 1,000 taxa x 1,000,000 site patterns, at 1/2/4/8 B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one evaluation of the whole alignment through the product's own multi-GPU object,
``phylo_utils_b200.parallel.ShardedTreeModel`` (the reference-shaped TreeModel API on every rank):

    tm.compute_partials()   new branch lengths -> device
    tm.lnl()                all transition matrices (one kernel) -> post-order pruning over all N-2 internal nodes ->
                            root combine, Gamma mixture, log, weighted reduction -> device all-reduce -> scalar on the host

  scaling   STRONG: the same 1M-pattern alignment is split into N contiguous pattern shards (rank r owns
            shard_bounds(1M, r, N)); the columns are generated in fixed blocks so the alignment - and hence lnL - is
            the same for every N.  `weak` (1M patterns PER GPU, what round 1 reported) is kept as an extra record.
  value     evaluations/s with the tip codes resident in HBM (CUDA events on the launching stream, max over ranks)
  e2e       the same evaluation starting from pinned HOST codes every step (every step copies one alignment to the device
            and reads one result back): ShardedTreeModel.lnl_from_host_submit / PendingLnl.result - C ABI
            phb_lnl_from_host_submit: packed tip codes, the copy engine feeds the kernel chunk by chunk, and two
            evaluations are in flight so that the copy of alignment i+1 runs under the walk of alignment i;
            `e2e.one_at_a_time` is the same call sequence with the result on the host before the next evaluation is issued
  roofline  the dominant kernel timed alone.  The lnL-only walk keeps every intermediate on chip: it is bound by the
            fp64 / shared-memory pipes, NOT by HBM, and is reported as such (bound "fp64", denominator = the DFMA
            rate measured in this run); its algorithmic-byte rate is given for the record and labelled as what it
            is.  The walks that keep all partials in HBM carry the HBM rooflines (`with_stored_partials`).
  configs   bounded sub-records for BASELINE configs 1, 3, 4, 5 (N = 1; cfg5 as named at N = 8)
  cpu_baseline / --impl reference   the reference's CPU algorithm (oracle port, OpenMP on all host cores) on a
            bounded pattern sample of the same workload; only the pattern-proportional part is extrapolated
"""
import argparse
import gc
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
CODE_BLOCK = 125000          # the alignment's columns are generated in blocks of this many patterns


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--taxa", type=int, default=1000)
    ap.add_argument("--patterns", type=int, default=1000000, help="site patterns of the WHOLE alignment (split over the GPUs)")
    ap.add_argument("--cpu-patterns", type=int, default=20000, help="pattern sample for the CPU baseline")
    ap.add_argument("--seed", type=int, default=2)
    ap.add_argument("--chunks", type=int, default=0, help="host->device pipeline depth of the e2e path (0: the library's rule, ~4 MB per chunk)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-stored", action="store_true", help="skip the 'partials stored' measurements (HBM rooflines)")
    ap.add_argument("--no-weak", action="store_true", help="skip the weak-scaling record (N > 1)")
    ap.add_argument("--no-configs", action="store_true", help="skip the cfg1/3/4/5 sub-records")
    return ap.parse_args()


# --------------------------------------------------------------------------------------------------
# synthetic workload: seeded random-join topology, branch lengths U(0.01, 0.3), iid uniform
# states with 1 % fully ambiguous codes; at these sizes every column is a distinct pattern.
# Pure numpy: the reference arm builds the SAME problem without importing phylo_utils_b200.
# --------------------------------------------------------------------------------------------------
def make_codes(n_taxa, lo, hi, seed):
    """Columns [lo, hi) of the global alignment; block b of CODE_BLOCK columns comes from rng([seed, b])."""
    out = np.empty((n_taxa, hi - lo), dtype=np.uint8)
    for b in range(lo // CODE_BLOCK, (hi + CODE_BLOCK - 1) // CODE_BLOCK):
        rng = np.random.default_rng([seed, b])
        blk = rng.integers(0, 4, size=(n_taxa, CODE_BLOCK), dtype=np.uint8)
        blk[rng.random((n_taxa, CODE_BLOCK), dtype=np.float32) < 0.01] = 4
        b0 = b * CODE_BLOCK
        s, e = max(lo, b0), min(hi, b0 + CODE_BLOCK)
        out[:, s - lo:e - lo] = blk[:, s - b0:e - b0]
    return out


def dna_lut():
    # rows in lexicographic rank order of the 0/1 vectors: T, G, C, A, then the all-ones gap row
    return np.vstack([np.eye(4)[::-1], np.ones((1, 4))])


def workload_tree(n_taxa, seed):
    """
    The bench tree as plain tables - the same random draws, in the same order, as phylo_utils_b200.tree.random_tree
    (join two uniformly chosen live subtrees until one is left; lengths U(0.01, 0.3) in pre-order), so that both
    bench arms evaluate one and the same tree.  Returns (rows, lengths, root_edge, root_length, leaf_order):
    rows (N-2, 3) = PAR, CH1, CH2 in post-order without the root, lengths (N-2, 2), the root edge (a, b) with the
    sum of the two root branches (what deroot() leaves, reference utils.py:114-118), and the tip node ids in the
    pre-order in which the alignment rows are assigned.  Tip i of the join order is node i.
    """
    rng = np.random.default_rng(seed)
    live = list(range(n_taxa))
    children, nxt = {}, n_taxa
    while len(live) > 1:
        i, j = rng.choice(len(live), size=2, replace=False)
        children[nxt] = (live[i], live[j])
        for k in sorted((int(i), int(j)), reverse=True):
            live.pop(k)
        live.append(nxt)
        nxt += 1
    root = live[0]
    length, leaves, stack = {}, [], [root]
    while stack:                                     # pre-order, first child first
        nd = stack.pop()
        if nd != root:
            length[nd] = float(rng.uniform(0.01, 0.3))
        if nd in children:
            stack.extend(reversed(children[nd]))
        else:
            leaves.append(nd)
    rows, lens, stack = [], [], [(root, 0)]
    while stack:                                     # post-order
        nd, i = stack.pop()
        kids = children.get(nd, ())
        if i < len(kids):
            stack.append((nd, i + 1))
            stack.append((kids[i], 0))
        elif kids and nd != root:
            rows.append((nd, kids[0], kids[1]))
            lens.append((length[kids[0]], length[kids[1]]))
    a, b = children[root]
    return np.asarray(rows, dtype=np.int64), np.asarray(lens), (a, b), length[a] + length[b], leaves


def gtr_eigen(rates6, freqs):
    """Reversible Q scaled to one substitution per site and its symmetric eigen-decomposition (textbook; the reference
    does the same in substitution_models/utils.py:45-98)."""
    pi = np.asarray(freqs, dtype=np.double)
    r = np.zeros((4, 4))
    r[np.triu_indices(4, 1)] = rates6
    r = r + r.T
    q = r * pi[None, :]
    np.fill_diagonal(q, 0.0)
    np.fill_diagonal(q, -q.sum(1))
    q /= -(pi * np.diag(q)).sum()
    sq = np.sqrt(pi)
    lam, u = np.linalg.eigh(q * sq[:, None] / sq[None, :])
    return u / sq[:, None], lam, u.T * sq[None, :]


def algorithmic_bytes(n_taxa, n_pat, K=NCAT, A=4):
    """S * [(N-2) * (2*K*A*8 + 16) + N + 16]; the pruning kernel's share excludes the root's 16."""
    b_node = 2 * K * A * 8 + 16
    prune = n_pat * ((n_taxa - 2) * b_node + n_taxa)
    return prune, prune + n_pat * 16


def algorithmic_flops(n_taxa, n_pat, K=NCAT, A=4):
    """F_node = K * (2 * 2 A^2 + A) fp64 flop per site-node update (272 at A = 4, K = 4)."""
    return float(n_taxa - 2) * n_pat * K * (4 * A * A + A)


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


def bind_to_gpu_numa_node(index):
    """Pin this rank's threads to the CPUs of the NUMA node its GPU hangs off, BEFORE the pinned host buffers are
    allocated (first touch places them on that node): the e2e path streams tip codes from host memory on every step,
    and at 4 - 8 GPUs the box's aggregate host-to-device bandwidth is what bounds it.  Returns what was found / done."""
    info = {"gpu": index}
    try:
        import torch
        bus = torch.cuda.get_device_properties(index).pci_bus_id
        dom = torch.cuda.get_device_properties(index).pci_domain_id
        dev = torch.cuda.get_device_properties(index).pci_device_id
        path = "/sys/bus/pci/devices/{:04x}:{:02x}:{:02x}.0/numa_node".format(dom, bus, dev)
        node = int(open(path).read().strip())
        info["numa_node"] = node
        nodes = [d for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit()]
        info["numa_nodes_on_host"] = len(nodes)
        if node >= 0 and len(nodes) > 1:
            cpus = set()
            for part in open("/sys/devices/system/node/node{}/cpulist".format(node)).read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
            cpus &= os.sched_getaffinity(0)
            if cpus:
                os.sched_setaffinity(0, cpus)
                info["bound_to_cpus"] = len(cpus)
        else:
            info["bound_to_cpus"] = None      # one node (or unknown): nothing to choose
    except Exception as exc:                   # sysfs layout differs, containers hide it, ...: measurement goes on unbound
        info["error"] = repr(exc)[:120]
    info["cpus_available"] = len(os.sched_getaffinity(0))
    return info


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_record(key):
    """Per-launch figures of a kernel from the committed `ncu --set full` summaries (profiles/roofline_traffic.json)."""
    path = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(path):
        with open(path) as fh:
            return json.load(fh).get(key)
    return None


# --------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference engine.  Nothing in here imports phylo_utils_b200.
# --------------------------------------------------------------------------------------------------
def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_eval_factory(n_taxa, n_pat, seed):
    """
    One reference-style evaluation on a sample of the workload's patterns, split into the two parts that scale
    differently: `p_matrices()` = the 2(N-2) + 2 Model.p calls of TreeModel.compute_partials (tree_model.py:166-169;
    independent of the pattern count) and `prune(pm)` = the clv loop + compute_likelihood_at_edge (proportional to it).
    """
    from oracle import oracle
    rows, lengths, (a, b), root_len, leaves = workload_tree(n_taxa, seed)
    codes = make_codes(n_taxa, 0, n_pat, seed)
    lut = dna_lut()
    evecs, evals, ivecs = gtr_eigen(GTR_RATES, GTR_FREQS)
    rates = oracle.ref_discrete_gamma(ALPHA, NCAT) if oracle.have_ref_gamma() else oracle.discrete_gamma_scipy(ALPHA, NCAT)
    weights = np.full(NCAT, 1.0 / NCAT)
    freqs = np.asarray(GTR_FREQS)
    threads = host_threads()        # explicit: torchrun exports OMP_NUM_THREADS=1, which would serialise the arm

    def model_p(t):                 # Model.p(t, rates): abstract.py:49-59
        return np.stack([(evecs * np.exp(evals * (t * r))).dot(ivecs) for r in rates])

    tips = {node: np.ascontiguousarray(lut[codes[i]]) for i, node in enumerate(leaves)}
    ot = oracle.OracleTree(2 * n_taxa - 1, tips, NCAT, n_threads=threads)   # TreeModel.initialise (not timed)

    def p_matrices():
        pm = np.empty((len(rows), 2, NCAT, 4, 4))
        for i in range(len(rows)):
            pm[i, 0] = model_p(lengths[i, 0])
            pm[i, 1] = model_p(lengths[i, 1])
        return pm, np.stack([model_p(0.0), model_p(root_len)])

    def prune(pms):
        pm, root_pm = pms
        ot.compute_partials(rows, pm)
        return float(ot.likelihood_at_edge(a, b, root_pm, freqs, weights).sum())

    return p_matrices, prune, threads


def time_cpu(args, steps, warmup):
    p_matrices, prune, threads = cpu_eval_factory(args.taxa, args.cpu_patterns, args.seed)
    for _ in range(max(warmup, 1)):
        pms = p_matrices()
        prune(pms)
    t0 = time.perf_counter()
    for _ in range(steps):
        pms = p_matrices()
    t_p = (time.perf_counter() - t0) / steps
    t0 = time.perf_counter()
    for _ in range(steps):
        lnl = prune(pms)
    t_prune = (time.perf_counter() - t0) / steps
    # only the pattern-proportional part is extrapolated; the P-matrix build does not depend on the pattern count
    per_eval = t_p + t_prune * (args.patterns / float(args.cpu_patterns))
    sample = ("{} taxa x {} of {} patterns (the first columns of the same alignment), {} timed evaluations: P-matrix build "
              "{:.4f} s (counted once) + pruning {:.4f} s (x {:.0f}, patterns are independent); oracle/pruning_oracle.c "
              "(C restatement of the reference's numba engine), OpenMP over {} threads").format(
                  args.taxa, args.cpu_patterns, args.patterns, steps, t_p, t_prune, args.patterns / float(args.cpu_patterns), threads)
    return {"value": 1.0 / per_eval, "unit": "lnL evals/s", "cores": threads, "kind": "port", "sample": sample,
            "sample_lnl": lnl, "sample_seconds": t_p + t_prune,
            "site_node_updates_per_s": (args.taxa - 2) * args.cpu_patterns / t_prune}


def run_cpu_baseline(args, budget_s=12.0):
    p_matrices, prune, _ = cpu_eval_factory(args.taxa, min(args.cpu_patterns, 2000), args.seed)
    t0 = time.perf_counter()
    prune(p_matrices())
    probe = (time.perf_counter() - t0) * args.cpu_patterns / float(min(args.cpu_patterns, 2000))
    return time_cpu(args, steps=int(min(20, max(2, budget_s / max(probe, 1e-3)))), warmup=1)


def run_reference_arm(args):
    """The reference's CPU algorithm (oracle port; the reference itself is Python + numba + dendropy/Biopython and its
    sources cannot travel to the GPU box) on the box's host cores.  Rank 0 alone works; other ranks exit."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    cpu = time_cpu(args, steps=args.steps, warmup=args.warmup)
    line = {
        "impl": "reference", "metric": metric_name(args), "value": cpu["value"], "unit": "lnL evals/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / cpu["value"],
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, max(1, args.gpus)),   # the other arm's config, key for key
        "cpu_baseline": cpu,
        "e2e": {"value": cpu["value"], "unit": "lnL evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "sample_lnl": cpu["sample_lnl"],
    }
    print(json.dumps(line), flush=True)


def metric_name(args):
    return "lnL evals/s (GTR+G4, {} taxa x {} site patterns)".format(args.taxa, args.patterns)


def workload_config(args, world):
    return {
        "workload": "GTR+G4 nucleotide full-tree lnL, ONE alignment of {} taxa x {} site patterns (BASELINE configs[1]), "
                    "split into {} pattern shard(s)".format(args.taxa, args.patterns, world),
        "taxa": args.taxa, "global_patterns": args.patterns, "patterns_per_gpu": args.patterns // world,
        "categories": NCAT, "states": 4, "tree": "random joins, seed {}".format(args.seed),
        "parallelism": "pattern shards x{} (ShardedTreeModel, the scalar summed over ranks on the device)".format(world),
        "l2": "every evaluation re-reads its shard's tip codes ({} MB per GPU) and, in the stored-partials walks, streams "
              ">= 16 GB of node blocks: far beyond the 126 MB L2, no flush needed".format(args.taxa * (args.patterns // world) // 1000000),
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
    from phylo_utils_b200.engine import fp64_peak
    from phylo_utils_b200.parallel import ShardedTreeModel, shard_bounds

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: phylo_utils_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = bind_to_gpu_numa_node(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    n_taxa, n_pat = args.taxa, args.patterns
    tree = phy.tree.random_tree(n_taxa, args.seed)
    names = {lf.taxon.label: i for i, lf in enumerate(tree.leaf_node_iter())}
    model = phy.substitution_models.GTR(GTR_RATES, GTR_FREQS)
    rate = phy.rate_models.GammaRateModel(NCAT, ALPHA)
    lut = dna_lut()
    lo, hi = shard_bounds(n_pat, rank, world)
    n_local = hi - lo

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(fn, steps):
        """K calls of fn between barriers, CUDA events on the stream the engine launches on; ms = max over ranks."""
        barrier()
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        out = None
        for _ in range(steps):
            out = fn()
        stop.record()
        barrier()
        return max_over_ranks(start.elapsed_time(stop)), out

    def sharded_model(codes, n_global, **kw):
        tm = ShardedTreeModel(device=local_rank, **kw)
        tm.set_tree(tree)
        tm.set_local_tip_codes(codes, lut, names, n_global)
        tm.set_rate_model(rate)
        tm.set_substitution_model(model)
        tm.initialise()
        return tm

    # ---- the headline: strong scaling of the one alignment, lnL-only walk ---------------------------------------
    codes_host = torch.from_numpy(make_codes(n_taxa, lo, hi, args.seed)).pin_memory()
    tm = sharded_model(codes_host.numpy(), n_pat, store_partials=False)
    eng = tm.local.engine
    peer_sums = bool(getattr(tm, "peer_sums", False))
    root_a, root_b = tm.traversal.root_edge

    def evaluate():
        tm.compute_partials()          # branch lengths -> device (an lnL-only model stores no partials)
        return tm.lnl()                # P build + walk + reduction + device all-reduce + 8 bytes to the host

    warmup = max(args.warmup, 3)
    for _ in range(warmup):
        lnl = evaluate()
    launches0, coll0 = eng.launch_count, tm.collectives
    with ClockSampler(local_rank) as clocks:
        ms, lnl = timed(evaluate, args.steps)
    launches = eng.launch_count - launches0
    collectives = tm.collectives - coll0
    ms_per_step = ms / args.steps
    value = 1e3 / ms_per_step

    # ---- dominant kernel alone (P build + walk + block-sum reduction enqueued back to back, no host round trips) --
    root_len = tm.traversal.brlens[(root_a, root_b)]
    torch.cuda.synchronize(dev)
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0.record()
    for _ in range(args.steps):
        eng.lnl_resident_async(root_a, root_b, root_len)
    k1.record()
    torch.cuda.synchronize(dev)
    kernel_ms = max_over_ranks(k0.elapsed_time(k1) / args.steps)
    peak_hbm, peak_src = measured_peak()
    dfma_peak = fp64_peak(local_rank)
    prune_bytes, _ = algorithmic_bytes(n_taxa, n_local)
    flops = algorithmic_flops(n_taxa, n_local)
    # the launch's own rules (clv_dna_pair.cu): 128-pattern tiles when the 64-pattern tiles need a second wave (pair_ppt);
    # one CTA of all resident warps per SM from six rounds of tiles on, independent one-warp CTAs below that (launch_pair)
    sms = torch.cuda.get_device_properties(local_rank).multi_processor_count
    ppt = 4 if ((n_local + 63) // 64 > 12 * sms and (n_local + 127) // 128 <= 8 * sms) else 2
    warps_per_sm = 12 if ppt == 2 else 8
    cta_form = 10 * ((n_local + 32 * ppt - 1) // (32 * ppt)) // (sms * warps_per_sm) >= 60
    kernel_name = "dna_pair_cta_kernel" if cta_form else "dna_pair_kernel"
    ncu = ncu_record("dna_pair_cta_kernel_lnl_only_1000x1M" if cta_form else "dna_pair_kernel_lnl_only_1000x1M") or {}
    achieved_tf = flops / (kernel_ms * 1e-3) / 1e12
    roofline = {
        "bound": "fp64", "achieved": achieved_tf, "peak": dfma_peak, "unit": "TFLOP/s", "frac": achieved_tf / dfma_peak,
        "traffic": ncu.get("dram_bytes_per_launch"),
        "kernel": "{}<K=4,NC=8,PPT={},SYM> ({} patterns per lane, {}; whole post-order walk + root + reduction in one "
                  "launch, operands on chip, symmetric P blocks, no partials stored)".format(
                      kernel_name, ppt, {2: "two", 4: "four"}[ppt],
                      "one CTA of {} warps per SM".format(warps_per_sm) if cta_form else "one-warp CTAs, {} per SM".format(warps_per_sm)),
        "kernel_ms": kernel_ms, "patterns_per_launch": n_local, "share_of_step": kernel_ms / ms_per_step,
        "algorithmic_flops_per_launch": flops,
        "flops_per_unit": "272 fp64 flop per site-node update (SURVEY.md 8(d): K (4 A^2 + A)), x (N-2) x patterns of the shard",
        "peak_source": "measured in this run: phb_op_fp64_peak (8 independent DFMA chains per thread on every SM); "
                       "nominal 64 FMA/clk/SM x 148 x 1.965 GHz = 37.2",
        "limiter": "per-warp latency at three warps per scheduler (168 registers): 20 % fewer shared-memory wavefronts changed "
                   "nothing, 6 % fewer instructions bought 1.3 % (DESIGN.md 3.1); the busiest units are the shared-memory (LSU) data "
                   "pipe and the fp64 pipe; DRAM is idle (see `ncu`)",
        "ncu": ncu or None,
        "hbm_algorithmic": {
            "bytes_per_launch": prune_bytes, "gbs": prune_bytes / (kernel_ms * 1e-3) / 1e9, "hbm_peak_gbs": peak_hbm,
            "ratio_to_hbm_peak": prune_bytes / (kernel_ms * 1e-3) / 1e9 / peak_hbm, "peak_source": peak_src,
            "note": "NOT a roofline fraction: SURVEY.md 8(d)'s algorithmic bytes (272 B per site-node update) assume every "
                    "node block is written to and read from HBM; this kernel keeps them in registers / shared memory / L2 "
                    "and moves ~1.5 % of that through DRAM (`traffic`).  The HBM rooflines are in `with_stored_partials`."},
    }

    # ---- e2e: the same evaluation from pinned HOST codes every step (TreeModel-level API) ---------------------------
    e2e = None
    if not args.no_e2e:
        # the alphabet has 5 code rows (<= 8): 3 bits per code in two planes (phb_split_codes, done once when an alignment
        # is loaded); the two-codes-per-byte format of round 1 is measured next to it
        planes = tuple(torch.from_numpy(x).pin_memory().numpy() for x in phy.LikelihoodEngine.split_codes(codes_host.numpy()))
        packed = torch.from_numpy(phy.LikelihoodEngine.pack_codes(codes_host.numpy())).pin_memory().numpy()

        def evaluate_e2e(host_codes):
            tm.compute_partials()
            return tm.lnl_from_host_codes(host_codes, n_chunks=args.chunks)

        def measure_e2e(host_codes, code_bytes, what):
            for _ in range(2):
                evaluate_e2e(host_codes)
            e_ms, e_lnl = timed(lambda: evaluate_e2e(host_codes), args.steps)
            e_ms /= args.steps
            # the same calls with TWO evaluations in flight (lnl_from_host_submit): every step still copies one alignment
            # to the device and reads one result back, but the copy of alignment i+1 runs under the walk of alignment i
            state = {"pending": None}

            def pipelined_step():
                tm.compute_partials()
                nxt = tm.lnl_from_host_submit(host_codes, n_chunks=args.chunks)
                out = state["pending"].result() if state["pending"] is not None else None
                state["pending"] = nxt
                return out
            for _ in range(3):
                pipelined_step()
            p_ms, p_lnl = timed(pipelined_step, args.steps)
            p_ms /= args.steps
            state["pending"].result()
            return {"value": 1e3 / p_ms, "unit": "lnL evals/s", "ms_per_step": p_ms,
                    "h2d_bytes_per_step": int(code_bytes + (2 * (n_taxa - 2)) * 8 + 16), "d2h_bytes_per_step": 8,
                    "host_to_device_gbs_per_gpu": code_bytes / (p_ms * 1e-3) / 1e9, "tip_code_format": what, "lnl": p_lnl,
                    "in_flight": 2,
                    "one_at_a_time": {"value": 1e3 / e_ms, "ms_per_step": e_ms, "lnl": e_lnl,
                                      "note": "lnl_from_host_codes: the result of an evaluation is on the host before the next one is issued"}}
        split = measure_e2e(planes, int(planes[0].nbytes + planes[1].nbytes), "3 bits per code: a plane of 2-bit values + a plane of high bits (phb_lnl_from_host_split_async)")
        nibble = measure_e2e(packed, int(packed.nbytes), "two 4-bit codes per byte (phb_lnl_from_host_packed_async)")
        # which format a deployment uses is a property of the box, decided before the run: one or two GPUs decode nibbles
        # faster than the copy engine delivers them (the 3-bit decode costs the kernel 3 %); from four GPUs on the shared
        # host-to-device link is the bound and fewer bytes win
        e2e, other = (split, nibble) if world > 2 else (nibble, split)
        e2e["bytes_are"] = "per GPU (its shard's tip codes + the branch lengths); x{} over the box".format(world)
        e2e["api"] = ("ShardedTreeModel.compute_partials() + lnl_from_host_submit(codes) / PendingLnl.result() -> TreeModel -> C ABI "
                      "phb_set_edge_lengths + phb_lnl_from_host_submit (pinned host tip codes, {} chunks, copied by the copy engine while "
                      "the kernel already walks the first chunks and the previous alignment), sum over ranks on the device, phb_result_post + "
                      "phb_result_wait".format(args.chunks or "~4 MB"))
        e2e["other_format"] = other
        del packed, planes

    # ---- parity inside the run ------------------------------------------------------------------------------------
    # (a) the sharded product path against the reference-generated golden case cfg1 (10 taxa x 1000 patterns, numba
    #     reference output committed under tests/golden/): total lnL over NCCL, per-edge derivative sums over NCCL
    parity = sharded_golden_check(phy, ShardedTreeModel, local_rank)

    # ---- weak scaling, for continuity with round 1 (1M patterns PER GPU) -----------------------------------------------
    weak = None
    if world > 1 and not args.no_weak:
        del tm, eng
        gc.collect()
        w_codes = make_codes(n_taxa, rank * n_pat, (rank + 1) * n_pat, args.seed)
        tmw = sharded_model(w_codes, n_pat * world, store_partials=False)

        def evaluate_weak():
            tmw.compute_partials()
            return tmw.lnl()
        for _ in range(warmup):
            evaluate_weak()
        w_ms, w_lnl = timed(evaluate_weak, args.steps)
        w_ms /= args.steps
        weak = {"scaling": "weak", "patterns_per_gpu": n_pat, "global_patterns": n_pat * world, "ms_per_step": w_ms,
                "value_1M_pattern_evals_per_s": world * 1e3 / w_ms, "lnl": w_lnl}
        del tmw, w_codes
        gc.collect()
    else:
        del tm, eng
        gc.collect()

    # ---- the walks that keep every node block in HBM (TreeModel.partials, the derivative path): HBM rooflines -------------
    stored = None
    if world == 1 and not args.no_stored:
        stored = stored_partials_records(phy, _lib, args, tree, names, model, rate, lut, codes_host, timed, peak_hbm, peak_src, dev)
    del codes_host
    gc.collect()
    torch.cuda.empty_cache()

    # ---- the other BASELINE configs, bounded ------------------------------------------------------------------------------
    configs = None
    if not args.no_configs and world in (1, 8):
        from tools import bench_configs
        try:
            configs = bench_configs.sub_records(world, rank, local_rank, peak_hbm, fp64_peak(local_rank, tensor=True), dfma_peak)
        except Exception as exc:            # a sub-record must never cost the headline line
            configs = {"error": repr(exc)[:300]}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = run_cpu_baseline(args)
        # the GPU path on exactly the CPU sample: per-evaluation parity of the two arms inside the run
        tm_s = phy.TreeModel(device=local_rank, store_partials=False)
        tm_s.set_tree(tree)
        tm_s.set_tip_codes(make_codes(n_taxa, 0, args.cpu_patterns, args.seed), lut, names)
        tm_s.set_rate_model(rate)
        tm_s.set_substitution_model(model)
        tm_s.initialise()
        gpu_sample = tm_s.lnl()
        cpu["gpu_lnl_on_the_same_sample"] = gpu_sample
        cpu["rel_diff"] = abs(gpu_sample - cpu["sample_lnl"]) / abs(cpu["sample_lnl"])
        del tm_s

    if rank == 0:
        line = {
            "metric": metric_name(args), "value": value, "unit": "lnL evals/s", "n_gpus": world, "steps": args.steps,
            "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args, world),
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
            "collectives_in_timed_region": int(collectives),
            "rank_sum": ("inside the reduction kernel: every rank stores its total into the peers' exchange buffers over NVLink "
                         "(CUDA IPC) and adds what arrives in rank order - phb_peer_sum_next, no collective-library call"
                         if peer_sums else
                         ("torch.distributed all_reduce in place on the device result buffer" if world > 1 else "single rank")),
            "path": "ShardedTreeModel(store_partials=False): lnL-only operand-resident walk per shard, sum over ranks on the device",
            "with_stored_partials": stored, "weak_scaling": weak, "configs": configs, "sharded_parity": parity,
            "clocks": clocks.summary(), "lnl": lnl, "host_placement_rank0": numa,
            "site_node_updates_per_s": (n_taxa - 2) * n_pat * 1e3 / ms_per_step,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def sharded_golden_check(phy, ShardedTreeModel, local_rank):
    """cfg1 golden case (output of the unmodified reference, tests/golden/cfg1_gtr_g4.npz) through ShardedTreeModel at
    this run's world size: total lnL, per-site lnL (all-gather) and the pulley property of the all-reduced derivative sums."""
    path = os.path.join(ROOT, "tests", "golden", "cfg1_gtr_g4.npz")
    if not os.path.exists(path):
        return None
    from phylo_utils_b200.alignment.alignment import SeqRecord
    with np.load(path) as z:
        g = {k: z[k] for k in z.files}
    records = [SeqRecord(str(n), bytes(row).decode("ascii")) for n, row in zip(g["names"], g["seqs"])]
    tm = ShardedTreeModel(device=local_rank, up_partials=True)
    tm.set_tree(phy.tree.parse_newick(str(g["newick"])))
    tm.set_alignment(records, int(g["alphabet"]))
    tm.set_rate_model(phy.rate_models.GammaRateModel(NCAT, ALPHA))
    tm.set_substitution_model(phy.substitution_models.GTR(GTR_RATES, GTR_FREQS))
    tm.initialise()
    want = float(g["total_lnl"])
    total = tm.lnl()
    site = tm.compute_likelihood_at_edge(*tm.traversal.root_edge)
    tm.compute_up_partials()
    d = tm.edge_derivatives(phy.optimise.edge_nodes(tm.traversal))
    return {"case": "tests/golden/cfg1_gtr_g4.npz (numba reference output), sharded x{}".format(tm.world),
            "total_rel_err": abs(total - want) / abs(want),
            "site_max_rel_err": float(np.max(np.abs(site - g["site_lnl"]) / np.abs(g["site_lnl"]))),
            "pulley_max_rel_err": float(np.max(np.abs(d[:, 0] - want)) / abs(want)),
            "collectives": tm.collectives,
            "ok": bool(abs(total - want) <= 1e-10 * abs(want) and np.allclose(site, g["site_lnl"], rtol=1e-10, atol=0)
                       and np.max(np.abs(d[:, 0] - want)) <= 1e-10 * abs(want))}


def stored_partials_records(phy, _lib, args, tree, names, model, rate, lut, codes_host, timed, peak_hbm, peak_src, dev):
    """Same evaluation with every node block kept in HBM (133 GB at 1000 x 1M): the operand-resident store walk and the
    streaming tile walk, each with its HBM roofline (algorithmic bytes = SURVEY.md 8(d), all of them really moved)."""
    import torch
    n_taxa, n_pat = args.taxa, args.patterns
    _, eval_bytes = algorithmic_bytes(n_taxa, n_pat)
    out = {}
    for label, mode, key in (("resident_store", "resident", "dna_pair_store_kernel_1000x1M"),
                             ("tile", "tile", "dna_prune_kernel_tile_1000x1M")):
        tm = phy.TreeModel(device=dev.index, mode=mode)
        tm.set_tree(tree)
        tm.set_tip_codes(codes_host.numpy(), lut, names)
        tm.set_rate_model(rate)
        tm.set_substitution_model(model)
        tm.initialise()

        def evaluate():
            tm.compute_partials()
            return tm.lnl()
        evaluate()
        steps = max(2, args.steps // 4)
        s_ms, s_lnl = timed(evaluate, steps)
        s_ms /= steps
        ncu = ncu_record(key) or {}
        gbs = eval_bytes / (s_ms * 1e-3) / 1e9
        out[label] = {"evals_per_s": 1e3 / s_ms, "ms_per_step": s_ms, "lnl": s_lnl,
                      "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak_hbm, "unit": "GB/s", "frac": gbs / peak_hbm,
                                   "traffic": ncu.get("dram_bytes_per_launch"), "algorithmic_bytes_per_step": eval_bytes,
                                   "peak_source": peak_src, "timed": "whole step (P build + walk + root kernel), CUDA events"}}
        del tm
        gc.collect()
        torch.cuda.empty_cache()
    return out


if __name__ == "__main__":
    main()
