"""
Builds phylo_utils_b200/libphylo_b200.so in-tree with nvcc for sm_100a.

    python phylo_utils_b200/csrc/build.py [--force] [--verbose] [--ptxas]

The shared object is git-ignored but travels to the GPU box with the working tree.
"""
import concurrent.futures
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
ROOT = os.path.dirname(PKG)
OUT = os.path.join(PKG, "libphylo_b200.so")
OBJ = os.path.join(HERE, "_obj")

CUDA_SOURCES = ["api.cu", "pmatrix.cu", "clv_dna.cu", "clv_dna_pair.cu", "clv_dna_pair_k4n8.cu", "clv_dna_pair_k4n16.cu", "clv_dna_pair_k12.cu",
                "clv_dna_pair_k8.cu", "up_dna_pair.cu", "clv_generic.cu", "clv_mma.cu", "ops.cu", "derivs.cu", "compress.cu"]
HOST_SOURCES = ["discrete_gamma.cpp"]
HEADERS = ["common.cuh", "resident_plan.cuh", "pair_common.cuh", "pair_walk.cuh", os.path.join(ROOT, "include", "phylo_b200.h")]

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
              "--expt-relaxed-constexpr"]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; cannot build libphylo_b200.so")
    return exe


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _run(cmd, verbose):
    if verbose:
        print(" ".join(cmd), flush=True)
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError("build step failed:\n  {}\n{}".format(" ".join(cmd), res.stdout))
    if verbose and res.stdout.strip():
        print(res.stdout)
    return res.stdout


def build(force=False, verbose=False, ptxas_info=False, checks=False):
    """checks=True: the same sources with -DPHB_DEVICE_CHECKS (device-side invariant asserts, common.cuh) into
    libphylo_b200_checks.so - a developer build the parity tests can be pointed at through PHB_LIBRARY."""
    nvcc = _nvcc()
    global OBJ, OUT
    saved = OBJ, OUT
    if checks:
        OBJ, OUT = os.path.join(HERE, "_obj_checks"), os.path.join(PKG, "libphylo_b200_checks.so")
    try:
        return _build(nvcc, force, verbose, ptxas_info, ["-DPHB_DEVICE_CHECKS"] if checks else [])
    finally:
        OBJ, OUT = saved


def _build(nvcc, force, verbose, ptxas_info, defines):
    os.makedirs(OBJ, exist_ok=True)
    headers = [h if os.path.isabs(h) else os.path.join(HERE, h) for h in HEADERS]
    headers.append(os.path.abspath(__file__))
    jobs = []
    objs = []
    for src in CUDA_SOURCES + HOST_SOURCES:
        s = os.path.join(HERE, src)
        o = os.path.join(OBJ, src + ".o")
        objs.append(o)
        if force or _stale(o, [s] + headers):
            if src.endswith(".cu"):
                cmd = [nvcc] + ARCH + NVCC_FLAGS + defines + (["-Xptxas", "-v"] if ptxas_info else []) + ["-c", s, "-o", o]
            else:
                cmd = [nvcc, "-O2", "-std=c++17", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
                       "-Xcompiler", "-ffp-contract=off", "-c", s, "-o", o]
            jobs.append(cmd)
    logs = []
    if jobs:
        with concurrent.futures.ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 2)) as pool:
            logs = list(pool.map(lambda c: _run(c, verbose), jobs))
    if jobs or force or _stale(OUT, objs):
        _run([nvcc] + ARCH + ["-shared", "-o", OUT] + objs + ["-cudart", "static"], verbose)
    return OUT, logs


if __name__ == "__main__":
    out, logs = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, ptxas_info="--ptxas" in sys.argv,
                      checks="--checks" in sys.argv)
    if "--ptxas" in sys.argv:
        print("\n".join(logs))
    print(out)
