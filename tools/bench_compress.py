"""
Timing of site-pattern compression: GPU (phb_compress_patterns: H2D + encode + radix sort + runs + D2H, wall
clock of the whole call) against the host implementation (numpy argsort on column byte strings).

    python tools/bench_compress.py [--shapes 100x20000,1000x200000,1000x1000000]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from phylo_utils_b200.alignment.alignment import compress_codes, compress_codes_gpu  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shapes", default="100x20000,1000x200000,1000x1000000")
    ap.add_argument("--host-limit", type=int, default=60000000, help="skip the host run above this many cells")
    args = ap.parse_args()
    for shape in args.shapes.split(","):
        ntax, nsite = (int(v) for v in shape.split("x"))
        rng = np.random.default_rng(ntax + nsite)
        base = rng.integers(0, 5, size=(ntax, max(1, nsite // 2)), dtype=np.uint8)
        codes = np.ascontiguousarray(base[:, rng.integers(0, base.shape[1], size=nsite)])   # ~43 % distinct columns
        compress_codes_gpu(codes[:, :1000])                                                  # context creation
        t0 = time.perf_counter()
        gp, gw, ginv = compress_codes_gpu(codes)
        t_gpu = time.perf_counter() - t0
        t0 = time.perf_counter()
        gp, gw, ginv = compress_codes_gpu(codes)
        t_gpu = min(t_gpu, time.perf_counter() - t0)
        line = {"taxa": ntax, "sites": nsite, "patterns": int(gp.shape[1]), "gpu_s": t_gpu,
                "gpu_msites_per_s": nsite / t_gpu / 1e6}
        if ntax * nsite <= args.host_limit:
            t0 = time.perf_counter()
            hp, hw, hinv = compress_codes(codes)
            line["host_s"] = time.perf_counter() - t0
            line["identical"] = bool(np.array_equal(hp, gp) and np.array_equal(hw, gw) and np.array_equal(hinv, ginv))
            line["speedup"] = line["host_s"] / t_gpu
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
