"""
SASS opcode histogram per kernel of libphylo_b200.so (cuobjdump -sass), for the kernels whose names match the given
regular expressions (default: the hot kernels of DESIGN.md section 3).  Runs on the build box - no GPU needed.

    python tools/sass_histogram.py [regex ...] > profiles/r02_sass_histogram.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "phylo_utils_b200", "libphylo_b200.so")
HOT = [r"dna_pair_cta_kernel", r"dna_pair_kernel", r"dna_pair_store_kernel", r"dna_up_kernel", r"dna_edge_st_kernel", r"dna_edge_sumtable_kernel",
       r"mma_prune_kernel", r"mma_edge_deriv_kernel", r"edge_st_kernel", r"dna_prune_kernel", r"dna_root_kernel",
       r"pmatrix_kernel", r"tip_table_kernel"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), stdout=subprocess.PIPE, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    pats = [re.compile(p) for p in (sys.argv[1:] or HOT)]
    sass = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, text=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True).stdout
    usage = {}
    for m in re.finditer(r"Function (\S+):\s*\n\s*(REG:\d+.*)", res):
        usage[m.group(1)] = m.group(2).strip()
    funcs = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = funcs.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_]+)*)", line)
        if m and cur is not None:
            op = m.group(1)
            # keep the width / kind suffix where it says something about the memory system or the maths
            suffix = m.group(2)
            if op in ("LDS", "STS", "LDG", "STG", "LDGSTS", "LD", "ST", "LDC", "ATOMG", "RED"):
                width = re.search(r"\.(U?8|U?16|32|64|128|256)\b", suffix)
                op += "." + (width.group(1) if width else "32")
            cur[op] += 1
    names = demangle(list(funcs))
    print("SASS opcode histogram, {} (sm_100a), cuobjdump -sass".format(os.path.relpath(LIB, ROOT)))
    print("legend: DFMA/DMUL/DADD = fp64 pipe; DMMA = fp64 tensor pipe; LDGSTS = cp.async; LDS/STS = shared memory;")
    print("        no HMMA/UTC*MMA/UTMALDG expected: the path is fp64 (tcgen05 has no fp64 kind) and stages 16-byte pieces\n")
    for raw, counts in funcs.items():
        name = names.get(raw, raw)
        if not any(p.search(name) for p in pats):
            continue
        total = sum(counts.values())
        print("{}\n  {} instructions; {}".format(name, total, usage.get(raw, "")))
        fp64 = sum(v for k, v in counts.items() if k in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"))
        smem = sum(v for k, v in counts.items() if k.startswith(("LDS", "STS")))
        print("  fp64 pipe {} ({:.1f} %), DMMA {}, shared-memory accesses {} ({:.1f} %), cp.async {}".format(
            fp64, 100.0 * fp64 / max(total, 1), counts.get("DMMA", 0), smem, 100.0 * smem / max(total, 1),
            sum(v for k, v in counts.items() if k.startswith("LDGSTS"))))
        print("  " + ", ".join("{} {}".format(k, v) for k, v in counts.most_common(28)))
        print()


if __name__ == "__main__":
    main()
