"""
Dynamic SASS profile of a kernel from an .ncu-rep captured with --set full --import-source on: executed warp
instructions per opcode (and per unit of work, e.g. per tile-row), shared-memory wavefronts and stall samples per
opcode, and the basic blocks that carry the instructions.  What tools/ncu_hotspots.py does per source line, per opcode.

    python tools/ncu_sass_profile.py gpurun_out/x.ncu-rep [units] [kernel-substring]

`units` = how many units of work the launch processed (15625 tiles x 999 rows for the headline walk); default 1.
"""
import collections
import csv
import re
import subprocess
import sys


def opname(src):
    src = re.sub(r"^@!?U?PT?\d*\s+", "", src.strip())
    op = src.split()[0] if src.split() else "?"
    if op.startswith("IMAD.MOV") or op == "MOV":
        return "MOV"
    parts = op.split(".")
    return ".".join(parts[:2]) if parts[0] in ("LDS", "STS", "LDG", "STG") and len(parts) > 1 and parts[1][0].isdigit() else parts[0]


def main(path, units=1.0, which=""):
    raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"],
                         stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    for si, h in enumerate(starts):
        name = rows[h][1]
        if which and which not in name:
            continue
        hdr = rows[h + 1]
        end = starts[si + 1] if si + 1 < len(starts) else len(rows)
        data = [r for r in rows[h + 2:end] if len(r) >= len(hdr)]
        i_s, i_e, i_a = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Address")
        i_w, i_n = hdr.index("L1 Wavefronts Shared"), hdr.index("# Samples")
        ops, wf, smp = collections.Counter(), collections.Counter(), collections.Counter()
        for r in data:
            o = opname(r[i_s])
            ops[o] += int(r[i_e] or 0)
            wf[o] += int(r[i_w] or 0)
            smp[o] += int(r[i_n] or 0)
        tot, tots = sum(ops.values()), max(1, sum(smp.values()))
        print("kernel:", name[:150])
        print("warp instructions executed: %d = %.1f per unit (%g units)" % (tot, tot / units, units))
        for o, v in ops.most_common(28):
            print("  %-12s %9.2f per unit %5.1f %%   shared wavefronts per unit %7.2f   stall samples %5.1f %%"
                  % (o, v / units, 100.0 * v / tot, wf[o] / units, 100.0 * smp[o] / tots))
        # runs of instructions with the same execution count = basic blocks as executed
        base = int(data[0][i_a], 16)
        print("blocks (>= 12 instructions, executed by >= 2 % of the units):")
        i = 0
        while i < len(data):
            j = i
            while j < len(data) and data[j][i_e] == data[i][i_e]:
                j += 1
            n = int(data[i][i_e] or 0)
            if j - i >= 12 and n >= 0.02 * units:
                c = collections.Counter(opname(r[i_s]) for r in data[i:j])
                print("  0x%05x  %4d instructions x %.3f per unit   %s" % (int(data[i][i_a], 16) - base, j - i, n / units,
                                                                       " ".join("%s=%d" % kv for kv in c.most_common(8))))
            i = j
        print()


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 1.0, sys.argv[3] if len(sys.argv) > 3 else "")
