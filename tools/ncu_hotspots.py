"""
Source-line hot spots of a kernel from an .ncu-rep captured with --import-source on (and -lineinfo):
per CUDA source line its share of the stall samples, of the executed instructions, and its two main stall reasons.

    python tools/ncu_hotspots.py gpurun_out/x.ncu-rep [n_lines]
"""
import csv
import subprocess
import sys

NAMES = ["stall_long_sb", "stall_wait", "stall_short_sb", "stall_barrier", "stall_math", "stall_mio", "stall_no_inst",
         "stall_selected", "stall_not_selected", "stall_dispatch", "stall_branch_resolving", "stall_lg"]


def main(path, n_lines=24):
    raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Line No" and "# Samples" in r]
    for si, h in enumerate(starts):
        hdr = rows[h]
        end = starts[si + 1] if si + 1 < len(starts) else len(rows)
        title = [r for r in rows[max(0, h - 3):h] if r and r[0] in ("Function Name", "File Path")]
        data = [r for r in rows[h + 1:end] if len(r) == len(hdr) and r[2] == "-"]
        i_s, i_e = hdr.index("# Samples"), hdr.index("Instructions Executed")
        idx = {n: hdr.index(n) for n in NAMES if n in hdr}
        tot = sum(int(r[i_s] or 0) for r in data)
        tote = sum(int(r[i_e] or 0) for r in data)
        if tot < 1000:
            continue
        for t in title:
            print(t[0] + ":", t[1][:160])
        agg = {n: sum(int(r[i] or 0) for r in data) for n, i in idx.items()}
        print("samples", tot, " ".join("%s=%.1f%%" % (k[6:], 100.0 * v / tot) for k, v in sorted(agg.items(), key=lambda x: -x[1])[:8]))
        top = sorted(data, key=lambda r: -int(r[i_s] or 0))[:n_lines]
        for r in sorted(top, key=lambda r: int(r[0])):
            st = sorted(idx, key=lambda n: -int(r[idx[n]] or 0))[:2]
            print(r[0].rjust(5), "%5.1f%% smp %5.1f%% ins " % (100.0 * int(r[i_s]) / tot, 100.0 * int(r[i_e] or 0) / max(1, tote)),
                  " ".join("%s=%.0f%%" % (n[6:], 100.0 * int(r[idx[n]] or 0) / max(1, int(r[i_s]))) for n in st).ljust(34),
                  r[1].strip()[:100])
        print()


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 24)
