"""
Condense an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` launch list into one
line per launch.

    python tools/launch_list.py gpurun_out/launches.csv "what was run" > profiles/rNN_launches_x.txt
"""
import csv
import sys
from collections import OrderedDict


def main(path, what):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ik, im, iv, iid = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
    d = OrderedDict()
    for r in rows[1:]:
        d.setdefault(r[iid], {"k": r[ik]})[r[im]] = float(r[iv].replace(",", ""))
    print("# " + what)
    print("# per launch, serialised by ncu (--clock-control none), cold cache")
    for i, v in d.items():
        t, rd, wr = v.get("gpu__time_duration.sum", 0), v.get("dram__bytes_read.sum", 0), v.get("dram__bytes_write.sum", 0)
        print("%3s %10.1f us   DRAM read %7.2f GB  written %7.2f GB   %s" % (i, t / 1e3, rd / 1e9, wr / 1e9, v["k"][:100]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "")
