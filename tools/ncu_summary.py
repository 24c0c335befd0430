"""
Condense an .ncu-rep (ncu --set full capture) into the handful of numbers the roofline discussion needs.

    python tools/ncu_summary.py gpurun_out/prune_r01a.ncu-rep > profiles/r01a_prune_full.txt
"""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
    "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
    "smsp__warps_active.avg.per_cycle_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__inst_executed_pipe_fp64.sum",
    # FP64 tensor cores (DMMA): the pipe-utilisation figures of the 20- / 61-state kernels
    "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    rows = [r for r in rows if len(r) > 10]
    hdr, units = rows[0], rows[1]
    for rec in rows[2:]:
        name = rec[hdr.index("Kernel Name")]
        print("kernel:", name)
        for k in KEYS:
            match = [h for h in hdr if h == k or h.endswith("." + k)]
            if match:
                i = hdr.index(match[0])
                print("  {:70s} {:>18s} {}".format(k, rec[i], units[i]))
        stalls = []
        for i, h in enumerate(hdr):
            if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
                try:
                    stalls.append((float(rec[i]), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
                except ValueError:
                    pass
        print("  stall reasons (warps stalled per issued instruction):")
        for v, h in sorted(stalls, reverse=True)[:8]:
            print("    {:30s} {:8.3f}".format(h, v))


if __name__ == "__main__":
    main(sys.argv[1])
