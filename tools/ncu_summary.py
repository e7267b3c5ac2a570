#!/usr/bin/env python
"""Print the metrics we care about from an .ncu-rep (read on the GPU-less box with `ncu -i`)."""
import csv
import io
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
    "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.sum",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.avg.per_cycle_active", "sm__cycles_elapsed.avg", "sm__cycles_active.avg",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum",
]


def main(path, kernel_filter=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        if kernel_filter and kernel_filter not in name:
            continue
        print("==", name[:70])
        for i, h in enumerate(hdr):
            if h in WANT or "warp_issue_stalled" in h and h.endswith("_per_warp_active.pct"):
                try:
                    v = float(r[i].replace(",", ""))
                except ValueError:
                    continue
                if "stalled" in h and v < 2.0:
                    continue
                print(f"  {h:75s} {r[i]:>18s} {units[i]}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)


def stalls(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[0]
    for r in rows[2:]:
        items = []
        for i, h in enumerate(hdr):
            if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h:
                try:
                    items.append((float(r[i].replace(",", "")), h.replace("smsp__pcsamp_warps_issue_stalled_", "")))
                except ValueError:
                    pass
        tot = sum(f for f, _ in items) or 1.0
        print("  warp stall samples (share of all samples):")
        for f, k in sorted(items, reverse=True)[:9]:
            print(f"    {k:28s} {100 * f / tot:5.1f} %")


if __name__ == "__main__" and len(sys.argv) > 1:
    stalls(sys.argv[1])
