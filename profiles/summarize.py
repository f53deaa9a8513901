#!/usr/bin/env python
"""Turns an .ncu-rep (brought back from gpurun_out/) into the small text summary committed here.

    python profiles/summarize.py gpurun_out/prof.ncu-rep > profiles/r01_<kernel>.txt
"""
import collections
import csv
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active",
]


def main(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    print(f"# {rep}: {len(rows) - 2} profiled launch(es); ncu --set full --clock-control none")
    for n, r in enumerate(rows[2:]):
        print(f"\n## launch {n}: {r[ki][:110]}")
        for m in METRICS:
            if m in hdr:
                i = hdr.index(m)
                print(f"{m:85s} {r[i]:>16s} {units[i]}")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    hdr, agg, k = None, collections.Counter(), 0
    for r in rows:
        if r and r[0] == "Kernel Name":
            k += 1
            continue
        if r and r[0] == "Address":
            hdr = r
            continue
        if k != 1 or hdr is None:
            continue
        for i, h in enumerate(hdr):
            if h.startswith("stall_") and "Not Issued" not in h:
                try:
                    agg[h] += int(r[i])
                except (ValueError, IndexError):
                    pass
    tot = sum(agg.values()) or 1
    print("\n## warp stall sampling, first launch (share of samples)")
    for h, v in agg.most_common(10):
        print(f"{h:28s} {100 * v / tot:5.1f} %")


if __name__ == "__main__":
    main(sys.argv[1])
