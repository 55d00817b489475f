#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into the handful of numbers DESIGN.md and
bench.py quote.  Usage: python profiles/summarize.py gpurun_out/x.ncu-rep > profiles/x.txt"""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__m_xbar2l1tex_read_bytes.sum",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum",
    "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__sass_inst_executed_op_shared_ld.sum",
    "smsp__warps_eligible.avg.per_cycle_active", "sm__cycles_elapsed.avg",
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("kernel:", r[hdr.index("Kernel Name")])
        for k in KEYS:
            if k in hdr:
                print(f"  {k:72s} {r[hdr.index(k)]:>22s} {units[hdr.index(k)]}")
        stalls = [(float(r[i].replace(",", "")), h) for i, h in enumerate(hdr)
                  if h.startswith("smsp__pcsamp_warps_issue_stalled") and "not_issued" not in h]
        tot = sum(v for v, _ in stalls) or 1.0
        print("  top warp-stall reasons (share of samples):")
        for v, h in sorted(stalls, reverse=True)[:6]:
            print(f"    {h.replace('smsp__pcsamp_warps_issue_stalled_', ''):28s} {100 * v / tot:5.1f} %")


if __name__ == "__main__":
    main(sys.argv[1])
