#!/usr/bin/env python
"""Per-opcode view of an .ncu-rep captured with --import-source on: share of warp-stall
samples and of executed instructions per SASS opcode, shared-memory wavefronts per
memory instruction, and instruction counts of the kernel's hottest loop.
Usage: python profiles/summarize_source.py gpurun_out/x.ncu-rep >> profiles/x.txt"""
import csv
import subprocess
import sys
from collections import defaultdict


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, data = rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}

    def num(r, name):
        return int(r[col[name]] or 0)

    tot = sum(num(r, "# Samples") for r in data) or 1
    totex = sum(num(r, "Instructions Executed") for r in data) or 1
    agg = defaultdict(lambda: [0, 0, 0, 0, 0])
    for r in data:
        words = r[col["Source"]].split()
        if not words:
            continue
        op = (words[1] if words[0].startswith("@") else words[0]).split(".")[0]
        a = agg[op]
        a[0] += num(r, "# Samples")
        a[1] += num(r, "Instructions Executed")
        a[2] += num(r, "stall_wait")
        a[3] += num(r, "stall_no_inst")
        a[4] += num(r, "stall_branch_resolving")
    print(f"per-opcode profile ({tot} stall samples, {totex} warp instructions)")
    print(f"  {'opcode':10s} {'samples':>8s} {'executed':>9s}   wait / no_inst / branch samples")
    for op, a in sorted(agg.items(), key=lambda x: -x[1][0])[:14]:
        print(f"  {op:10s} {100 * a[0] / tot:7.1f}% {100 * a[1] / totex:8.1f}%   "
              f"{a[2]:8d} {a[3]:8d} {a[4]:8d}")
    print("shared-memory instructions (wavefronts per warp instruction: measured / ideal)")
    seen = set()
    for r in data:
        wf, ideal, ex = (num(r, "L1 Wavefronts Shared"), num(r, "L1 Wavefronts Shared Ideal"),
                         num(r, "Instructions Executed"))
        words = r[col["Source"]].split()
        if wf and ex > totex / 1000 and words:
            op = words[1] if words[0].startswith("@") else words[0]
            key = (op, round(wf / ex, 1), round(ideal / ex, 1))
            if key not in seen:
                seen.add(key)
                print(f"  {op:24s} {wf / ex:5.1f} / {ideal / ex:4.1f}")


if __name__ == "__main__":
    main(sys.argv[1])
