#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.
Usage: python profiles/summarize_launches.py gpurun_out/launches.csv > profiles/x.txt
(per-launch times under ncu are cold-cache and serialised: compare SHARES)."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
hdr = rows[hi]
kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hi + 1:]:
    if len(r) <= mv:
        continue
    name = r[kn].split("(")[0].replace("void ", "")
    if "::k_" not in name and not name.startswith("k_"):
        name = "(torch index/sort/fill helpers of table building)"
    agg[name][0] += 1
    agg[name][1] += float(r[mv]) / 1e6
tot = sum(v[1] for v in agg.values())
print(f"{'kernel':52s} {'launches':>8s} {'total ms':>10s} {'share':>7s} {'avg ms':>9s}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:52s} {v[0]:8d} {v[1]:10.2f} {100 * v[1] / tot:6.1f}% {v[1] / v[0]:9.3f}")
ex = {k: v for k, v in agg.items() if "k_gather" in k or "k_mix" in k or "k_init" in k}
te = sum(v[1] for v in ex.values())
print("\nexchange step only (gather + mix + init):")
for k, v in sorted(ex.items(), key=lambda kv: -kv[1][1]):
    print(f"  {k:50s} share of step {100 * v[1] / te:5.1f}%")
