#!/usr/bin/env python
"""Per-role view of an ncu report of k_gather_tmem captured with --import-source on:
the kernel's warps are specialised (producer / fill / consumer, separated by the
USETMAXREG instructions), so stall samples are attributed per role and the hottest
instructions of each role are listed.
Usage: python tools/ncu_roles.py gpurun_out/x.ncu-rep [n_top]"""
import csv
import subprocess
import sys


def main(path, n_top=12):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, data = rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]

    def num(r, k):
        return int(r[col[k]] or 0)

    idx = [i for i, r in enumerate(data) if "USETMAXREG" in r[col["Source"]]]
    bounds = [0] + idx + [len(data)]
    names = ["prologue", "producer", "fill", "consumer"][:len(bounds) - 1]
    tot = sum(num(r, "# Samples") for r in data)
    print(f"total samples {tot} (= 16 warps x kernel duration)")
    for k, name in enumerate(names):
        seg = data[bounds[k]:bounds[k + 1]]
        s = sum(num(r, "# Samples") for r in seg)
        d = {h: sum(num(r, h) for r in seg) for h in stalls}
        top = sorted(d.items(), key=lambda x: -x[1])[:6]
        print(f"\n{name}: {100 * s / tot:.1f}% of samples; "
              + ", ".join(f"{a.replace('stall_', '')} {100 * b / max(s, 1):.0f}%" for a, b in top))
        for r in sorted(seg, key=lambda r: -num(r, "# Samples"))[:n_top]:
            d = {h: num(r, h) for h in stalls}
            t = sorted(d.items(), key=lambda x: -x[1])[:2]
            print(f"  {num(r, '# Samples'):8d} {num(r, 'Instructions Executed'):10d}  "
                  f"{r[col['Source']].strip()[:58]:58s} "
                  + ", ".join(f"{a.replace('stall_', '')} {b}" for a, b in t))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 12)
