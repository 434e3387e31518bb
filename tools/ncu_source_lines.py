#!/usr/bin/env python
"""tools/ncu_source_lines.py <report.ncu-rep> [top] -- per CUDA source line: warp instructions executed, stall
samples and shared-memory wavefronts (ideal / excessive), from the source page of an ncu report."""
import csv
import subprocess
import sys
from collections import defaultdict

rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
agg = defaultdict(lambda: [0, 0, 0, 0, ""])
fname, hdr = "", None
for r in rows:
    if len(r) == 2 and r[0] in ("File Path", "File Name"):
        fname = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        col = {h: i for i, h in enumerate(hdr)}
        continue
    if hdr is None or len(r) < len(hdr) or not r[0].isdigit():
        continue
    key = (fname, int(r[0]))
    a = agg[key]

    def num(name):
        v = r[col[name]].replace(",", "").split("(")[0]
        try:
            return int(float(v)) if v not in ("", "-") else 0
        except ValueError:      # a source line with commas / quotes of its own shifted the columns
            return 0
    a[0] += num("Instructions Executed")
    a[1] += num("# Samples")
    a[2] += num("L1 Wavefronts Shared")
    a[3] += num("L1 Wavefronts Shared Excessive")
    a[4] = r[1].strip()[:90]
tot_i = sum(a[0] for a in agg.values()) or 1
tot_s = sum(a[1] for a in agg.values()) or 1
print(f"total warp instructions {tot_i}, samples {tot_s}")
print(f"{'file:line':28s} {'inst%':>6s} {'samp%':>6s} {'smem wavefronts':>16s} {'excessive':>10s}  source")
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{key[0] + ':' + str(key[1]):28s} {100 * a[0] / tot_i:6.2f} {100 * a[1] / tot_s:6.2f} {a[2]:16d} {a[3]:10d}  {a[4]}")
