#!/usr/bin/env python
"""Condense an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals and shares.
    python tools/launch_summary.py gpurun_out/launches.csv profiles/r1_launches_bench.txt"""
import csv
import sys
from collections import OrderedDict

src, out = sys.argv[1], sys.argv[2]
rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = OrderedDict()
for r in rows[1:]:
    name = r[ki].split("(")[0].replace("void ", "")
    v = float(r[vi].replace(",", ""))
    scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1e-6)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v * scale
total = sum(a[1] for a in agg.values())
lines = [f"# per-kernel device time from {src} (ncu launch list: cold-cache, serialised -> compare SHARES, not absolutes)",
         f"# total {total:.3f} ms over {sum(a[0] for a in agg.values())} launches", f"{'launches':>8s} {'total_ms':>12s} {'share':>7s}  kernel"]
for name, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    lines.append(f"{n:8d} {ms:12.3f} {100 * ms / total:6.2f}%  {name}")
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
