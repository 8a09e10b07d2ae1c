#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel share of device time.
  python tools/summarize_launches.py gpurun_out/r1_launches.csv > profiles/r1_launch_summary.md
"""
import csv
import re
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
agg = defaultdict(lambda: [0, 0.0])
for r in rows:
    name = re.sub(r"\(.*", "", r[4]).replace("void ", "").replace("<unnamed>::", "").strip()
    agg[(name, r[8])][0] += 1
    agg[(name, r[8])][1] += float(r[14])
total = sum(v[1] for v in agg.values())
print(f"# ncu launch list summary ({len(rows)} launches, {total / 1e6:.2f} ms of device time; cold-cache, serialised)\n")
print("| kernel | grid | launches | total ms | avg us | share |")
print("|---|---|---|---|---|---|")
for (name, grid), (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{name}` | {grid} | {n} | {t / 1e6:.3f} | {t / n / 1e3:.1f} | {100 * t / total:.1f}% |")
