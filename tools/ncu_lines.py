#!/usr/bin/env python3
"""Aggregate an ncu SASS source page by CUDA source line.
  ncu -i rep.ncu-rep --page source --csv --kernel-name regex:K > sass.csv
  nvdisasm -g -c kernel.cubin > k.sass
  python tools/ncu_lines.py sass.csv k.sass <mangled-kernel-substring> [top]
"""
import csv
import re
import sys

sass_csv, disasm, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
# address -> line from nvdisasm
addr2line = {}
cur = None
insec = False
for ln in open(disasm, errors="ignore"):
    if ln.startswith("//---") and ".text." in ln:
        insec = kname in ln
    if not insec:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m and cur:
        addr2line[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(sass_csv)))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
ia, ii, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
base = None
agg = {}
for r in rows[h + 1:]:
    try:
        a = int(r[ia], 16)
    except ValueError:
        continue
    if base is None:
        base = a
    key = addr2line.get(a - base, ("?", 0))
    e = agg.setdefault(key, [0, 0])
    e[0] += int(r[ii] or 0)
    e[1] += int(r[isamp] or 0)
ti = sum(v[0] for v in agg.values())
ts = sum(v[1] for v in agg.values())
print(f"total warp instructions {ti}, stall samples {ts}")
srcs = {}
for (f, l), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    text = ""
    try:
        if f not in srcs:
            import glob
            cand = glob.glob(f"**/{f}", recursive=True)
            srcs[f] = open(cand[0]).read().split("\n") if cand else []
        text = srcs[f][l - 1].strip() if srcs[f] else ""
    except Exception:
        pass
    print(f"{100 * v[0] / ti:5.1f}% inst {100 * v[1] / max(ts, 1):5.1f}% stall  {f}:{l:<4} {text[:100]}")
