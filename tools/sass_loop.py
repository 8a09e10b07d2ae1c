#!/usr/bin/env python3
"""Kernel tuning helper: static SASS statistics of one kernel in an object file.
   python tools/sass_loop.py <file.o> <substring of the mangled name>
Prints the instruction count, the backward branches (loops) with their body length and an opcode histogram
of each loop body."""
import collections
import re
import subprocess
import sys

obj, pat = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
funcs = txt.split("Function : ")
for f in funcs[1:]:
    name = f.split("\n", 1)[0]
    if pat not in name:
        continue
    ins = []
    for l in f.split("\n"):
        m = re.match(r"\s*/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    print(name[:150], "\n  instructions:", len(ins))
    for ad, t in ins:
        m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d+,\s*)?(0x[0-9a-f]+)", t)
        if m and int(m.group(1), 16) < ad:
            tgt = int(m.group(1), 16)
            body = [x for x in ins if tgt <= x[0] <= ad]
            c = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", x[1]).split()[0].split(".")[0] for x in body)
            print(f"  loop {tgt:#x}..{ad:#x}: {len(body)} instr", c.most_common(14))
