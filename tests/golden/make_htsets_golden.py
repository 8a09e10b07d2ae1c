#!/usr/bin/env python3
"""Generates tests/golden/frames_htsets.npz with the REFERENCE (oracle/_ref/libj2kref.so built from /root/reference):
HTDECLARED and MIXED streams (blocks of one frame on both block coders), MIXED streams with random byte damage, and
the Psot bound cases, palettised JP2 files (PAL8: index plane + palette), Part-1 streams re-labelled with
other code-block styles (context reset, causal contexts, segmentation symbols), D-Cinema profile
streams (XYZ12 and the planar formats of that list) -- reference verdict and a SHA-1 digest of its picture.  Same layout as frames_tiles.npz
(key = name|flags).  Run where the reference is mounted:  python tests/golden/make_htsets_golden.py"""
import hashlib
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path[:0] = [str(ROOT), str(ROOT / "tools"), str(ROOT / "tests")]
from helpers import ref_decode  # noqa: E402
from test_oracle_vs_ref import (HT_SETS, PAL8_CASES, dcinema_cases, ht_set_damaged_streams, ht_set_stream,  # noqa: E402
                                pal8_file, part1_styled_streams, psot_bound_cases)

items = {}
for i, t in enumerate(HT_SETS):
    for be in ((0,) if t["cfg"].get("reversible", True) else (0, 1)):
        items[f"ht_set{i:02d}|{be}"] = ht_set_stream(t)[0]
for i, bad in enumerate(ht_set_damaged_streams()):
    if i % 5 < 3:
        continue            # two in five keep the fixture small
    items[f"ht_set_damaged{i:02d}|0"] = bad
for i, c in enumerate(psot_bound_cases()):
    items[f"psot{i}|0"] = c
for i, (name, c) in enumerate(part1_styled_streams()):
    if i % 8 < 6 and (i // 8) in (0, 3, 5):          # three layouts x six served styles
        items[f"part1_styled{i:02d}|0"] = c
for i, (name, c) in enumerate(dcinema_cases()):
    if "x8" not in name and "8-bit" not in name and i % 3 != 1:
        items[f"dcinema{i:02d}|0"] = c
for i, (bd, csp, pal) in enumerate(PAL8_CASES):
    items[f"pal8_{i}|0"] = pal8_file(bd, csp, pal, tile=(64, 32))[0]
pack = {}
for k, data in items.items():
    r, info, out, pix = ref_decode(data, int(k.split("|")[1]))
    pack[k + "|cs"] = np.frombuffer(data, np.uint8)
    pack[k + "|meta"] = np.array([info[0], info[1], r], np.int64)
    pack[k + "|pix"] = np.frombuffer((pix or "none").encode(), np.uint8)
    pack[k + "|out"] = np.frombuffer(hashlib.sha1(out.tobytes()).digest(), np.uint8) if r == 0 else np.zeros(0, np.uint8)
np.savez_compressed(ROOT / "tests" / "golden" / "frames_htsets.npz", **pack)
print("wrote frames_htsets.npz:", len(items), "streams,", sum(1 for k in items if pack[k + "|meta"][2] == 0), "decodable")
