#!/usr/bin/env python3
"""Generates tests/golden/frames_packed.npz with the REFERENCE (oracle/_ref/libj2kref.so built from
/root/reference): streams whose packet headers sit in PPT / PPM marker segments, damaged PPT / PPM markers and
tiles without a first tile-part -- reference verdict and pixels for each.  Same layout as frames_wide.npz
(key = name|flags).  Run where the reference is mounted:  python tests/golden/make_packed_golden.py"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path[:0] = [str(ROOT), str(ROOT / "tools"), str(ROOT / "tests")]
import j2kgen  # noqa: E402
from helpers import ref_decode  # noqa: E402
from test_oracle_vs_ref import PACKED, packed_marker_cases, packed_stream  # noqa: E402

items = {}
for i, t in enumerate(PACKED):
    cs, _ = packed_stream(t)
    for be in ((0,) if t["cfg"].get("reversible", True) else (0, 1)):
        items[f"packed{i}_{'ppt' if t['mode'] == 1 else 'ppm'}|{be}"] = cs
for i, bad in enumerate(packed_marker_cases()):
    items[f"marker{i:02d}|0"] = bad
planes = j2kgen.synth_image(96, 64, 3, 8, seed=5)
cs = j2kgen.encode(planes, 8, j2kgen.EncParams(levels=2, tile=(64, 64)))
sot = [i for i in range(len(cs) - 3) if cs[i:i + 4] == b"\xff\x90\x00\x0a"]
renumbered = bytearray(cs)
renumbered[sot[1] + 10] = 1
for i, bad in enumerate((cs[:sot[1]], cs[:sot[1] + 3], bytes(renumbered))):
    items[f"tile_without_first_part{i}|0"] = bad
pack = {}
for k, data in items.items():
    r, info, out, pix = ref_decode(data, int(k.split("|")[1]))
    pack[k + "|cs"] = np.frombuffer(data, np.uint8)
    pack[k + "|meta"] = np.array([info[0], info[1], r], np.int64)
    pack[k + "|pix"] = np.frombuffer((pix or "none").encode(), np.uint8)
    pack[k + "|out"] = out if r == 0 else np.zeros(0, np.uint8)
np.savez_compressed(ROOT / "tests" / "golden" / "frames_packed.npz", **pack)
print("wrote frames_packed.npz:", len(items), "streams,", sum(1 for k in items if pack[k + "|meta"][2] == 0), "decodable")
