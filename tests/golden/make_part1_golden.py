#!/usr/bin/env python3
"""Generates tests/golden/frames_part1.npz with the REFERENCE (oracle/_ref/libj2kref.so built from /root/reference):
Part-1 (EBCOT / MQ) streams made by OpenJPEG through Pillow, plain and damaged -- reference verdict and pixels.
Same layout as frames_wide.npz (key = name|flags).  Run where the reference is mounted:
    python tests/golden/make_part1_golden.py"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path[:0] = [str(ROOT), str(ROOT / "tools"), str(ROOT / "tests")]
from helpers import ref_decode  # noqa: E402
from test_oracle_vs_ref import PART1, part1_stream  # noqa: E402

items = {}
for i, t in enumerate(PART1):
    for be in ((0,) if not t["kw"].get("irreversible") else (0, 1)):
        items[f"part1_{i:02d}_{t['mode']}|{be}"] = part1_stream(t, jp2=(i % 3 == 1))
rng = np.random.default_rng(79)
for i in (0, 3, 7):
    cs = part1_stream(PART1[i])
    sod = cs.index(b"\xff\x93")
    for k in range(4):
        bad = bytearray(cs)
        for _ in range(int(rng.integers(1, 6))):
            bad[int(rng.integers(sod + 2, len(bad) - 2))] = int(rng.integers(0, 256))
        if k == 3:
            bad = bad[:int(rng.integers(sod + 10, len(bad)))]
        items[f"part1_damaged_{i:02d}_{k}|0"] = bytes(bad)
pack = {}
for k, data in items.items():
    r, info, out, pix = ref_decode(data, int(k.split("|")[1]))
    pack[k + "|cs"] = np.frombuffer(data, np.uint8)
    pack[k + "|meta"] = np.array([info[0], info[1], r], np.int64)
    pack[k + "|pix"] = np.frombuffer((pix or "none").encode(), np.uint8)
    pack[k + "|out"] = out if r == 0 else np.zeros(0, np.uint8)
np.savez_compressed(ROOT / "tests" / "golden" / "frames_part1.npz", **pack)
print("wrote frames_part1.npz:", len(items), "streams,", sum(1 for k in items if pack[k + "|meta"][2] == 0), "decodable")
