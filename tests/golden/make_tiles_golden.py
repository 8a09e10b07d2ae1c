#!/usr/bin/env python3
"""Generates tests/golden/frames_tiles.npz with the REFERENCE (oracle/_ref/libj2kref.so built from /root/reference):
several (interleaved) tile-parts per tile, tiles with their own COD / QCD / POC, damaged tile-header markers --
reference verdict and pixels.  Same layout as frames_wide.npz (key = name|flags).  Run where the reference is
mounted:  python tests/golden/make_tiles_golden.py"""
import hashlib
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path[:0] = [str(ROOT), str(ROOT / "tools"), str(ROOT / "tests")]
from helpers import ref_decode  # noqa: E402
from test_oracle_vs_ref import (TILE_PARTS, TILE_STYLES, tile_part_stream, tile_style_marker_cases,  # noqa: E402
                                tile_style_stream)

items = {}
for i, t in enumerate(TILE_PARTS):
    items[f"tile_parts{i}|0"] = tile_part_stream(t)[0]
for i, t in enumerate(TILE_STYLES):
    for be in ((0,) if t["cfg"].get("reversible", True) else (0, 1)):
        items[f"tile_styles{i}|{be}"] = tile_style_stream(t)[0]
for i, bad in enumerate(tile_style_marker_cases()):
    items[f"tile_marker{i:02d}|0"] = bad
pack = {}
for k, data in items.items():
    r, info, out, pix = ref_decode(data, int(k.split("|")[1]))
    pack[k + "|cs"] = np.frombuffer(data, np.uint8)
    pack[k + "|meta"] = np.array([info[0], info[1], r], np.int64)
    pack[k + "|pix"] = np.frombuffer((pix or "none").encode(), np.uint8)
    # the pictures are kept as SHA-1 digests: 31 pictures of 144 KB would be a 3 MB fixture
    pack[k + "|out"] = np.frombuffer(hashlib.sha1(out.tobytes()).digest(), np.uint8) if r == 0 else np.zeros(0, np.uint8)
np.savez_compressed(ROOT / "tests" / "golden" / "frames_tiles.npz", **pack)
print("wrote frames_tiles.npz:", len(items), "streams,", sum(1 for k in items if pack[k + "|meta"][2] == 0), "decodable")
