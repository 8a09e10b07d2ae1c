#!/usr/bin/env python3
"""Frames/s of the FFCodec shim through the reference-side harness (bench.py's e2e_ffcodec_shim leg alone).
  python tools/shim_fps.py [contexts ...]"""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "tools")]
import bench  # noqa: E402

ctx = tuple(int(a) for a in sys.argv[1:]) or (1, 2, 4, 8, 16)
print(json.dumps(bench.shim_leg(bench.make_frames(4), ctx), indent=1))
