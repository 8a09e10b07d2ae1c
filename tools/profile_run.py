#!/usr/bin/env python3
"""Small driver for ncu: stage one frame of a BASELINE config and re-run its kernels a few times.
  python tools/profile_run.py [cfg4] [runs]
"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))
import ffmpeg_ht_b200 as fb  # noqa: E402
import j2kgen  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
runs = int(sys.argv[2]) if len(sys.argv) > 2 else 3
cs, planes, p = j2kgen.make_config(name)
with fb.Jpeg2000Decoder(device=0, slots=1) as dec:
    info = dec.stage(0, cs)
    for _ in range(runs):
        ms = dec.run_staged_timed(0)
    print(name, info.width, info.height, info.ncodeblocks, "stage ms", ms)
