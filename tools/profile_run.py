#!/usr/bin/env python3
"""Small driver for ncu and quick timing: stage one frame of a BASELINE config and re-run its kernels.
  python tools/profile_run.py [cfg4[,cfg2,...]] [runs]
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))
import ffmpeg_ht_b200 as fb  # noqa: E402
import j2kgen  # noqa: E402

names = (sys.argv[1] if len(sys.argv) > 1 else "cfg4").split(",")
runs = int(sys.argv[2]) if len(sys.argv) > 2 else 3
for name in names:
    cs, planes, p = j2kgen.make_config(name)
    with fb.Jpeg2000Decoder(device=0, slots=1) as dec:
        info = dec.stage(0, cs)
        ms = np.zeros(6)
        dec.run_staged_timed(0)
        for _ in range(runs):
            ms += np.array(dec.run_staged_timed(0))
        ms /= runs
        print(name, info.width, info.height, info.ncodeblocks, "stage ms [unstuff, vlc, magsgn, refine-blocks, idwt levels, final]",
              [round(float(x), 4) for x in ms], "total", round(float(ms.sum()), 4), flush=True)
