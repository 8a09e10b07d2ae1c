#!/usr/bin/env python3
"""Experiment: how many blocks should share a warp in ht_vlc_kernel (lane per block)?  Run once per setting:
   B200J2K_VLC_ACTIVE=8 python tools/vlc_lanes_experiment.py
Prints the per-stage latency of one 4K frame alone on a stream and the pipelined kernel-only frame rate (8 frames
in flight), for the bench workload and for the 8K grey one."""
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "tools")]
import torch  # noqa: E402

import bench  # noqa: E402
import ffmpeg_ht_b200 as fb  # noqa: E402

torch.cuda.set_device(0)
frames = bench.make_frames(4)
S = int(os.environ.get("SLOTS", "8"))
dec = fb.Jpeg2000Decoder(device=0, slots=S)
for s in range(S):
    dec.stage(s, frames[s % len(frames)])
for _ in range(3):
    for i in range(64):
        dec.run_staged(i % S)
for s in range(S):
    dec.sync(s)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    for i in range(64):
        dec.run_staged(i % S)
for s in range(S):
    dec.sync(s)
dt = time.perf_counter() - t0
st = np.zeros(6)
for _ in range(2):
    dec.run_staged_timed(0)
for _ in range(10):
    st += np.array(dec.run_staged_timed(0))
print("active", os.environ.get("B200J2K_VLC_ACTIVE", "default"), "slots", S, "pipelined fps %.0f" % (320 / dt),
      "stage ms", [round(float(x) / 10, 4) for x in st])
dec.close()
