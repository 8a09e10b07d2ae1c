#!/usr/bin/env python3
"""Frames per second of the FFCodec shim (oracle/_ref/libj2kb200shim.so: the shim compiled against the reference's
headers and driven by the test harness the way libavcodec drives a decoder, one context per thread = frame threading)
on the bench workload, with the shim's pinned frame pool on and off.   python tests/shim_throughput.py [threads]"""
import ctypes
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "tools")]
import bench  # noqa: E402

threads = int(sys.argv[1]) if len(sys.argv) > 1 else 4
frames = bench.make_frames(4)
lib = ctypes.CDLL(str(ROOT / "oracle" / "_ref" / "libj2kb200shim.so"))
fn = lib.b200shim_decode_bench
fn.restype = ctypes.c_double
fn.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]
n = 16
bufs = [np.frombuffer(frames[i % len(frames)], np.uint8) for i in range(n)]
ptrs = (ctypes.c_void_p * n)(*[b.ctypes.data for b in bufs])
lens = (ctypes.c_int64 * n)(*[b.size for b in bufs])
for pinned in ("1", "0"):
    os.environ["REF_OPT_pinned_frames"] = pinned
    fn(ptrs, lens, n, threads, 1, 0)
    reps = 24
    s = fn(ptrs, lens, n, threads, reps, 0)
    print(f"pinned_frames={pinned} threads={threads}: {n * reps / s:.0f} frames/s (4K 4:4:4 12-bit, pageable packets, "
          f"a new AVFrame per picture)")
