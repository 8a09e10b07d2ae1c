#!/usr/bin/env python3
"""Kernel tuning helper: the kernel-only leg of bench.py alone (frames staged in HBM, S slots in flight).
  python tools/pipelined_fps.py [cfg4] [slots] [frames]      B200J2K_SKIP=<mask> leaves kernels out (marginal costs)"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "tools")]
import torch  # noqa: E402

import ffmpeg_ht_b200 as fb  # noqa: E402
import j2kgen  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
S = int(sys.argv[2]) if len(sys.argv) > 2 else 12
N = int(sys.argv[3]) if len(sys.argv) > 3 else 256
frames = [j2kgen.make_config(name, phase=0.37 * i)[0] for i in range(4)]
dec = fb.Jpeg2000Decoder(device=0, slots=S)
for s in range(S):
    dec.stage(s, frames[s % 4])
streams = [torch.cuda.ExternalStream(dec.stream(s), device=0) for s in range(S)]
for i in range(3 * S):
    dec.run_staged(i % S)
torch.cuda.synchronize()
ev0 = [torch.cuda.Event(enable_timing=True) for _ in streams]
ev1 = [torch.cuda.Event(enable_timing=True) for _ in streams]
for e, st in zip(ev0, streams):
    e.record(st)
for i in range(N):
    dec.run_staged(i % S)
for e, st in zip(ev1, streams):
    e.record(st)
torch.cuda.synchronize()
ms = max(a.elapsed_time(b) for a in ev0 for b in ev1)
print(f"{name} slots {S}: {N / ms * 1e3:.0f} frames/s, {ms / N * 1e3:.1f} us per frame", flush=True)
dec.close()
