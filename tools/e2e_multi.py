#!/usr/bin/env python3
"""End-to-end leg of bench.py alone, under torchrun, for host-side settings at N > 1 (decoder contexts per rank, frames
in flight per context, frames per draw from the shared counter):
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/e2e_multi.py \
      T:S:chunk [T:S:chunk ...]
Prints whole-job frames/s per setting (max over ranks of the wall time, frames drawn from one FrameQueue)."""
import os
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "tools")]
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import ffmpeg_ht_b200 as fb  # noqa: E402
import j2kgen  # noqa: E402
from ffmpeg_ht_b200 import sharding  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
store = sharding.default_store()
frames = [j2kgen.make_config("cfg4", phase=0.37 * i)[0] for i in range(4)]
info = fb.probe(frames[0])
pkts = [fb.PinnedPacket(f) for f in frames]
pic_bytes = sum(info.plane_row_bytes[i] * info.plane_height[i] for i in range(info.nplanes))
PER_RANK = 256
for n_set, spec in enumerate(sys.argv[1:]):
    T, S, chunk = (int(v) for v in spec.split(":"))
    decs = [fb.Jpeg2000Decoder(device=local, slots=S) for _ in range(T)]
    pins = [[fb.PinnedFrame(info) for _ in range(S)] for _ in range(T)]
    done = [0] * T

    def worker(t, fq):
        torch.cuda.set_device(local)
        d, pf, pending = decs[t], pins[t], []
        for n, i in enumerate(fq):
            if len(pending) == len(pf):
                d.wait(pending.pop(0))
            pending.append(d.submit(pkts[i % 4], pf[n % len(pf)]))
            done[t] += 1
        for s in pending:
            d.wait(s)

    def run(name, total):
        fq = sharding.FrameQueue(store, name, total, chunk)
        th = [threading.Thread(target=worker, args=(t, fq)) for t in range(T)]
        for x in th:
            x.start()
        for x in th:
            x.join()

    run(f"warm{n_set}", 32 * world)
    torch.cuda.synchronize()
    for t in range(T):
        done[t] = 0
    dist.barrier()
    t0 = time.perf_counter()
    run(f"timed{n_set}", PER_RANK * world)
    torch.cuda.synchronize()
    sec = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    mine = torch.zeros(world, dtype=torch.float64, device="cuda")
    mine[rank] = sum(done)
    dist.all_reduce(sec, op=dist.ReduceOp.MAX)
    dist.all_reduce(mine, op=dist.ReduceOp.SUM)
    if rank == 0:
        fps = PER_RANK * world / float(sec.item())
        print(f"contexts {T} x {S} in flight, draws of {chunk}: {fps:.0f} frames/s ({fps * pic_bytes / 1e9:.1f} GB/s of "
              f"pictures), frames per rank {[int(v) for v in mine.tolist()]}", flush=True)
    for d in decs:
        d.close()
    del pins
    dist.barrier()
dist.destroy_process_group()
