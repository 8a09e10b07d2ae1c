"""Multi-GPU partitioning of the decode path.

Frames of a sequence and tiles of one oversized frame are independent units: the reference decodes them on
independent frame-thread contexts (libavcodec/pthread_frame.c:248-295) and independent execute2 tile jobs
(jpeg2000_decode_tile, libavcodec/jpeg2000dec.c:2368-2395, :2894).  Here they are sharded round-robin over GPUs with
NO data-path collective:

  frames   rank r of N decodes frames r, r + N, ...            (shard_indices; bench.py's kernel-only leg), or
           the ranks draw runs of frames from one counter        (FrameQueue; bench.py's end-to-end legs at N > 1:
           end to end a GPU is as fast as its PCIe path, and on an 8-GPU host those differ by half -- equal shares
           wait for the slowest one)
  tiles    context r of N decodes tiles r, r + N, ... of EVERY frame: a decoder context created with
           B200J2K_FLAG_TILES(r, N) parses only its tiles' packets, uploads only their tile-parts, reconstructs
           only them and writes only their rectangles of the output planes (tile_flags; TileShardedDecoder drives
           one such context per GPU of this process into one shared frame, the way one libavcodec process would)

torch.distributed is used only for the barrier around a timed region and for reducing the bookkeeping (max of the
elapsed time, sum of the unit counts)."""
from __future__ import annotations


def shard_indices(n_units: int, rank: int, world: int) -> list[int]:
    """Unit i (frame, or tile of a frame) goes to rank i % world."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return list(range(rank, n_units, world))


def owner(unit: int, world: int) -> int:
    return unit % world


def tile_flags(rank: int, world: int) -> int:
    """B200J2K_FLAG_TILES(rank, world) of include/b200j2k.h: flags of the decoder context that owns tiles
    rank, rank + world, ... (0 = every tile when there is only one context)."""
    if world < 1 or world > 255 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return 0 if world == 1 else ((world & 255) << 16) | ((rank & 255) << 24)


class TileShardedDecoder:
    """One oversized frame over several GPUs of this process: context i (on devices[i]) owns tiles i, i + n, ...
    All contexts write their tiles' rectangles into the SAME page-locked frame; decode() returns when every one has."""

    def __init__(self, devices, flags: int = 0, slots: int = 1):
        import ffmpeg_ht_b200 as fb
        self.devices = list(devices)
        n = len(self.devices)
        self.ctx = [fb.Jpeg2000Decoder(device=d, slots=slots, flags=flags | tile_flags(i, n))
                    for i, d in enumerate(self.devices)]

    def decode(self, pkt, frame):
        """pkt: bytes or PinnedPacket; frame: PinnedFrame (or DeviceFrame) every context writes its tiles into."""
        pending = [(c, c.submit(pkt, frame)) for c in self.ctx]
        for c, slot in pending:
            c.wait(slot)

    def close(self):
        for c in self.ctx:
            c.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


class FrameQueue:
    """Frames of a sequence handed out in runs of `chunk` from one counter shared by all ranks: whoever is free takes
    the next run, so a GPU behind a faster PCIe path decodes more frames than one behind a contended switch.  The
    counter lives in a torch.distributed store (TCPStore / FileStore / HashStore: `add` is atomic there); no
    collective, nothing on the data path.  One queue per pass over the sequence: `name` must be new each time."""

    def __init__(self, store, name: str, total: int, chunk: int = 4):
        if total < 0 or chunk < 1:
            raise ValueError("bad total/chunk")
        self.store, self.key, self.total, self.chunk = store, "framequeue/" + name, int(total), int(chunk)

    def take(self):
        """-> range of frame indices to decode next, or None when the sequence is handed out."""
        if self.store is None:                         # single process: a local counter does
            start = getattr(self, "_next", 0)
            self._next = start + self.chunk
        else:
            start = int(self.store.add(self.key, self.chunk)) - self.chunk
        if start >= self.total:
            return None
        return range(start, min(start + self.chunk, self.total))

    def __iter__(self):
        while True:
            r = self.take()
            if r is None:
                return
            yield from r


def default_store():
    """The store torch.distributed was initialised with (None without a process group)."""
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized():
        return None
    from torch.distributed import distributed_c10d
    return distributed_c10d._get_default_store()


def reduce_stats(elapsed_s: float, units_done: int, dist=None, device=None):
    """-> (max elapsed over ranks, total units over ranks).  dist: an initialised torch.distributed module
    (nccl on GPUs, gloo in the CPU tests) or None for a single process."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(elapsed_s), int(units_done)
    import torch
    t = torch.tensor([float(elapsed_s)], dtype=torch.float64, device=device)
    n = torch.tensor([int(units_done)], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(n, op=dist.ReduceOp.SUM)
    return float(t.item()), int(n.item())
