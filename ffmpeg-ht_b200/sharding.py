"""Multi-GPU partitioning of the decode path: frames of a sequence (and tiles of one oversized frame) are
independent units -- the reference decodes them on independent frame-thread contexts
(libavcodec/pthread_frame.c:248-295) and independent execute2 tile jobs (libavcodec/jpeg2000dec.c:2894) --
so ranks shard them round-robin with NO data-path collective.  torch.distributed is used only for the
barrier around the timed region and for reducing the bookkeeping (max of elapsed time, sum of counts)."""
from __future__ import annotations


def shard_indices(n_units: int, rank: int, world: int) -> list[int]:
    """Unit i (frame, or tile of a frame) goes to rank i % world."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return list(range(rank, n_units, world))


def owner(unit: int, world: int) -> int:
    return unit % world


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
