"""N>1 host logic on the CPU: two gloo ranks shard a small sequence round-robin, each runs the host side of
the path (codestream -> plan) on its shard, and the reduced bookkeeping must cover every frame exactly once
with the max-over-ranks time.  No GPU, no data-path collective."""
import os
import socket
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, nframes, q):
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tools"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    import ffmpeg_ht_b200 as fb
    from ffmpeg_ht_b200 import sharding
    import j2kgen
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = sharding.shard_indices(nframes, rank, world)
    blocks = 0
    for i in mine:
        planes = j2kgen.synth_image(96, 64, 3, 12, seed=100 + i)
        cs = j2kgen.encode(planes, 12, j2kgen.EncParams(levels=3, cblk=(32, 32)))
        plan = fb.plan(cs)
        blocks += len(plan["cblk"])
        assert plan["info"].width == 96
    dist.barrier()
    elapsed, total = sharding.reduce_stats(0.5 + rank, len(mine), dist)
    q.put((rank, mine, elapsed, total, blocks))
    dist.destroy_process_group()


def test_two_ranks_shard_a_sequence():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    nframes, world = 7, 2
    procs = [ctx.Process(target=_worker, args=(r, world, port, nframes, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    seen = sorted(i for _, mine, *_ in res for i in mine)
    assert seen == list(range(nframes))                      # every frame exactly once
    for rank, mine, elapsed, total, blocks in res:
        assert elapsed == pytest.approx(0.5 + (world - 1))   # max over ranks
        assert total == nframes                              # sum over ranks
        assert blocks > 0


def _tile_worker(rank, world, port, q):
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tools"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import numpy as np
    import torch
    import torch.distributed as dist
    import ffmpeg_ht_b200 as fb
    from ffmpeg_ht_b200 import sharding
    import j2kgen
    dist.init_process_group("gloo", rank=rank, world_size=world)
    planes = j2kgen.synth_image(300, 200, 3, 12, seed=7)
    cs = j2kgen.encode(planes, 12, j2kgen.EncParams(levels=3, tile=(128, 96)))      # 3 x 3 tiles
    whole = fb.plan(cs)
    mine = fb.plan(cs, sharding.tile_flags(rank, world))
    live = (mine["cblk"]["flags"] & 0x80) == 0                # J2K_CBLK_SKIP marks the other ranks' blocks
    tiles = sorted({i // 3 for i, t in enumerate(mine["tcs"]) if int(t["transform"]) >= 0})
    # no data-path collective: only the bookkeeping is reduced -- every block and every tile has exactly one owner
    owners = torch.from_numpy(live.astype(np.int64))
    dist.all_reduce(owners, op=dist.ReduceOp.SUM)
    q.put((rank, tiles, int(live.sum()), len(whole["cblk"]), bool((owners == 1).all())))
    dist.destroy_process_group()


def test_two_ranks_shard_the_tiles_of_a_frame():
    """Tiles of one frame over two ranks (B200J2K_FLAG_TILES): each rank's plan keeps the tables of the whole frame
    but only its own tiles are live; together the ranks cover every code-block exactly once."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    world = 2
    procs = [ctx.Process(target=_tile_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == [0, 2, 4, 6, 8] and res[1][1] == [1, 3, 5, 7]
    assert res[0][2] + res[1][2] == res[0][3]
    assert all(r[4] for r in res)


def test_tile_flags():
    sys.path.insert(0, str(ROOT))
    from ffmpeg_ht_b200 import sharding
    assert sharding.tile_flags(0, 1) == 0
    assert sharding.tile_flags(3, 8) == (8 << 16) | (3 << 24)
    for bad in ((1, 1), (-1, 2), (0, 256)):
        with pytest.raises(ValueError):
            sharding.tile_flags(*bad)


def test_shard_indices_properties():
    sys.path.insert(0, str(ROOT))
    from ffmpeg_ht_b200 import sharding
    for world in (1, 2, 4, 8):
        for n in (0, 1, 4, 7, 960):
            allu = sorted(i for r in range(world) for i in sharding.shard_indices(n, r, world))
            assert allu == list(range(n))
            sizes = [len(sharding.shard_indices(n, r, world)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
            for u in range(min(n, 16)):
                assert u in sharding.shard_indices(n, sharding.owner(u, world), world)
    with pytest.raises(ValueError):
        sharding.shard_indices(4, 2, 2)


def _queue_worker(rank, world, port, nframes, q):
    sys.path.insert(0, str(ROOT))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import threading
    import time
    import torch.distributed as dist
    from ffmpeg_ht_b200 import sharding
    dist.init_process_group("gloo", rank=rank, world_size=world)
    fq = sharding.FrameQueue(sharding.default_store(), "pass0", nframes, chunk=3)
    mine, lock = [], threading.Lock()

    def work():                                   # two host threads per rank draw from the same queue
        for i in fq:
            time.sleep(0.002 * (1 + 3 * rank))    # rank 1 is four times slower: it must end up with fewer frames
            with lock:
                mine.append(i)

    dist.barrier()
    th = [threading.Thread(target=work) for _ in range(2)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert fq.take() is None
    q.put((rank, sorted(mine)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_draw_frames_from_one_queue():
    """FrameQueue: every frame exactly once over all ranks and threads, and the faster rank takes more."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    nframes, world = 200, 2
    procs = [ctx.Process(target=_queue_worker, args=(r, world, port, nframes, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res[0] + res[1]) == list(range(nframes))
    assert len(res[0]) > 1.5 * len(res[1])


def test_frame_queue_single_process():
    from ffmpeg_ht_b200 import sharding
    fq = sharding.FrameQueue(None, "x", 10, chunk=4)
    assert [list(fq.take()), list(fq.take()), list(fq.take()), fq.take()] == [[0, 1, 2, 3], [4, 5, 6, 7], [8, 9], None]
