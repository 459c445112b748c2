"""N>1 host logic on CPU: world_size-2 gloo processes agree on a partition of the batch (no image lost
or duplicated, balanced bytes), parse their shard with the host marker walk, and reduce timings with max."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import synth
from jpeglibrary_b200.sharding import gather_counts, max_over_ranks, shard_by_size


def test_shard_by_size_partitions_and_balances():
    sizes = [5, 9, 1, 7, 3, 8, 2, 6, 4, 10]
    shards = shard_by_size(sizes, 3)
    flat = sorted(i for s in shards for i in s)
    assert flat == list(range(len(sizes)))
    loads = [sum(sizes[i] for i in s) for s in shards]
    assert max(loads) - min(loads) <= max(sizes)
    assert shard_by_size([], 4) == [[], [], [], []]
    assert shard_by_size([3, 1], 4) == [[0], [1], [], []]


def _worker(rank, world, port, blobs, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import jpeglibrary_b200 as J
        sizes = [len(b) for b in blobs]
        mine = shard_by_size(sizes, world)[rank]
        pixels = 0
        for i in mine:  # host side of the hot path only: no GPU in this test
            d = J.Parsed(blobs[i]).desc
            pixels += d.width * d.height
        counts = gather_counts(len(mine))
        slowest = max_over_ranks(10.0 + rank)
        total_px = torch.tensor([pixels], dtype=torch.int64)
        dist.all_reduce(total_px)
        q.put((rank, mine, counts, slowest, int(total_px.item())))
    finally:
        dist.destroy_process_group()


def test_two_rank_scatter_over_gloo():
    blobs = [synth.synth_jpeg(i, 64 + 16 * (i % 3), 48, restart_rows=1) for i in range(7)]
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, blobs, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, m0, c0, t0, px0), (r1, m1, c1, t1, px1) = res
    assert sorted(m0 + m1) == list(range(7)) and not set(m0) & set(m1)
    assert c0 == c1 == [len(m0), len(m1)]
    assert t0 == t1 == 11.0                      # max over ranks
    assert px0 == px1 == sum((64 + 16 * (i % 3)) * 48 for i in range(7))
