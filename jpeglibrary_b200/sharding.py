"""Multi-GPU plumbing: images are independent, so an N-GPU box is driven by host-side
scatter -- one process per GPU, no collective on the data path.  torch.distributed is used only to
agree on the partition and to take the max over ranks of the device-timed duration."""
from typing import List, Sequence


def shard_by_size(sizes: Sequence[int], world_size: int) -> List[List[int]]:
    """Greedy longest-processing-time partition of image indices by compressed size: every rank gets
    nearly the same number of entropy-coded bytes (the Huffman stage dominates)."""
    order = sorted(range(len(sizes)), key=lambda i: (-sizes[i], i))
    loads = [0] * world_size
    shards: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (loads[k], k))
        shards[r].append(i)
        loads[r] += sizes[i]
    for s in shards:
        s.sort()
    return shards


def my_shard(sizes: Sequence[int], rank: int, world_size: int) -> List[int]:
    return shard_by_size(sizes, world_size)[rank]


def max_over_ranks(value: float, device=None) -> float:
    """Whole-job time = the slowest rank's device time."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_counts(count: int, device=None) -> List[int]:
    """Units processed by every rank (rank 0 reports the aggregate)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [int(count)]
    t = torch.zeros(dist.get_world_size(), dtype=torch.int64, device=device)
    t[dist.get_rank()] = count
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [int(x) for x in t.tolist()]
