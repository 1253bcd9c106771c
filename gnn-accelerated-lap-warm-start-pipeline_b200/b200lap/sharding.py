"""Instance-level sharding across the GPUs of one box.

Instances are independent, so the batch is split into contiguous blocks, one per rank (one process
per GPU under torchrun); there is no data-path collective.  Results (x, y int32 per instance) are
gathered onto rank 0 with one ``gather`` over the process group (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple


def shard_bounds(batch: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous block of instance indices owned by ``rank``: sizes differ by at most one."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    base, extra = divmod(max(batch, 0), world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def solve_sharded(batch: int, solve_block: Callable[[int, int], "object"], group=None, dst: int = 0):
    """Run ``solve_block(lo, hi)`` on this rank's block; it returns a tensor whose first dimension is
    ``hi - lo`` (e.g. the int32 row->column assignments).  Rank ``dst`` receives the list of every
    rank's block in rank order (concatenate for the whole batch); other ranks receive None."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return [solve_block(0, batch)]
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = shard_bounds(batch, world, rank)
    mine = solve_block(lo, hi).contiguous()
    # blocks differ in length by at most one: pad to the longest so gather() sees equal shapes
    longest = shard_bounds(batch, world, 0)[1] - shard_bounds(batch, world, 0)[0]
    padded = mine
    if mine.shape[0] < longest:
        pad = torch.zeros((longest - mine.shape[0],) + tuple(mine.shape[1:]), dtype=mine.dtype, device=mine.device)
        padded = torch.cat([mine, pad], dim=0)
    bucket = [torch.empty_like(padded) for _ in range(world)] if rank == dst else None
    dist.gather(padded, bucket, dst=dst, group=group)
    if rank != dst:
        return None
    out = []
    for r in range(world):
        a, b = shard_bounds(batch, world, r)
        out.append(bucket[r][: b - a])
    return out
