"""Instance-level sharding across the GPUs of one box.

Instances are independent, so there is no data-path collective.  Two ways to hand them out:
  * ``shard_bounds`` / ``solve_sharded`` -- contiguous static blocks, one per rank (one process per GPU under
    torchrun); results (x, y int32 per instance) are gathered onto rank 0 with one ``gather`` over the process group
    (NCCL on GPUs, gloo in the CPU tests);
  * ``WorkQueue`` -- a shared queue of work units (batches of instances) that the ranks drain dynamically: solve
    times are data dependent (a sparse-family instance takes twice a uniform one), so with static blocks the slowest
    rank sets the time of the job.  The queue is one atomic counter in the process group's key-value store
    (``store.add``), i.e. control-plane traffic only; a rank claims its next unit when one of its lanes frees up.
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


class WorkQueue:
    """``total`` work units, claimed one at a time by whichever rank is ready (dynamic distribution
    instead of static equal splits).  ``store`` is a ``torch.distributed`` store (default: the one of the initialised
    default process group); without a process group the queue is local.  Every queue of a job needs its own ``name``."""

    def __init__(self, total: int, name: str = "b200lap_queue", store=None):
        import torch.distributed as dist
        self.total, self.key, self._local = int(total), f"{name}/next", 0
        self.store = store
        if self.store is None and dist.is_available() and dist.is_initialized():
            from torch.distributed import distributed_c10d
            self.store = distributed_c10d._get_default_store()
        self.claimed: List[int] = []

    def claim(self) -> Optional[int]:
        """The next unclaimed unit id, or None when the queue is drained."""
        if self.store is None:
            k = self._local
            self._local += 1
        else:
            k = int(self.store.add(self.key, 1)) - 1
        if k >= self.total:
            return None
        self.claimed.append(k)
        return k


def drain_queue(queue: "WorkQueue", launch: Callable[[int], "object"], in_flight: int = 1, wait: Optional[Callable[["object"], None]] = None):
    """Claims units until the queue is empty, keeping at most ``in_flight`` launched units outstanding on this rank:
    ``launch(unit)`` enqueues the unit's work asynchronously and returns a handle, ``wait(handle)`` blocks until it is
    done (default: ``handle.synchronize()``, e.g. a CUDA event).  The back-pressure is what makes the distribution
    dynamic -- a rank only claims when one of its lanes is free.  Returns the units this rank processed."""
    pending: List[Tuple[int, object]] = []
    done: List[int] = []
    wait = wait or (lambda h: h.synchronize())
    while True:
        if len(pending) >= max(1, in_flight):
            u, h = pending.pop(0)
            wait(h)
            done.append(u)
        unit = queue.claim()
        if unit is None:
            break
        pending.append((unit, launch(unit)))
    for u, h in pending:
        wait(h)
        done.append(u)
    return done


def bind_to_device_numa_node(device: int) -> Optional[int]:
    """Restrict this process to the CPUs of the NUMA node the GPU hangs off, so that pinned host buffers allocated and
    first-touched afterwards live in that node's memory: with one process per GPU and 2 GiB of binary64 matrices
    uploaded per step, host->device copies that cross the socket interconnect are what limits the end-to-end rate of an
    8-GPU box.  Returns the node, or None when the topology cannot be read (single node, containers without sysfs)."""
    import os
    try:
        import torch
        prop = torch.cuda.get_device_properties(device)
        bus = f"{prop.pci_domain_id:04x}:{prop.pci_bus_id:02x}:{prop.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = cpus & set(os.sched_getaffinity(0))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except Exception:  # noqa: BLE001
        return None
