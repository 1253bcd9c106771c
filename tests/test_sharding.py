"""CPU tests of the N>1 path: contiguous instance sharding + gather, world_size 2 over gloo."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from b200lap.sharding import WorkQueue, drain_queue, shard_bounds, solve_sharded


def test_shard_bounds_partition():
    for batch in (0, 1, 7, 64, 257):
        for world in (1, 2, 3, 8):
            blocks = [shard_bounds(batch, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == batch
            for (a, b), (c, d) in zip(blocks, blocks[1:]):
                assert b == c
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, batch, n, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        def solve_block(lo, hi):
            # stand-in for the per-rank device solve: instance k's "assignment" is a k-dependent permutation
            return torch.stack([torch.roll(torch.arange(n, dtype=torch.int32), k) for k in range(lo, hi)]) if hi > lo \
                else torch.zeros((0, n), dtype=torch.int32)
        res = solve_sharded(batch, solve_block)
        if rank == 0:
            full = torch.cat(res, dim=0)
            torch.save(full, out)
        else:
            assert res is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("batch", [5, 8])
def test_solve_sharded_world2_gloo(tmp_path, batch):
    n = 6
    out = str(tmp_path / "gathered.pt")
    mp.spawn(_worker, args=(2, _free_port(), batch, n, out), nprocs=2, join=True)
    full = torch.load(out)
    assert full.shape == (batch, n)
    for k in range(batch):
        assert torch.equal(full[k], torch.roll(torch.arange(n, dtype=torch.int32), k))


def _queue_worker(rank, world, port, total, out_dir):
    import time
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        q = WorkQueue(total, name="t")

        class Handle:                       # stand-in for a CUDA event: rank 1's units take four times as long
            def __init__(self, until): self.until = until
            def synchronize(self):
                while time.perf_counter() < self.until: time.sleep(0.001)

        def launch(unit):
            return Handle(time.perf_counter() + (0.02 if rank == 0 else 0.08))
        done = drain_queue(q, launch, in_flight=2)
        assert done == q.claimed
        torch.save(torch.tensor(done, dtype=torch.int64), os.path.join(out_dir, f"done{rank}.pt"))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_work_queue_world2_gloo(tmp_path):
    """Every unit is processed exactly once, and the faster rank ends up with more of them (dynamic balance)."""
    total = 40
    mp.spawn(_queue_worker, args=(2, _free_port(), total, str(tmp_path)), nprocs=2, join=True)
    a = torch.load(str(tmp_path / "done0.pt")).tolist()
    b = torch.load(str(tmp_path / "done1.pt")).tolist()
    assert sorted(a + b) == list(range(total))
    assert len(a) > len(b) + 4, (len(a), len(b))


def test_work_queue_without_a_process_group():
    q = WorkQueue(3)
    assert drain_queue(q, lambda u: u, in_flight=2, wait=lambda h: None) == [0, 1, 2]
    assert q.claim() is None
