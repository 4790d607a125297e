"""Multi-GPU host logic on CPU: batch sharding and the max-over-ranks timing reduce under gloo, world_size 2."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lgteun_b200.sharding import forward_sharded, max_over_ranks, shard_range, sum_over_ranks


@pytest.mark.parametrize("n,world", [(512, 1), (512, 2), (512, 8), (7, 4), (3, 8), (64, 3)])
def test_shard_range_partitions_the_batch(n, world):
    spans = [shard_range(n, world, r) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == n
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    sizes = [hi - lo for lo, hi in spans]
    assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        ms = torch.rand(5, 4, 4, 4, generator=g)
        pan = torch.rand(5, 1, 16, 16, generator=g)
        fake = lambda a, b: torch.nn.functional.interpolate(a, scale_factor=4) + b       # stand-in for the GPU module
        part = forward_sharded(fake, ms, pan, world, rank)
        parts = [None] * world
        dist.all_gather_object(parts, part)
        whole = torch.cat(parts)
        ok = torch.equal(whole, fake(ms, pan))
        t = max_over_ranks(10.0 + rank)
        s = sum_over_ranks(part.shape[0])
        q.put((rank, ok, t, s))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_sharded_forward():
    world, port = 2, 29500 + os.getpid() % 2000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, t, s in res:
        assert ok and t == 11.0 and s == 5.0


@pytest.mark.parametrize("n,chunk", [(512, 128), (512, 64), (100, 64), (64, 64), (65, 64), (1, 4), (0, 4), (130, 128), (7, 1)])
def test_host_pipeline_chunk_plan(n, chunk):
    """HostPipeline's chunk plan covers the batch exactly once, never exceeds the staging size, and keeps the two chunks
    whose copies are exposed (first upload, last download) short."""
    from lgteun_b200.hostio import plan_chunks
    plan = plan_chunks(n, chunk)
    assert sum(hi - lo for lo, hi in plan) == n
    assert all(0 < hi - lo <= chunk for lo, hi in plan)
    assert all(a[1] == b[0] for a, b in zip(plan, plan[1:]))
    if n > chunk:
        edge = max(chunk // 4, 1)
        assert plan[0] == (0, edge) and plan[-1][1] - plan[-1][0] <= edge
