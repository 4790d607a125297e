"""Data-parallel rule of the training step on CPU (gloo, world_size 2): lgteun_b200.train.allreduce_gradients over a flat
gradient + the 1/world factor reproduces the gradient of the GLOBAL batch (each rank's nn.L1Loss is a mean over its own
half), and the Adam update that follows leaves both ranks with identical parameters.  Gradients come from the oracle."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import load_weights


def _flatten(grads, keys):
    return torch.cat([grads[k].reshape(-1) for k in keys])


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(2)
        from oracle import lgteun_oracle as O
        from lgteun_b200.train import allreduce_gradients
        sd = load_weights(4)
        gen = torch.Generator().manual_seed(21)
        ms, pan, gt = torch.rand(2, 4, 4, 4, generator=gen), torch.rand(2, 1, 16, 16, generator=gen), torch.rand(2, 4, 16, 16, generator=gen)
        _, _, g_all = O.train_step_grads(sd, ms, pan, gt)
        keys = [k for k, g in g_all.items() if g is not None]
        _, _, g_loc = O.train_step_grads(sd, ms[rank:rank + 1], pan[rank:rank + 1], gt[rank:rank + 1])
        flat = _flatten(g_loc, keys)
        scale = allreduce_gradients(flat)
        ref = _flatten(g_all, keys)
        err = ((flat * scale) - ref).abs().max().item() / ref.abs().max().item()
        p0 = _flatten(sd, keys)
        p1, _, _ = O.adam_step(p0, flat * scale, torch.zeros_like(p0), torch.zeros_like(p0), 1, 1.5e-3)
        others = [torch.zeros_like(p1) for _ in range(world)]
        dist.all_gather(others, p1)
        q.put((rank, scale, err, all(torch.equal(o, p1) for o in others)))
    finally:
        dist.destroy_process_group()


def test_two_rank_gradient_average_matches_global_batch():
    world, port = 2, 31500 + os.getpid() % 2000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, scale, err, same in res:
        assert scale == 0.5 and err <= 1e-5 and same


def test_single_process_has_no_collective():
    from lgteun_b200.train import allreduce_gradients
    g = torch.ones(8)
    assert allreduce_gradients(g) == 1.0 and torch.equal(g, torch.ones(8))
