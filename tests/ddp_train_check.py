"""Launched by tests/test_gpu_train.py under torchrun (one process per GPU, NCCL): lgteun_b200.Trainer.step with per-rank batches.
Checks (a) the all-reduced flat gradient equals the sum of the ranks' local gradients, (b) every rank holds bit-identical
parameters after the step, (c) the loss is finite, (d) ranks constructed with different weights hold rank 0's after
Trainer() (broadcast), (e) the dropout seed differs per rank.  Prints 'DDP_OK <world>' on rank 0."""
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rank, world = dist.get_rank(), dist.get_world_size()
    import lgteun_b200
    z = np.load(os.path.join(ROOT, "tests", "golden", "weights_b4.npz"))
    net = lgteun_b200.Pansharpening(SimpleNamespace(ms_chans=4), None, stage=2)
    sd = {k: torch.from_numpy(z[k]) for k in z.files}
    if rank > 0:       # ranks start from DIFFERENT weights: the Trainer has to broadcast rank 0's (as torch DDP does)
        sd = {k: v + 0.01 * rank for k, v in sd.items()}
    net.load_state_dict(sd)
    net = net.to(dev).train()
    tr = lgteun_b200.Trainer(net, lr=1.5e-3, dropout_p=0.1, seed=5)
    p0 = [torch.zeros_like(tr.flat.param) for _ in range(world)]
    dist.all_gather(p0, tr.flat.param)
    ok_bcast = all(torch.equal(p, p0[0]) for p in p0)
    seeds = [None] * world
    dist.all_gather_object(seeds, tr.dropout_seed())
    ok_seed = len(set(seeds)) == world          # every rank draws its own dropout masks
    gen = torch.Generator().manual_seed(40 + rank)
    ms, pan, gt = (torch.rand(2, 4, 16, 16, generator=gen).to(dev), torch.rand(2, 1, 64, 64, generator=gen).to(dev),
                   torch.rand(2, 4, 64, 64, generator=gen).to(dev))
    tr.forward_backward(ms, pan, gt)
    loc = tr.flat.grad.clone()
    parts = [torch.zeros_like(loc) for _ in range(world)]
    dist.all_gather(parts, loc)
    total = torch.stack(parts).sum(0)
    tr2_scale = lgteun_b200.train.allreduce_gradients(tr.flat.grad)
    ok_a = (tr.flat.grad - total).abs().max().item() <= 1e-6 * max(total.abs().max().item(), 1e-30) and tr2_scale == 1.0 / world
    loss = tr.step(ms, pan, gt)
    torch.cuda.synchronize()
    ps = [torch.zeros_like(tr.flat.param) for _ in range(world)]
    dist.all_gather(ps, tr.flat.param)
    ok_b = all(torch.equal(p, ps[0]) for p in ps)
    ok_c = bool(torch.isfinite(loss).all()) and bool(torch.isfinite(tr.flat.param).all())
    flags = torch.tensor([int(ok_a), int(ok_b), int(ok_c), int(ok_bcast), int(ok_seed)], device=dev)
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(("DDP_OK" if int(flags.min()) == 1 else f"DDP_FAIL {flags.tolist()}"), world, flush=True)
    dist.destroy_process_group()
    return 0 if int(flags.min()) == 1 else 1


if __name__ == "__main__":
    sys.exit(main())
