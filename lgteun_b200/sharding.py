"""Multi-GPU = batch sharding.  Image pairs are independent (LayerNorm per pixel, FFT and
attention per image, no BatchNorm), so rank r of G processes owns a contiguous slice of the batch and the
data path needs no collective; the only communication is the barrier / max-reduce used for timing and an
optional gather of results.  (The reference's own multi-GPU path is nn.DataParallel,
models/base/base_model.py:95-97.)"""
from typing import Tuple

import torch


def shard_range(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced [lo, hi) slice of n items for `rank` of `world` (first n % world ranks get one more)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def forward_sharded(model, ms: torch.Tensor, pan: torch.Tensor, world: int, rank: int) -> torch.Tensor:
    """Run `model` on this rank's slice of the global batch; returns the slice's output."""
    lo, hi = shard_range(ms.shape[0], world, rank)
    if hi == lo:
        return ms.new_empty((0, ms.shape[1], 4 * ms.shape[2], 4 * ms.shape[3]))
    return model(ms[lo:hi], pan[lo:hi])


def max_over_ranks(value: float, device=None) -> float:
    """Max of a per-rank scalar (elapsed ms) over the default process group; identity when not distributed."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device=None) -> float:
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
