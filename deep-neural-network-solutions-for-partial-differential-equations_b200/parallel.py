"""Path sharding over the GPUs of one box: one process per GPU, parameters and Adam state
replicated, Brownian paths partitioned; the loss and its gradient are *sums* over paths, so one sum-allreduce of
[flat gradient | loss] per iteration makes every rank apply the identical Adam update.  The MC pricer shards
global path ids the same way and all-reduces three scalars at the end.  Backend: torch.distributed (NCCL over
NVLink on the GPU box; gloo in the CPU test-suite)."""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def is_distributed() -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def rank() -> int:
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def world_size() -> int:
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def shard_range(n: int, r: int, world: int) -> Tuple[int, int]:
    """Contiguous, near-even split of n units: the first n % world ranks get one extra (100 over 8 -> 13,13,13,13,12,...)."""
    if world < 1 or not (0 <= r < world):
        raise ValueError(f"bad rank/world {r}/{world}")
    base, extra = divmod(n, world)
    lo = r * base + min(r, extra)
    return lo, lo + base + (1 if r < extra else 0)


def allreduce_grads_and_loss(grad_flat: torch.Tensor, loss: torch.Tensor, group=None) -> None:
    """In-place sum over ranks of the flat gradient buffer and the scalar loss, as ONE collective (the loss rides
    in a packed copy so that small-M steps pay a single launch latency)."""
    if not is_distributed():
        return
    packed = torch.cat((grad_flat.reshape(-1), loss.reshape(-1)))
    dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
    n = grad_flat.numel()
    grad_flat.reshape(-1).copy_(packed[:n])
    loss.reshape(-1).copy_(packed[n:])


def allreduce_sums(values: torch.Tensor, group=None) -> torch.Tensor:
    """Sum a small fp64 vector (MC pricer: sum, sum of squares, count) over ranks."""
    if is_distributed():
        dist.all_reduce(values, op=dist.ReduceOp.SUM, group=group)
    return values
