"""Data-parallel plumbing of the update: samplers shard across ranks, ONE all-reduce per optimizer step.

The reference engine (allenact fork; one learner process per GPU, training/online/base.py:194-234) all-reduces every
parameter gradient separately (255 NCCL calls per step at this model).  Here the gradients of
all three towers live in one flat fp32 arena, so a step needs a single collective; the two scalars the
Lagrange-multiplier update needs (sum of finished-episode costs, number of finished episodes) ride in the arena's
tail, which makes lambda identical on every rank without a broadcast.  Device-agnostic (NCCL on GPUs; the CPU tests
run it over gloo)."""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist

TAIL = 64  # floats reserved after the gradients


def shard_samplers(num_samplers: int, world: int, rank: int) -> Tuple[int, int]:
    """[lo, hi) of the sampler (environment column) range owned by `rank`; every rank keeps whole trajectories."""
    if num_samplers % world != 0:
        raise ValueError(f"{num_samplers} samplers do not divide evenly over {world} ranks")
    per = num_samplers // world
    return rank * per, (rank + 1) * per


def allreduce_arena(comm: torch.Tensor, n_grad: int, cost_sum_cnt: Optional[torch.Tensor],
                    group: Optional[dist.ProcessGroup] = None) -> float:
    """comm = [gradients (n_grad) | tail].  Writes the local cost pair into the tail (when given), sums the whole
    buffer over the ranks in place and returns the factor that turns the summed gradients into the global-batch
    mean (equal shards: 1 / world)."""
    world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    if cost_sum_cnt is not None:  # 2 floats per cost channel
        assert cost_sum_cnt.numel() <= TAIL
        comm[n_grad:n_grad + cost_sum_cnt.numel()].copy_(cost_sum_cnt)
    else:
        comm[n_grad:n_grad + TAIL].zero_()
    if world > 1:
        dist.all_reduce(comm, op=dist.ReduceOp.SUM, group=group)
    return 1.0 / world
