"""Data-parallel plumbing (torch.distributed; NCCL on GPUs, gloo in the CPU tests).

The reference is single-device (train.py:312-315 is dead code).  Env sharding is exact w.r.t. the reference because
every loss term is a mean over equal-sized shards (train.py:204-212, values.py:63): averaging shard gradients equals
the gradient of the global mean; the global advantage normalisation (rollout.py:159-162) needs the global first and
second moments, i.e. a sum all-reduce of (sum adv, sum adv^2).
"""

from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_bounds(num_agents: int, rank: int, world: int) -> Tuple[int, int]:
    """Agents [lo, hi) owned by `rank` (contiguous, equal shards)."""
    if num_agents % world != 0:
        raise ValueError(f"num_agents={num_agents} is not divisible by world={world}")
    per = num_agents // world
    return rank * per, (rank + 1) * per


def allreduce_adv_moments(stats: torch.Tensor, local_count: int) -> float:
    """In-place sum all-reduce of stats = (sum adv, sum adv^2); returns the global element count."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(stats)
        return float(local_count * dist.get_world_size())
    return float(local_count)


def allreduce_gradients(flat_grad: torch.Tensor) -> float:
    """In-place sum all-reduce of the flat fp32 gradient; returns the scale (1/world) the optimizer applies."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(flat_grad)
        return 1.0 / dist.get_world_size()
    return 1.0


def broadcast_params(flat_params: torch.Tensor, src: int = 0) -> None:
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(flat_params, src)
