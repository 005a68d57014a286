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


def _world() -> int:
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def reduce_scatter_shard(partition: torch.Tensor, shard: int) -> float:
    """Sum reduce-scatter of `partition` (world equal shards of `shard` floats, the Muon-labelled matrices grouped by owner
    rank - params.ParamStore(world=N)): afterwards THIS rank's shard holds the sum over ranks (in place, the other shards
    are left as they were).  Returns the scale (1/world) the optimizer applies.  NCCL reduce-scatters in place
    (recvbuff = sendbuff + rank * count); gloo has no reduce-scatter, so the CPU tests all-reduce the partition."""
    w = _world()
    if w == 1:
        return 1.0
    assert partition.numel() == w * shard
    if dist.get_backend() == "gloo":
        dist.all_reduce(partition)
        return 1.0 / w
    r = dist.get_rank()
    dist.reduce_scatter_tensor(partition[r * shard:(r + 1) * shard], partition)
    return 1.0 / w


def all_gather_shards(partition: torch.Tensor, shard: int) -> None:
    """All-gather in place: every rank contributes its own shard of `partition` and receives the others'."""
    w = _world()
    if w == 1:
        return
    assert partition.numel() == w * shard
    r = dist.get_rank()
    if dist.get_backend() == "gloo":
        parts = [torch.empty(shard, dtype=partition.dtype, device=partition.device) for _ in range(w)]
        dist.all_gather(parts, partition[r * shard:(r + 1) * shard].clone())
        for i, t in enumerate(parts):
            partition[i * shard:(i + 1) * shard].copy_(t)
        return
    dist.all_gather_into_tensor(partition, partition[r * shard:(r + 1) * shard])
