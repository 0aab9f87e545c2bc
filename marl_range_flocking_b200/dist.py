"""Multi-GPU plumbing: env instances shard over ranks, statistics are all-reduced.

Env instances are independent (no inter-env term anywhere in step/reward), so rank r
of W simply owns the contiguous global env range `shard_range(total, W, r)` on its own GPU and no
data-path collective exists. Philox streams are keyed by the GLOBAL env index (`env_offset`), so
results do not depend on W. The one collective is a SUM all-reduce of the 8-element int64 episode
statistics vector -- NCCL over NVLink for CUDA tensors, gloo for the CPU tests -- issued per
rollout, never per step.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.distributed as dist

from ._lib import STAT_NAMES


def shard_range(total_envs: int, world_size: int, rank: int) -> Tuple[int, int]:
    """(offset, count) of the env range owned by `rank`: contiguous, sizes differ by at most one."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} not in [0, {world_size})")
    base, rem = divmod(int(total_envs), int(world_size))
    count = base + (1 if rank < rem else 0)
    offset = rank * base + min(rank, rem)
    return offset, count


def make_sharded_env(variant: str, total_envs: int, agents: int, *args, rank: Optional[int] = None,
                     world_size: Optional[int] = None, **kwargs):
    """Build this rank's `VecEnv` shard of a `total_envs`-wide job (one process per GPU)."""
    from .vec_env import VecEnv

    if world_size is None:
        world_size = dist.get_world_size() if dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    offset, count = shard_range(total_envs, world_size, rank)
    return VecEnv(variant, count, agents, *args, env_offset=offset, **kwargs)


def allreduce_stats(stats: torch.Tensor, group=None) -> torch.Tensor:
    """SUM all-reduce of an int64 statistics vector across ranks (returns a new tensor).

    The episode-return slot is a two's-complement fixed-point sum, so integer SUM is exact and
    independent of the reduction order."""
    out = stats.clone()
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
    return out


def stats_dict(stats: torch.Tensor, agents: int) -> Dict[str, float]:
    s = [int(v) for v in stats.tolist()]
    n = max(s[0], 1)
    d = {name: s[i] for i, name in enumerate(STAT_NAMES)}
    d["mean_episode_length"] = s[1] / n
    d["mean_episode_return"] = s[2] / 4294967296.0 / agents / n
    return d


def global_stats(env, group=None) -> Dict[str, float]:
    """All-reduced episode statistics of a sharded job (one device->host read)."""
    return stats_dict(allreduce_stats(env.stats_tensor(), group), env.num_particles)
