"""Env sharding across the GPUs of one box and the path's only collective.

Envs are independent: GPU g owns the contiguous block
[g*N/G, (g+1)*N/G) of global env indices, all buffers are local, and nothing crosses
GPUs on the data path.  The single exchange is the all-reduce of the 16-entry
episode-statistics vector (the reference's `_step_info` means and counts,
trifinger_env.py:554, :1067-1068, :1076, :1098-1099) when a caller wants job-wide
numbers: NCCL over NVLink on GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch

from . import _native as nat

MEAN_SLOTS = tuple(range(7)) + (nat.STAT_SUCCESSES, nat.STAT_REWARD)


def shard_range(global_num_envs: int, rank: int, world_size: int) -> Tuple[int, int]:
    """[first, last) global env indices owned by `rank`."""
    if global_num_envs % world_size != 0:
        raise ValueError("num_instances must be divisible by world_size")
    n = global_num_envs // world_size
    return rank * n, (rank + 1) * n


def stats_to_sums(step_stats: torch.Tensor, local_num_envs: int) -> torch.Tensor:
    """Shard statistics hold means for the mean-type slots and counts for the rest; sums add up
    across shards, means do not."""
    w = torch.ones_like(step_stats)
    w[list(MEAN_SLOTS)] = float(local_num_envs)
    return step_stats * w


def all_reduce_stats(step_stats: torch.Tensor, local_num_envs: int, global_num_envs: int,
                     group=None) -> torch.Tensor:
    """Job-wide statistics vector in the same layout (means / counts) from every rank's shard vector."""
    sums = stats_to_sums(step_stats.to(torch.float64), local_num_envs)
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        torch.distributed.all_reduce(sums, op=torch.distributed.ReduceOp.SUM, group=group)
    w = torch.ones_like(sums)
    w[list(MEAN_SLOTS)] = 1.0 / float(global_num_envs)
    return sums * w


def stats_to_info(stats: torch.Tensor, active_terms) -> Dict[str, float]:
    """The reference's `_step_info` keys from a statistics vector."""
    v = stats.detach().cpu().tolist()
    out = {}
    for i, name in enumerate(nat.TERM_NAMES):
        if name in active_terms:
            out[f"env/rewards/{name}"] = v[i]
    out["env/current_position_goal/count"] = v[nat.STAT_POSITION_GOAL]
    out["env/current_orientation_goal/count"] = v[nat.STAT_ORIENTATION_GOAL]
    out["env/average_consecutive_success"] = v[nat.STAT_SUCCESSES]
    return out
