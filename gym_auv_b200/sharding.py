"""Multi-GPU plumbing: the env batch shards trivially (every env is independent, SURVEY.md
section 8e); the only exchange is a sum of the episode-statistic accumulators."""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np


def shard_range(n_total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block of global env indices owned by `rank` (sizes differ by <= 1)."""
    base, rem = divmod(int(n_total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def reduce_stats(stats):
    """Sum the [AUV_N_STATS] accumulator over ranks (NCCL for CUDA tensors, gloo for CPU);
    no-op without an initialised process group."""
    import torch
    import torch.distributed as dist

    out = stats.clone()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(out, op=dist.ReduceOp.SUM)
    return out


def summarize_stats(v, t_step_size: float = 1.0) -> Dict[str, float]:
    """Per-episode means under the keys of ``env.history`` (environment.py:476-489)."""
    v = np.asarray(v, dtype=np.float64)
    n = max(v[0], 1.0)
    mean_r = v[1] / n
    return dict(
        episodes=float(v[0]),
        reward=float(mean_r),
        reward_std=float(np.sqrt(max(v[2] / n - mean_r * mean_r, 0.0))),
        progress=float(v[3] / n),
        collision=float(v[4] / n),
        reached_goal=float(v[5] / n),
        timesteps=float(v[6] / n),
        duration=float(v[6] / n * t_step_size),
        cross_track_error=float(v[7] / n),
        pathlength=float(v[8] / n),
        steps=float(v[9]),
    )
