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


def bind_to_gpu_numa(device_index: int) -> Tuple[int, ...]:
    """Pin this process to the CPU cores NVML reports as local to GPU `device_index` (the host
    side of its PCIe link), so that pinned staging buffers are first-touched on that NUMA node.
    With one process per GPU the host-buffer step is bound by the D2H link; cross-socket staging
    buffers halve it.  Returns the cores bound to (empty tuple: NVML / affinity unavailable)."""
    import os

    try:
        import pynvml

        pynvml.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = int(visible.split(",")[device_index]) if visible and visible.split(",")[device_index].isdigit() else device_index
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cores = [64 * w + b for w, m in enumerate(words) for b in range(64) if (int(m) >> b) & 1]
        allowed = sorted(set(cores) & os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
        return tuple(allowed)
    except Exception:
        return ()
