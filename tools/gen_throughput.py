#!/usr/bin/env python
"""GPU-side scenario generation throughput (auv_generate_moving_obstacles + reset-cache refill)
and the cost of the ping-pong refresh during stepping.  Usage: python tools/gen_throughput.py [N]"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gym_auv_b200 import lidar_config, scenarios as S  # noqa: E402
from gym_auv_b200.vec_env import AUVVecEnv  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
cfg = lidar_config()
t0 = time.perf_counter()
scn = S.moving_obstacles_template(2 * N, 16, 16, seed=0, n_paths=1024)
t_host = time.perf_counter() - t0
env = AUVVecEnv(scn, N, cfg, auto_reset=True, chunks=4)
out = {"envs": N, "pool": 2 * N, "host_path_bank_s": t_host}
for rep in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    env.regenerate_scenarios(seed=1, epoch=rep + 1)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
out["regenerate_all_s"] = dt
out["scenarios_per_s_incl_reset_cache"] = 2 * N / dt
env.reset()
g = torch.Generator(device="cuda").manual_seed(0)
acts = [torch.rand((N, 2), device="cuda", generator=g) * 2 - 1 for _ in range(8)]
for i in range(20):
    env.step(acts[i % 8])
env.refresh_finished(seed=3)
torch.cuda.synchronize()
K, every = 200, 10
t0 = time.perf_counter()
for i in range(K):
    env.step(acts[i % 8])
torch.cuda.synchronize()
t_plain = time.perf_counter() - t0
refreshed = 0
t0 = time.perf_counter()
for i in range(K):
    env.step(acts[i % 8])
    if i % every == every - 1:
        refreshed += env.refresh_finished(seed=3)
torch.cuda.synchronize()
t_ref = time.perf_counter() - t0
out.update(steps=K, refresh_every=every, env_steps_per_s_plain=N * K / t_plain, env_steps_per_s_with_refresh=N * K / t_ref,
           scenarios_refreshed=refreshed, refreshed_per_s=refreshed / t_ref)
print(json.dumps(out))
