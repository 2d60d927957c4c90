#!/usr/bin/env python
"""Small end-to-end case for compute-sanitizer (one tool per gpurun call):
   compute-sanitizer --tool memcheck python tools/sanitize_case.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gym_auv_b200 import lidar_config, scenarios as S  # noqa: E402
from gym_auv_b200.vec_env import AUVVecEnv  # noqa: E402

cfg = lidar_config()
cfg.episode.max_timesteps = 7  # force auto-resets
for scn, n in ((S.moving_obstacles(37, 17, 11, seed=1), 37), (S.land_scenarios(19, 96, 3, 3, seed=2, extent=1500.0), 19),
               (S.test_scenario2(), 1)):
    env = AUVVecEnv(scn, n, cfg, auto_reset=True, debug=True, sector_outputs=True)
    env.reset()
    g = torch.Generator(device="cuda").manual_seed(0)
    for t in range(30):
        a = torch.rand((n, 2), device="cuda", generator=g) * 2 - 1
        env.step(a)
    torch.cuda.synchronize()
    print(scn.name, env.episode_stats(reduce=False)["episodes"])
acts = np.zeros((37, 2), np.float32)
env = AUVVecEnv(S.moving_obstacles(37, 4, 4, seed=5), 37, cfg, auto_reset=True)
env.reset()
for _ in range(5):
    env.step_host(acts)
print("ok")
