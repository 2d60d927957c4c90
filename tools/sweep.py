#!/usr/bin/env python
"""BASELINE config 5: throughput sweep over env count x ray count on one GPU (run under
gpurun; multi-GPU points come from bench.py under torchrun).  Writes JSON lines."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gym_auv_b200 import scenarios as S  # noqa: E402
from gym_auv_b200.config import Config  # noqa: E402
from gym_auv_b200.vec_env import AUVVecEnv  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, nargs="+", default=[1024, 4096, 16384, 65536, 262144, 1048576, 4194304])
ap.add_argument("--rays", type=int, nargs="+", default=[64, 128, 180, 360])
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--out", default="gpurun_out/sweep.jsonl")
args = ap.parse_args()

import dataclasses  # noqa: E402

dev = torch.device("cuda:0")
t0 = time.time()
# path bank on the host (once), vessel starts and obstacles sampled on the GPU per env object
full = S.moving_obstacles_template(max(args.envs), 16, 16, seed=0, n_paths=min(1024, min(args.envs)))
bank_s = time.time() - t0
with open(args.out, "w") as f:
    for N in args.envs:
        t0 = time.time()
        scn = dataclasses.replace(
            full, path_id=full.path_id[:N], vessel_init=full.vessel_init[:N], mov_start=full.mov_start[:N],
            mov_width=full.mov_width[:N], mov_track=full.mov_track[:N], vel_table=full.vel_table[:N * 16],
            st_pos=full.st_pos[:N], st_radius=full.st_radius[:N], _bank=full.bank)
        gen_s = bank_s + time.time() - t0
        gen = torch.Generator(device=dev).manual_seed(1)
        lo, hi = torch.tensor([-1.0, -0.15], device=dev), torch.tensor([1.0, 0.15], device=dev)
        acts = [lo + (hi - lo) * torch.rand((N, 2), device=dev, generator=gen) for _ in range(8)]
        for R in args.rays:
            cfg = Config()
            cfg.vessel.use_lidar = True
            cfg.vessel.n_sectors = 9 if R % 9 == 0 else 8
            cfg.vessel.n_sensors_per_sector = R // cfg.vessel.n_sectors
            env = AUVVecEnv(scn, N, cfg, device=dev, auto_reset=True, chunks=4 if N >= 16384 else 1)
            env.regenerate_scenarios(seed=0, epoch=1)
            env.reset()
            for i in range(5):
                env.step(acts[i % 8])
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(args.steps):
                env.step(acts[i % 8])
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.steps
            rec = dict(envs=N, rays=R, obstacles=32, ms_per_step=ms, env_steps_per_s=N / (ms * 1e-3), scenario_gen_s=gen_s)
            print(json.dumps(rec), flush=True)
            f.write(json.dumps(rec) + "\n")
            del env
            torch.cuda.empty_cache()
