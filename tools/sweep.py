#!/usr/bin/env python
"""BASELINE config 5: throughput sweep over env count x ray count, 32 obstacles (16 + 16).

    python tools/sweep.py                                   # one GPU
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/sweep.py --envs 65536 1048576

`--envs` is the env count PER GPU (the path shards trivially: weak scaling, no data-path collective); every
rank sweeps its own shard, the time of a point is the max over ranks, rank 0 writes JSON lines with the
whole-job throughput.  Paths are built and scenarios sampled on the GPU; `--preroll` untimed steps come first."""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gym_auv_b200 import scenarios as S  # noqa: E402
from gym_auv_b200.config import Config  # noqa: E402
from gym_auv_b200.vec_env import AUVVecEnv  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, nargs="+", default=[1024, 4096, 16384, 65536, 262144, 1048576, 4194304])
ap.add_argument("--rays", type=int, nargs="+", default=[64, 128, 180, 360])
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--preroll", type=int, default=100)
ap.add_argument("--out", default="gpurun_out/sweep.jsonl")
args = ap.parse_args()

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    import torch.distributed as dist

    dist.init_process_group("nccl", device_id=dev)
out = open(args.out, "w") if rank == 0 else None
for N in args.envs:
    t0 = time.time()
    scn = S.moving_obstacles_template(N, 16, 16, seed=rank, n_paths=min(1024, N), device_paths=True)
    gen = torch.Generator(device=dev).manual_seed(1 + rank)
    lo, hi = torch.tensor([-1.0, -0.15], device=dev), torch.tensor([1.0, 0.15], device=dev)
    acts = [lo + (hi - lo) * torch.rand((N, 2), device=dev, generator=gen) for _ in range(8)]
    for R in args.rays:
        cfg = Config()
        cfg.vessel.use_lidar = True
        cfg.vessel.n_sectors = 9 if R % 9 == 0 else 8
        cfg.vessel.n_sensors_per_sector = R // cfg.vessel.n_sectors
        env = AUVVecEnv(scn, N, cfg, device=dev, auto_reset=True, chunks=2 if N >= 16384 else 1)
        env.regenerate_scenarios(seed=rank, epoch=1)
        env.reset()
        for i in range(args.preroll):
            env.step(acts[i % 8])
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        setup_s = time.time() - t0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            env.step(acts[i % 8])
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        rec = dict(n_gpus=world, envs_per_gpu=N, envs=N * world, rays=R, obstacles=32, ms_per_step=ms,
                   env_steps_per_s=N * world / (ms * 1e-3), steps_before=args.preroll,
                   records_per_env=float(env._scratch["rec_cnt"].float().mean().item()), setup_s=setup_s)
        if rank == 0:
            print(json.dumps(rec), flush=True)
            out.write(json.dumps(rec) + "\n")
            out.flush()
        env.close()
        del env
        torch.cuda.empty_cache()
if world > 1:
    dist.destroy_process_group()
