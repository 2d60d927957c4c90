#!/usr/bin/env python
"""bench.py -- env-steps/s of the gym-auv step path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this framework
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm (oracle port)

Workload (BASELINE.json configs[2]): MovingObstacles, 65536 envs per GPU x 180 rays x
32 obstacles (16 moving vessel pentagons + 16 static polygonised circles), paths shared
from a bank of 1024 random curves, random actions, auto-reset on (non-test mode).
One "step" = one full env.step() for every env of the batch (obstacle update, RKF45
vessel step, path projection / navigation, 180-ray LiDAR with reference culling,
reward, done, auto-reset).

Prints ONE JSON line (rank 0).  `value` = whole-job env-steps/s with actions resident
in HBM; `e2e` = the same through auv_step_host with host buffers (H2D actions, D2H
obs/reward/done inside the timed region).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "env-steps/sec (180-ray LiDAR, whole box)"
UNIT = "env-steps/s"
WORKLOAD = "MovingObstacles 65536 envs/GPU x 180 rays x 32 obstacles (16 moving + 16 static), 1024-path bank"

# algorithmic bytes per env-step of the fused step, SURVEY.md section 8(d), FP64 state build:
#   read : state 48 + action 8 + aux 40 + path tables ~ 3 x 64 + moving 16 x 40 + static 16 x 24
#   write: state 48 + aux 40 + moving 16 x 40 + obs 186 x 4 + reward/done/info 14
ALGO_BYTES_PER_ENV_STEP = (48 + 8 + 40 + 192 + 16 * 40 + 16 * 24) + (48 + 40 + 16 * 40 + 186 * 4 + 14)
FLOP_PER_SEG_TEST = 16
FLOP_PER_RAY = 60


def n_ranges(n, chunks):
    """env ranges auv_step_chunked cuts n envs into (range size = ceil(n/chunks) rounded up to 64)."""
    if chunks <= 1:
        return 1
    cs = -(-(-(-n // chunks)) // 64) * 64
    return -(-n // cs)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs", type=int, default=65536, help="envs per GPU")
    ap.add_argument("--n-moving", type=int, default=16)
    ap.add_argument("--n-static", type=int, default=16)
    ap.add_argument("--n-paths", type=int, default=1024)
    ap.add_argument("--rays", type=int, default=180)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--workload", default="moving", choices=["moving", "land", "pathfollow"],
                    help="moving = BASELINE config 3 (headline); land = config 4 shape (shared world of "
                         "--n-polygons static land polygons); pathfollow = config 2 (no LiDAR, PathFollowRewarder)")
    ap.add_argument("--n-polygons", type=int, default=512)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-numa-bind", action="store_true",
                    help="do not pin the process to the CPU cores local to its GPU (NVML affinity)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-sample-envs", type=int, default=8)
    ap.add_argument("--cpu-sample-steps", type=int, default=400)
    ap.add_argument("--chunks", type=int, default=4,
                    help="env ranges per step on the pipeline's own streams (auv_step_chunked); 1 = single stream")
    ap.add_argument("--chunk-streams", type=int, default=None)
    ap.add_argument("--host-chunks", type=int, default=4,
                    help="env ranges of the host-buffer step (auv_step_host_chunked): D2H of a range overlaps the next")
    ap.add_argument("--gpu-scenarios", action="store_true",
                    help="sample vessel starts and obstacles on the GPU (auv_generate_moving_obstacles) instead of "
                         "the host generator; same distributions, seconds instead of ~16 s of set-up")
    ap.add_argument("--preroll-steps", type=int, default=0,
                    help="untimed steps of the side pre-roll (0 = run for ~0.4 s); a fixed count makes profiler "
                         "captures of the steady state addressable by launch index")
    ap.add_argument("--scenario-cache", default=None,
                    help="pickle the generated scenario set here / reuse it (tuning sweeps; same seeds => same set)")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows = []
        self.proc = None
        self.gpu_index = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu_index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {
            "sm_mhz": float(np.median(sm)) if sm else None,
            "sm_max_mhz": float(max(mx)) if mx else None,
            "samples": len(sm),
            "reasons": sorted(reasons),
        }


def build_workload(args, rank):
    import pickle

    cache = getattr(args, "scenario_cache", None)
    if cache:
        cache = f"{cache}.{args.workload}.{args.envs}.{args.rays}.{args.n_moving}.{args.n_static}.{args.n_paths}.{args.seed}.{rank}"
        if os.path.exists(cache):
            with open(cache, "rb") as f:
                return pickle.load(f)
    out = _build_workload(args, rank)
    if cache:
        with open(cache, "wb") as f:
            pickle.dump(out, f, protocol=4)
    return out


def _build_workload(args, rank):
    from gym_auv_b200 import scenarios as S
    from gym_auv_b200.config import Config

    cfg = Config()
    if getattr(args, "workload", "moving") == "pathfollow":
        return cfg, S.path_follow_no_obstacles(args.envs, seed=args.seed + 1000 * rank, n_paths=args.n_paths)
    cfg.vessel.use_lidar = True
    per_sector = args.rays // cfg.vessel.n_sectors
    if per_sector * cfg.vessel.n_sectors != args.rays:
        cfg.vessel.n_sectors = 8
        per_sector = args.rays // 8
    cfg.vessel.n_sensors_per_sector = per_sector
    if getattr(args, "workload", "moving") == "land":
        scn = S.land_scenarios(args.envs, n_polygons=args.n_polygons, n_moving=args.n_moving, n_static=args.n_static,
                               seed=args.seed + 1000 * rank, n_paths=args.n_paths)
    elif getattr(args, "gpu_scenarios", False):
        scn = S.moving_obstacles_template(args.envs, args.n_moving, args.n_static, seed=args.seed + 1000 * rank,
                                          n_paths=args.n_paths)
    else:
        scn = S.moving_obstacles(args.envs, args.n_moving, args.n_static, seed=args.seed + 1000 * rank,
                                 n_paths=args.n_paths)
    return cfg, scn


# --------------------------------------------------------------------------------------
# CPU arm: the oracle port (restated reference; Shapely/GEOS not installable)
# --------------------------------------------------------------------------------------
def _cpu_worker(job):
    descs, cfgd, actions, warmup = job
    from oracle.sim import OracleEnv

    envs = [OracleEnv(d, cfgd, test_mode=False) for d in descs]
    T = actions.shape[0]
    t0 = None
    n = 0
    for t in range(T):
        if t == warmup:
            t0 = time.perf_counter()
        for i, env in enumerate(envs):
            _, _, done, _ = env.step(actions[t, i])
            if done:
                env.reset()
            if t >= warmup:
                n += 1
    return n, time.perf_counter() - t0


def cpu_baseline(args, cfg, scn, n_envs, steps, warmup, procs):
    """env-steps/s of the oracle port on `procs` host processes (each its own envs)."""
    import multiprocessing as mp
    from tests._parity import oracle_cfg

    cfgd = oracle_cfg(cfg)
    rng = np.random.RandomState(123)
    per = max(1, n_envs // procs)
    jobs = []
    for p in range(procs):
        ids = [(p * per + i) % scn.n_scenarios for i in range(per)]
        acts = rng.uniform([-1, -0.15], [1, 0.15], size=(steps + warmup, per, 2))
        jobs.append(([scn.describe(i) for i in ids], cfgd, acts, warmup))
    if procs == 1:
        res = [_cpu_worker(jobs[0])]
    else:
        with mp.get_context("fork").Pool(procs) as pool:
            res = pool.map(_cpu_worker, jobs)
    total = sum(r[0] for r in res)
    tmax = max(r[1] for r in res)
    return total / tmax, total, tmax


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    small = argparse.Namespace(**vars(args))
    small.envs = max(cores * 2, 16)
    small.n_paths = min(args.n_paths, small.envs)
    cfg, scn = build_workload(small, 0)
    per_step_envs = small.envs
    val, total, tmax = cpu_baseline(args, cfg, scn, per_step_envs, args.steps, args.warmup, cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tmax / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": f"{per_step_envs} envs per step on {cores} processes"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{total} env-steps ({per_step_envs} envs x {args.steps} steps), "
                                   "oracle/sim.py FP64 restatement (Shapely/GEOS not installable)"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from gym_auv_b200 import _lib
    from gym_auv_b200.vec_env import AUVVecEnv

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback on the product path)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    from gym_auv_b200.sharding import bind_to_gpu_numa

    numa_cores = bind_to_gpu_numa(local_rank) if not args.no_numa_bind else ()
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    _lib.load()

    cfg, scn = build_workload(args, rank)
    N, K, Wm = args.envs, args.steps, max(args.warmup, 3)
    R = cfg.vessel.n_sensors
    env = AUVVecEnv(scn, N, cfg, device=device, test_mode=False, auto_reset=True, env_offset=0,
                    chunks=args.chunks, chunk_streams=args.chunk_streams, host_chunks=args.host_chunks)
    scenario_gen = None
    if args.gpu_scenarios and args.workload == "moving":
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        env.regenerate_scenarios(seed=args.seed + 1000 * rank, epoch=1)
        g1.record()
        torch.cuda.synchronize()
        scenario_gen = {"where": "gpu", "scenarios": scn.n_scenarios, "ms_incl_reset_cache": g0.elapsed_time(g1)}
    gen = torch.Generator(device=device)
    gen.manual_seed(1234 + rank)
    lo = torch.tensor([-1.0, -0.15], device=device)
    hi = torch.tensor([1.0, 0.15], device=device)
    n_act = 16
    actions = [lo + (hi - lo) * torch.rand((N, 2), device=device, generator=gen) for _ in range(n_act)]
    env.reset()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up: W steps after reset(); the timed region starts from THIS state (the same
    #      episode phase the reference arm measures).
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for i in range(Wm):
        env.step(actions[i % n_act])
    barrier()
    # snapshot so that the pre-roll below can be undone and the counting pass can replay the K steps
    snap = {k: v.clone() for k, v in env._st.items()}
    # ---- pre-roll: ~0.4 s of the same steps.  (1) the clock sampler (nvidia-smi every 100 ms)
    #      gets several samples under exactly this load -- K steps of 0.2 ms are over before a second
    #      sample would arrive; (2) its last block measures the steady state thousands of steps into
    #      the run, when vessels have left their start areas and more obstacles are in range.
    t_pre = time.perf_counter()
    i, blk = Wm, 50
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    while True:
        last = (i - Wm + blk >= args.preroll_steps) if args.preroll_steps > 0 else (time.perf_counter() - t_pre >= 0.4)
        s0.record()
        for _ in range(blk):
            env.step(actions[i % n_act])
            i += 1
        s1.record()
        torch.cuda.synchronize()
        if last:
            break
    steady_ms = s0.elapsed_time(s1) / blk
    steady_records = float(env._scratch["rec_cnt"].sum().item()) / N
    import ctypes as C

    cfgp, rays, paths, pool, batch = env._refs()
    tm = env.lib.auv_timer_create(20)
    sps = C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
    for j in range(20):  # per-kernel split of the steady state (single stream)
        a = actions[(i + j) % n_act]
        _lib.check(env.lib.auv_step_timed(cfgp, rays, paths, pool, batch, C.c_void_p(a.data_ptr()), C.byref(env.out),
                                          sps, tm, j), "auv_step_timed")
    torch.cuda.synchronize()
    sk = np.zeros((20, 3), dtype=np.float32)
    for j in range(20):
        _lib.check(env.lib.auv_timer_read(tm, j, sk[j].ctypes.data_as(C.POINTER(C.c_float))), "auv_timer_read")
    env.lib.auv_timer_destroy(tm)
    cross_track = float(env._st["nav"][:, 2].abs().mean().item())
    t_s = torch.tensor([steady_ms], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_s, op=dist.ReduceOp.MAX)
    steady = {"value": world * N / (float(t_s.item()) * 1e-3), "unit": UNIT, "ms_per_step": float(t_s.item()),
              "steps_since_reset": i, "records_per_env_step": steady_records,
              "kernel_ms": {"k_vessel_nav": float(sk[:, 1].mean()), "k_lidar": float(sk[:, 2].mean())},
              "mean_abs_cross_track_m": cross_track,
              "note": "same kernels, measured over the last %d of %d untimed steps (auto-reset running)" % (blk, i)}
    for k, v in snap.items():
        env._st[k].copy_(v)
    barrier()
    stream = torch.cuda.current_stream(device)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cfgp, rays, paths, pool, batch = env._refs()
    import ctypes as C
    sp = C.c_void_p(stream.cuda_stream)

    def product_step(a):
        if env._pipe and env.chunks > 1:
            _lib.check(env.lib.auv_step_chunked(cfgp, rays, paths, pool, batch, C.c_void_p(a.data_ptr()),
                                                C.byref(env.out), sp, env._pipe, env.chunks), "auv_step_chunked")
        else:
            _lib.check(env.lib.auv_step(cfgp, rays, paths, pool, batch, C.c_void_p(a.data_ptr()), C.byref(env.out), sp),
                       "auv_step")

    # ---- timed region: K product steps (the chunked step forks from / joins into `stream`, so
    #      the two events on `stream` bracket all of its work)
    barrier()
    ev0.record(stream)
    for i in range(K):
        product_step(actions[(Wm + i) % n_act])
    ev1.record(stream)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = ev0.elapsed_time(ev1)
    t_local = torch.tensor([ms], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_local, op=dist.ReduceOp.MAX)
    ms_max = float(t_local.item())
    value = world * N * K / (ms_max * 1e-3)

    # ---- per-kernel attribution: replay the same K steps on ONE stream with CUDA events around
    #      each kernel (auv_step_timed); these are the launch durations the roofline uses
    for k, v in snap.items():
        env._st[k].copy_(v)
    timer = env.lib.auv_timer_create(K)
    assert timer, "auv_timer_create failed"
    torch.cuda.synchronize()
    es0, es1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    es0.record(stream)
    for i in range(K):
        a = actions[(Wm + i) % n_act]
        _lib.check(env.lib.auv_step_timed(cfgp, rays, paths, pool, batch, C.c_void_p(a.data_ptr()), C.byref(env.out),
                                          sp, timer, i), "auv_step_timed")
    es1.record(stream)
    torch.cuda.synchronize()
    serial_ms_per_step = es0.elapsed_time(es1) / K
    kms = np.zeros((K, 3), dtype=np.float32)
    for i in range(K):
        _lib.check(env.lib.auv_timer_read(timer, i, kms[i].ctypes.data_as(C.POINTER(C.c_float))), "auv_timer_read")
    env.lib.auv_timer_destroy(timer)
    kernel_ms = {"k_obstacle_update": float(kms[:, 0].mean()), "k_vessel_nav": float(kms[:, 1].mean()),
                 "k_lidar": float(kms[:, 2].mean())}
    obs_ms = kernel_ms["k_lidar"]
    kernel_ms["k_lidar_min_med_max"] = [float(kms[:, 2].min()), float(np.median(kms[:, 2])), float(kms[:, 2].max())]
    kernel_ms["k_vessel_nav_min_med_max"] = [float(kms[:, 1].min()), float(np.median(kms[:, 1])), float(kms[:, 1].max())]
    kernel_ms["single_stream_ms_per_step"] = serial_ms_per_step

    # ---- counting pass (untimed): replay the same K steps with the seg-test counter on
    for k, v in snap.items():
        env._st[k].copy_(v)
    seg = torch.zeros(1, dtype=torch.int64, device=device)
    env.out.seg_tests = seg.data_ptr()
    recs = torch.zeros(1, dtype=torch.int64, device=device)
    for i in range(K):
        env.step(actions[(Wm + i) % n_act])
        recs += env._scratch["rec_cnt"].sum()
    torch.cuda.synchronize()
    env.out.seg_tests = None
    seg_tests_per_step = float(seg.item()) / K
    records_per_step = float(recs.item()) / K  # obstacle records k_vessel_nav hands to k_lidar
    dones_per_step = float(env._out["stats"][0].item()) / max(env.total_steps, 1)

    # ---- FP32 FMA peak probe (roofline denominator for the LiDAR kernel), measured live
    sink = torch.zeros(1, device=device)
    fl = C.c_double(0)
    blocks = 148 * 8
    for _ in range(2):
        env.lib.auv_fma_probe(C.c_void_p(sink.data_ptr()), blocks, 256, 1 << 15, C.c_void_p(stream.cuda_stream), C.byref(fl))
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record(stream)
    env.lib.auv_fma_probe(C.c_void_p(sink.data_ptr()), blocks, 256, 1 << 15, C.c_void_p(stream.cuda_stream), C.byref(fl))
    p1.record(stream)
    torch.cuda.synchronize()
    fp32_peak_tflops = fl.value / (p0.elapsed_time(p1) * 1e-3) / 1e12

    # ---- e2e through the host-buffer C ABI call
    e2e = None
    if not args.no_e2e:
        acts_np = [a.cpu().numpy() for a in actions[:4]]
        for i in range(2):
            env.step_host(acts_np[i % 4])
        barrier()
        ke = max(3, min(K, 20))
        t0 = time.perf_counter()
        for i in range(ke):
            env.step_host(acts_np[i % 4])
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t_e = torch.tensor([dt], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
        # what the link alone allows: the same D2H copies (obs, reward, done) with no kernels
        pin = env._pinned
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        c0.record(stream)
        for _ in range(10):
            pin["obs"].copy_(env._out["obs"], non_blocking=True)
            pin["reward"].copy_(env._out["reward"], non_blocking=True)
            pin["done"].copy_(env._out["done"], non_blocking=True)
        c1.record(stream)
        torch.cuda.synchronize()
        d2h_ms = c0.elapsed_time(c1) / 10
        e2e_ms = 1e3 * float(t_e.item()) / ke
        e2e = {"value": world * N * ke / float(t_e.item()), "unit": UNIT,
               "h2d_bytes_per_step": env.h2d_bytes_per_step * world,
               "d2h_bytes_per_step": env.d2h_bytes_per_step * world, "steps": ke, "ms_per_step": e2e_ms,
               "host_chunks": env.host_chunks,
               "pcie": {"d2h_only_ms_per_step": d2h_ms, "d2h_gbs": env.d2h_bytes_per_step / (d2h_ms * 1e-3) / 1e9,
                        "frac_of_link_bound": d2h_ms / e2e_ms,
                        "note": "rank-0 link; the e2e step is bound by the D2H of the observations "
                                "(pinned host memory); frac = D2H-only time / e2e step time"}}

    # ---- e2e, asynchronous interface: two env groups of N/2 stepped alternately with
    #      step_async / step_wait (stable-baselines' VecEnv interface) so that one group's
    #      observations travel while the other group is computed.  Same env count, same host
    #      buffers, every step's actions come from host memory and every observation lands there.
    if e2e is not None and N >= 128:
        half = N // 2
        groups = env.groups(2)  # two envs of N/2 sharing this env's device tables
        for g in groups:
            g.reset()
        ah = [[a[:half].copy() for a in acts_np], [a[half:2 * half].copy() for a in acts_np]]
        for g, a in zip(groups, ah):
            g.step_async(a[0])
        for i in range(3):  # warm-up (graph capture happens here)
            for g, a in zip(groups, ah):
                g.step_wait()
                g.step_async(a[i % 4])
        for g in groups:
            g.step_wait()
        barrier()
        t0 = time.perf_counter()
        for g, a in zip(groups, ah):
            g.step_async(a[0])
        for i in range(ke):
            for g, a in zip(groups, ah):
                g.step_wait()
                g.step_async(a[(i + 1) % 4])
        for g in groups:
            g.step_wait()
        dt = time.perf_counter() - t0
        t_a = torch.tensor([dt], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t_a, op=dist.ReduceOp.MAX)
        a_ms = 1e3 * float(t_a.item()) / (ke + 1)
        sync_part = {"value": e2e["value"], "ms_per_step": e2e["ms_per_step"], "host_chunks": e2e["host_chunks"],
                     "call": "AUVVecEnv.step_host (one synchronous call per step)"}
        e2e.update({"value": world * 2 * half * (ke + 1) / float(t_a.item()), "ms_per_step": a_ms, "steps": ke + 1,
                    "mode": "AUVVecEnv.step_async / step_wait, two groups of N/2 envs stepped alternately",
                    "host_chunks": groups[0].host_chunks, "sync": sync_part})
        e2e["pcie"]["frac_of_link_bound"] = e2e["pcie"]["d2h_only_ms_per_step"] / a_ms
        for g in groups:
            g.close()
        del groups

    stats = env.episode_stats(reduce=True)  # the only collective on the path (NCCL all-reduce)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    print(f"[bench] value={value:.4g} ms/step={ms_max / K:.4f} kernel_ms={kernel_ms} e2e={e2e}", file=sys.stderr)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    flops_per_launch = FLOP_PER_SEG_TEST * seg_tests_per_step + FLOP_PER_RAY * (R if cfg.vessel.use_lidar else 0) * N
    achieved_tflops = flops_per_launch / (obs_ms * 1e-3) / 1e12
    # Algorithmic bytes per launch of the two step kernels.
    #   k_lidar per env: reads the navigation record 192 and its obstacle records (80 B each),
    #     writes the closeness part of obs 4*(obs_dim-6), reward/done/info 15, counters 20.
    #   k_vessel_nav per env (SURVEY 8d accounting, FP64 state): reads state 48 + action 8 +
    #     counters 40 + path tables ~192 + moving Km x (40 state + 40 pool) + static Ks x 24, writes
    #     state 48 + counters 8 + moving Km x 40 + navigation record 192 + obs[0..5] 24 + its
    #     obstacle records (80 B each).
    Km, Ks = env.k_moving, env.k_static
    kbytes = {
        "k_lidar": N * (192 + 4 * (env.obs_dim - 6) + 15 + 20) + 80.0 * records_per_step,
        "k_vessel_nav": N * (48 + 8 + 40 + 192 + Km * 80 + Ks * 24 + 48 + 8 + Km * 40 + 192 + 24) + 80.0 * records_per_step,
    }
    traffic = {}
    try:  # dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture (profiles/)
        tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        if args.workload == "moving" and N == tr["envs"] and R == tr["rays"]:
            traffic = {k: tr[k]["dram_bytes_per_launch"] for k in kbytes}
    except Exception:
        pass
    kernels = {}
    for k, nbytes in kbytes.items():
        gbs = nbytes / (kernel_ms[k] * 1e-3) / 1e9
        kernels[k] = {"ms_per_launch": kernel_ms[k], "algo_bytes_per_launch": nbytes, "achieved": gbs,
                      "frac": gbs / hbm_peak, "traffic": traffic.get(k)}
    dominant = max(kernels, key=lambda k: kernels[k]["ms_per_launch"])
    achieved_gbs = kernels[dominant]["achieved"]
    step_gbs = ALGO_BYTES_PER_ENV_STEP * N / (ms_max / K * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
        "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 ray casting on f64 state/culling", "data": "synthetic",
        "config": {"workload": WORKLOAD if args.workload == "moving" and N == 65536 and R == 180 else
                   f"{args.workload}: {N} envs/GPU x {R if cfg.vessel.use_lidar else 0} rays x {args.n_moving}+{args.n_static} obstacles"
                   + (f" + {args.n_polygons} shared land polygons" if args.workload == "land" else ""),
                   "envs_per_gpu": N, "rays": R, "obstacles": args.n_moving + args.n_static,
                   "paths": args.n_paths, "l2": "per-step working set (state+obstacles ~%.0f MB, path bank ~%.0f MB) exceeds the 126 MB L2; no explicit flush"
                   % (N * ALGO_BYTES_PER_ENV_STEP / 2e6, scn.bank.poly_xy.nbytes * 1.5 / 1e6 + scn.bank.coef.nbytes / 1e6),
                   "auto_reset": True, "dones_per_step": dones_per_step,
                   "chunks": env.chunks, "chunk_streams": getattr(env, "chunk_streams", 1),
                   "scenario_generation": scenario_gen or {"where": "host"},
                   "host_cores_bound": len(numa_cores)},
        "clocks": clocks,
        "steady_state": steady,
        "e2e": e2e,
        "gpu_launches": 2 * K * n_ranges(N, env.chunks),  # k_vessel_nav + k_lidar per env range
        "roofline": {
            "kernel": dominant, "bound": "hbm", "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s",
            "frac": achieved_gbs / hbm_peak, "traffic": kernels[dominant]["traffic"],
            "ms_per_launch": kernels[dominant]["ms_per_launch"],
            "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s",
            "algo_bytes_per_launch": kernels[dominant]["algo_bytes_per_launch"],
            "records_per_env_step": records_per_step / N,
            "kernels": kernels,
            "note": "Neither step kernel is HBM-bound: k_vessel_nav is bound by the latency of dependent FP64 chains "
                    "and table look-ups (ncu: 41 %% issue slots at 42 %% occupancy, dram 20 %% of peak), k_lidar by "
                    "instruction issue (68 %% issue slots, dram 6 %% of peak), see fp32_view; the HBM fractions are "
                    "low by construction.  Launch durations are CUDA-event times of a single-stream replay of the same "
                    "K steps (auv_step_timed); the headline `value` runs the same kernels as %d env ranges on %d "
                    "streams." % (env.chunks, getattr(env, "chunk_streams", 1)),
            "kernel_ms": kernel_ms,
            "fp32_view": {"kernel": "k_lidar", "achieved": achieved_tflops, "peak": fp32_peak_tflops, "unit": "TFLOP/s",
                          "frac": achieved_tflops / fp32_peak_tflops if fp32_peak_tflops else None,
                          "seg_tests_per_env_step": seg_tests_per_step / N,
                          "note": "algorithmic FLOPs = 16 x reference-semantics ray/segment tests + 60 x rays "
                                  "(SURVEY 8d); peak = FP32 FMA probe measured in this run"},
            "step_hbm_view": {"achieved": step_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": step_gbs / hbm_peak,
                              "algo_bytes_per_env_step": ALGO_BYTES_PER_ENV_STEP,
                              "note": "whole step (both kernels) algorithmic bytes / timed-region time per step"},
        },
        "episode_stats": stats,
    }
    if not args.no_cpu_baseline and world == 1:  # reported on rank 0 at N=1 only
        small = argparse.Namespace(**vars(args))
        small.envs = args.cpu_sample_envs
        small.n_paths = min(args.n_paths, small.envs)
        ccfg, cscn = build_workload(small, 0)
        v, total, tmax = cpu_baseline(args, ccfg, cscn, small.envs, args.cpu_sample_steps, 4, 1)
        line["cpu_baseline"] = {
            "value": v, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"{total} env-steps ({small.envs} envs x {args.cpu_sample_steps} steps) of the same workload "
                      "distribution, oracle/sim.py FP64 restatement (Shapely/GEOS not installable), %.1f s" % tmax,
        }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    # stdout carries exactly ONE JSON line: libraries that print to fd 1 (NCCL's version
    # banner, torchrun children) are diverted to stderr for the whole run
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w", buffering=1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
    sys.stdout.flush()


if __name__ == "__main__":
    main()
