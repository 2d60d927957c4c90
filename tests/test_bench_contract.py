"""bench.py contract checks that run without a GPU: the flags the driver passes parse, and the
reference arm (CPU oracle port) prints exactly one JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1",
                          "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, res.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "env-steps/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["metric"].startswith("env-steps/sec")


def test_reference_arm_is_silent_on_other_ranks():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert res.returncode == 0 and res.stdout.strip() == ""


def test_gpu_arm_fails_loudly_without_cuda():
    import pytest

    torch = pytest.importorskip("torch")
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert res.returncode != 0 and "no CPU fallback" in (res.stderr + res.stdout)


def test_roofline_block_names_the_longest_single_launch():
    """bench.roofline_block: three step kernels, each with its own CUDA-event interval; the headline roofline is that of
    the single launch with the longest duration (k_lidar -> FP32 roof, a navigation kernel -> HBM roof); a timer that
    brackets the navigation pair as one interval keeps the pair whole.  Pure host logic: no GPU, no oracle."""
    import numpy as np

    import bench

    kms = np.array([[0.052, 0.037, 0.075], [0.054, 0.039, 0.077]], dtype=np.float32)
    t = bench.kernel_times(kms)
    assert abs(t["k_vessel_nav"] - 0.053) < 1e-6 and abs(t["k_nav_cull"] - 0.038) < 1e-6 and abs(t["nav_pair"] - 0.091) < 1e-6
    assert t["k_lidar_min_med_max"][0] <= t["k_lidar"] <= t["k_lidar_min_med_max"][2]
    kw = dict(N=65536, R=180, obs_dim=186, step_ms=0.164, records_per_step=1.4 * 65536, seg_tests_per_step=1200.0 * 65536,
              k_moving=16, k_static=16, refresh_interval=25, table_tracks=False, fp32_peak_tflops=64.0, workload="moving")
    r = bench.roofline_block(kernel_ms=t, **kw)
    assert r["kernel"] == "k_lidar" and r["bound"] == "fp32" and r["unit"] == "TFLOP/s"
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12 and 0.2 < r["frac"] < 0.6
    assert set(r["kernels"]) == {"k_vessel_nav", "k_nav_cull", "k_lidar"}
    assert all(k["hbm_frac"] > 0 and k["algo_bytes_per_launch"] > 0 for k in r["kernels"].values())
    assert r["traffic"] is None or r["traffic"] > 0  # profiles/ncu_traffic.json: per launch, like `achieved`
    # the navigation kernel as the longest launch -> HBM roof
    slow = dict(t, k_vessel_nav=0.2, nav_pair=0.238)
    r = bench.roofline_block(kernel_ms=slow, **kw)
    assert r["kernel"] == "k_vessel_nav" and r["bound"] == "hbm" and r["unit"] == "GB/s" and 0 < r["frac"] < 1
    # an older timer: interval 0 ~ 0, interval 1 = the pair
    t2 = bench.kernel_times(np.array([[0.0004, 0.091, 0.075]], dtype=np.float32))
    assert t2["k_nav_cull"] is None and abs(t2["k_vessel_nav"] - 0.0914) < 1e-5
    r = bench.roofline_block(kernel_ms=t2, **kw)
    assert r["kernel"] == "k_vessel_nav + k_nav_cull" and set(r["kernels"]) == {"k_vessel_nav", "k_lidar"}
    assert r["traffic"] is None or r["traffic"] > 5e7
    import json

    json.dumps(r)  # serialisable
