"""GPU parity tests proper: the CUDA path (through the C ABI) against the CPU oracle on
identical injected scenarios and actions, and against golden vectors produced by the
reference's own dynamics files.  Run on the B200 box:  pytest tests -m gpu"""
import math
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from gym_auv_b200 import Config, effective_reference_config, lidar_config, scenarios as S  # noqa: E402
from tests._parity import compare, rollout_gpu, rollout_oracle  # noqa: E402

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def f32(a):
    return np.asarray(a, dtype=np.float32).astype(np.float64)


def random_actions(T, M, seed):
    return f32(np.random.RandomState(seed).uniform([-1, -0.15], [1, 0.15], size=(T, M, 2)))


@pytest.fixture(scope="module", autouse=True)
def _lib_loaded(built_lib):
    from gym_auv_b200 import _lib

    _lib.load()
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"


# ---------------------------------------------------------------------------------------
# dynamics kernel against the REFERENCE'S OWN odesolver/constants output (pinned)
# ---------------------------------------------------------------------------------------
def test_vessel_step_matches_reference_goldens():
    from gym_auv_b200.vec_env import AUVVecEnv

    g = np.load(os.path.join(GOLD, "reference_dynamics.npz"))
    worst = 0.0
    for dt in np.unique(g["dts"]):
        sel = np.nonzero(g["dts"] == dt)[0]
        base = S.empty_scenario()
        sets = []
        for i in sel:
            one = S.empty_scenario()
            one.vessel_init = g["inits"][i][None, :]
            sets.append(one)
        scn = S.concat(sets)
        cfg = Config()
        cfg.simulation.t_step_size = float(dt)
        env = AUVVecEnv(scn, len(sel), cfg, test_mode=True, auto_reset=False)
        env.reset()
        acts = torch.as_tensor(g["actions"][sel], dtype=torch.float32, device="cuda")  # [n, T, 2]
        for t in range(acts.shape[1]):
            env.vessel_step(acts[:, t].contiguous())
            got = env.state.cpu().numpy().T
            want = g["trajs"][sel, t + 1]
            worst = max(worst, float(np.abs(got - want).max()))
    assert worst <= 1e-11, worst


def test_vessel_step_nan_action_is_zero_action():
    from gym_auv_b200.vec_env import AUVVecEnv

    scn = S.concat([S.empty_scenario(), S.empty_scenario()])
    env = AUVVecEnv(scn, 2, Config(), test_mode=True, auto_reset=False)
    env.reset()
    a = torch.tensor([[float("nan"), 0.1], [0.0, 0.0]], device="cuda")
    for _ in range(3):
        env.vessel_step(a)
    st = env.state.cpu().numpy()
    assert np.array_equal(st[:, 0], st[:, 1])  # environment.py:314-315


# ---------------------------------------------------------------------------------------
# full step rollouts against the oracle
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("dt,seed", [(1.0, 3), (0.5, 4)])
def test_moving_obstacles_rollout(dt, seed):
    cfg = lidar_config()
    cfg.simulation.t_step_size = dt
    scn = S.moving_obstacles(8, 17, 11, seed=seed)
    actions = random_actions(64, 8, 100 + seed)
    ref = rollout_oracle(scn, cfg, actions)
    gpu, _ = rollout_gpu(scn, cfg, actions)
    rep = compare(ref, gpu, cfg, f"moving dt={dt}")
    assert rep["env_steps"] >= 300
    if ref["alive"].all():  # (the GPU env keeps stepping after done when auto_reset is off)
        assert gpu["seg_tests"] == int(ref["n_tests"].sum())  # reference-semantics ray/segment tests
    assert np.allclose(gpu["obs0"], np.array(ref["obs0"]), atol=2e-5)
    # moving-obstacle positions are integrated in FP64 on both sides
    for t in range(actions.shape[0]):
        for m in range(8):
            if ref["alive"][t, m]:
                assert np.abs(gpu["mov_pos"][t, m] - ref["mov_pos"][t][m]).max() < 1e-9


def test_moving_obstacles_close_quarters():
    """Vessels started right next to obstacles so that collisions, inside-polygon and
    sub-10 m ranges actually occur within the horizon."""
    cfg = lidar_config()
    scn = S.moving_obstacles(12, 17, 11, seed=21)
    rng = np.random.RandomState(5)
    for m in range(12):
        if m % 2 == 0:  # next to a static circle
            j = rng.randint(11)
            ang = rng.uniform(-np.pi, np.pi)
            r = scn.st_radius[m, j] + rng.uniform(1.5, 6.0)
            scn.vessel_init[m, :2] = scn.st_pos[m, j] + r * np.array([np.cos(ang), np.sin(ang)])
            scn.vessel_init[m, 2] = ang + np.pi + rng.uniform(-0.4, 0.4)
        else:  # in the way of a moving vessel
            j = rng.randint(17)
            v = scn.vel_table[scn.mov_track[m, j, 0]]
            scn.vessel_init[m, :2] = scn.mov_start[m, j] + v * rng.uniform(8, 20) + rng.uniform(-3, 3, size=2)
            scn.vessel_init[m, 2] = rng.uniform(-np.pi, np.pi)
    actions = random_actions(40, 12, 77)
    actions[..., 0] = np.abs(actions[..., 0])
    ref = rollout_oracle(scn, cfg, actions)
    gpu, _ = rollout_gpu(scn, cfg, actions)
    rep = compare(ref, gpu, cfg, "close quarters")
    assert ref["collision"].any(), "fixture should produce at least one collision"
    assert (ref["min_dist"][ref["alive"]] < 10).any()
    assert rep["windows_checked"] > 100


@pytest.mark.parametrize("n_sectors,per_sector", [(8, 8), (8, 16), (9, 40), (9, 7), (3, 11), (1, 33)])
def test_ray_counts_rollout(n_sectors, per_sector):
    """BASELINE config 5 ray counts (64 / 128 / 360) plus odd counts that are not a multiple of
    the warp size and give an odd observation width (scalar store path of k_lidar)."""
    cfg = lidar_config()
    cfg.vessel.n_sectors = n_sectors
    cfg.vessel.n_sensors_per_sector = per_sector
    assert cfg.vessel.n_sensors == n_sectors * per_sector
    scn = S.moving_obstacles(6, 17, 11, seed=40 + per_sector)
    actions = random_actions(30, 6, 200 + per_sector)
    ref = rollout_oracle(scn, cfg, actions)
    gpu, env = rollout_gpu(scn, cfg, actions)
    compare(ref, gpu, cfg, f"rays={n_sectors}x{per_sector}")
    assert env.obs_dim == 6 + n_sectors * per_sector
    if ref["alive"].all():
        assert gpu["seg_tests"] == int(ref["n_tests"].sum())


def test_dense_harbour_many_records_per_env():
    """Dozens of obstacles inside sensor range of every vessel: more records than one staging
    round of k_lidar holds (6 records / 192 vertices), 64-gons (r > 62.3 m) that exhaust the
    vertex budget on their own, vessels inside rings, overlapping windows."""
    cfg = lidar_config()
    M, Km, Ks = 6, 8, 40
    scn = S.moving_obstacles(M, Km, Ks, seed=77)
    rng = np.random.RandomState(9)
    for m in range(M):
        p0 = scn.vessel_init[m, :2]
        for j in range(Ks):
            ang = rng.uniform(-np.pi, np.pi)
            r = [2.0, 6.0, 20.0, 45.0, 70.0, 90.0][j % 6] * rng.uniform(0.9, 1.1)
            dist = r + rng.uniform(3.0, 120.0)
            if j == 0 and m % 2 == 0:
                dist = 0.5 * r  # own-ship starts inside a ring: rays measure the exit distance
                r = max(r, 30.0)
            scn.st_radius[m, j] = r
            scn.st_pos[m, j] = p0 + dist * np.array([np.cos(ang), np.sin(ang)])
    actions = random_actions(30, M, 78)
    ref = rollout_oracle(scn, cfg, actions)
    gpu, env = rollout_gpu(scn, cfg, actions)
    rep = compare(ref, gpu, cfg, "dense harbour")
    assert int(env._scratch["rec_cnt"].max().item()) > 12, "fixture should need several staging rounds"
    assert rep["windows_checked"] > 1000
    if ref["alive"].all():
        assert gpu["seg_tests"] == int(ref["n_tests"].sum())


def test_linear_closeness_transform():
    """sensor_log_transform=False (vessel.py:88-95: closeness = 1 - clip(d / range))."""
    cfg = lidar_config()
    cfg.vessel.sensor_log_transform = False
    scn = S.moving_obstacles(6, 17, 11, seed=5)
    actions = random_actions(25, 6, 6)
    ref = rollout_oracle(scn, cfg, actions)
    gpu, _ = rollout_gpu(scn, cfg, actions)
    compare(ref, gpu, cfg, "linear closeness")


DETERMINISTIC = ["TestScenario1-v0", "TestScenario2-v0", "TestScenario3-v0", "TestScenario4-v0", "TestHeadOn-v0",
                 "TestCrossing-v0", "TestCrossing1-v0", "EmptyScenario-v0", "DebugScenario-v0"]


@pytest.mark.parametrize("name", DETERMINISTIC)
def test_deterministic_scenarios(name):
    scn = S.SCENARIOS[name]()
    cfg = lidar_config()
    if name in ("EmptyScenario-v0", "DebugScenario-v0"):  # registered with DEBUG_CONFIG
        cfg.simulation.t_step_size = 0.5
        cfg.episode.min_goal_distance = 0.1
    T = 20 if name == "TestScenario2-v0" else 48
    a = random_actions(T, 1, 9)
    a[..., 0] = 0.9
    a[..., 1] *= 0.3
    a = f32(a)
    ref = rollout_oracle(scn, cfg, a)
    gpu, _ = rollout_gpu(scn, cfg, a)
    compare(ref, gpu, cfg, name)
    assert np.allclose(gpu["obs0"][0], ref["obs0"][0], atol=2e-5)


def test_reference_hierarchical_collision_detector_case_on_gpu():
    """tests/test_hierarchical_collision_detector.py of the reference, through the CUDA path."""
    from gym_auv_b200.vec_env import AUVVecEnv

    scn = S._single([[0, 100], [0, 100]], vessel_init=[5.0, -5.0, np.deg2rad(45)], static=[((0.0, -9.5), 1.5)])
    env = AUVVecEnv(scn, 1, lidar_config(), test_mode=True, auto_reset=False, debug=True)
    closeness = env.reset()[0, 6:].cpu().numpy()
    assert len(closeness) == 180
    assert not (0 < closeness[90] < 1)
    assert 0 < closeness[-1] < 1
    assert 0 < closeness[0] < 1
    assert env.get_attr("windows")[0, 0].tolist() == [-9, 5]
    assert np.nonzero(closeness)[0].tolist() == [0, 1, 2, 3, 172, 173, 174, 175, 176, 177, 178, 179]


@pytest.mark.parametrize("heading_deg", [45, 90, 170, -135, -170, 0, 10, -10, 179.5, -179.5])
def test_seam_windows_bit_exact(heading_deg):
    from oracle import sim as OS
    from gym_auv_b200.vec_env import AUVVecEnv

    psi = np.deg2rad(heading_deg)
    pos = np.array([5.0, -5.0])
    sets, want = [], []
    for off in (np.pi, np.pi - 0.2, np.pi + 0.2, 0.0, 1.0, -2.5):
        c = pos + 6.5 * np.array([np.cos(psi + off), np.sin(psi + off)])
        sets.append(S._single([[0, 100], [0, 100]], vessel_init=[pos[0], pos[1], psi], static=[(c, 1.5)]))
        v = OS.OracleVessel(dict(OS.DEFAULT_CFG), [pos[0], pos[1], psi])
        v.perceive([OS.OracleCircle(c, 1.5)])
        want.append((v.windows[0], v.dists.copy()))
    scn = S.concat(sets)
    env = AUVVecEnv(scn, len(sets), lidar_config(), test_mode=True, auto_reset=False, debug=True)
    env.reset()
    win = env.get_attr("windows").cpu().numpy()
    d = env.get_attr("lidar_dist").cpu().numpy()
    for i, (w, dist) in enumerate(want):
        assert tuple(win[i, 0]) == w
        assert np.allclose(d[i], dist, rtol=1e-4, atol=1e-4)


def test_exact_cull_mode_sees_obstacle_the_reference_misses():
    from gym_auv_b200.vec_env import AUVVecEnv

    psi = np.deg2rad(-135)
    pos = np.array([5.0, -5.0])
    c = pos - 6.5 * np.array([np.cos(psi), np.sin(psi)])
    scn = S._single([[0, 100], [0, 100]], vessel_init=[pos[0], pos[1], psi], static=[(c, 1.5)])
    ref_env = AUVVecEnv(scn, 1, lidar_config(), test_mode=True, auto_reset=False, debug=True)
    ref_env.reset()
    ex_env = AUVVecEnv(scn, 1, lidar_config(), test_mode=True, auto_reset=False, debug=True, cull_mode="exact")
    ex_env.reset()
    d_ref = ref_env.get_attr("lidar_dist")[0].cpu().numpy()
    d_ex = ex_env.get_attr("lidar_dist")[0].cpu().numpy()
    assert d_ref.min() == 150.0  # quirk #6: invisible to all rays
    assert d_ex.min() < 6.0
    assert np.all(d_ex <= d_ref)


# ---------------------------------------------------------------------------------------
# static land polygons (PolygonObstacle, BASELINE config 4 shape)
# ---------------------------------------------------------------------------------------
def test_land_polygons_rollout():
    from gym_auv_b200.polygons import random_land, star_polygon

    cfg = lidar_config()
    scn = S.moving_obstacles(8, 3, 2, seed=41)
    rng = np.random.RandomState(4)
    # a dense little world around the vessels: ~40 star polygons (non-convex) within 400 m
    polys = []
    for m in range(8):
        for _ in range(5):
            ang, dist = rng.uniform(-np.pi, np.pi), rng.uniform(25, 300)
            c = scn.vessel_init[m, :2] + dist * np.array([np.cos(ang), np.sin(ang)])
            polys.append(star_polygon(rng, c, min(rng.uniform(8, 60), dist - 8), int(rng.randint(8, 65))))
    # one vessel starts INSIDE a polygon (filled => range 0 => collision on the first step)
    polys.append(star_polygon(rng, scn.vessel_init[7, :2] + 1.0, 30.0, 12) * 1.0)
    scn.world_polygons = polys
    actions = random_actions(30, 8, 55)
    actions[..., 0] = np.abs(actions[..., 0])
    ref = rollout_oracle(scn, cfg, actions)
    gpu, env = rollout_gpu(scn, cfg, actions)
    rep = compare(ref, gpu, cfg, "land polygons")
    assert env.n_world == len(polys) and rep["windows_checked"] > 200
    assert ref["collision"][0, 7] and gpu["collision"][0, 7]  # inside the filled polygon
    assert gpu["dists"][0, 7].min() == 0.0
    assert (ref["min_dist"][ref["alive"]] < 150).mean() > 0.5
    if ref["alive"].all():
        assert gpu["seg_tests"] == int(ref["n_tests"].sum())


def test_land_polygons_512_world_sample_vs_oracle():
    """Config-4-sized shared world (512 polygons = 16 extra mask words per env)."""
    cfg = lidar_config()
    scn = S.land_scenarios(6, n_polygons=512, n_moving=2, n_static=2, seed=9, extent=2500.0)
    actions = random_actions(10, 6, 5)
    ref = rollout_oracle(scn, cfg, actions)
    gpu, env = rollout_gpu(scn, cfg, actions)
    compare(ref, gpu, cfg, "land 512")
    assert env.n_world == 512


# ---------------------------------------------------------------------------------------
# BASELINE config 2: no LiDAR, PathFollowRewarder
# ---------------------------------------------------------------------------------------
def test_path_follow_no_lidar_rollout():
    cfg = Config()  # use_lidar False (reference default)
    scn = S.path_follow_no_obstacles(16, seed=5, n_paths=4)
    actions = random_actions(80, 16, 31)
    actions[..., 0] = np.abs(actions[..., 0])
    ref = rollout_oracle(scn, cfg, actions)
    gpu, env = rollout_gpu(scn, cfg, actions)
    assert env.obs_dim == 6
    compare(ref, gpu, cfg, "pathfollow")


def test_colav_rewarder_without_lidar():
    cfg = Config()
    scn = S.empty_scenario()  # ColavRewarder, no LiDAR: every ray keeps sensor_range
    a = random_actions(30, 1, 2)
    ref = rollout_oracle(scn, cfg, a)
    gpu, _ = rollout_gpu(scn, cfg, a)
    compare(ref, gpu, cfg, "colav-nolidar")


def test_velocity_observation_channel_is_zero_and_shaped():
    from gym_auv_b200.vec_env import AUVVecEnv

    cfg = lidar_config(sensor_use_velocity_observations=True)
    scn = S.test_scenario3()
    env = AUVVecEnv(scn, 1, cfg, test_mode=True, auto_reset=False)
    obs = env.reset()
    assert obs.shape == (1, 6 + 540)  # reference tests/test_config.py intent
    assert (obs[0, 6:186] > 0).any() and (obs[0, 186:] == 0).all()  # sensor.py:159


# ---------------------------------------------------------------------------------------
# sector pooling (utils/sector_partitioning.py + LidarPreprocessor._feasibility_pooling)
# ---------------------------------------------------------------------------------------
def test_sector_pooling_outputs():
    from oracle.sim import sector_pool
    from gym_auv_b200.vec_env import AUVVecEnv

    cfg = lidar_config()
    scn = S.moving_obstacles(24, 17, 11, seed=33)
    rng = np.random.RandomState(8)
    for m in range(0, 24, 2):  # half of the envs start close to a static obstacle
        j = rng.randint(11)
        ang = rng.uniform(-np.pi, np.pi)
        scn.vessel_init[m, :2] = scn.st_pos[m, j] + (scn.st_radius[m, j] + rng.uniform(3, 40)) * np.array(
            [np.cos(ang), np.sin(ang)])
    env = AUVVecEnv(scn, 24, cfg, test_mode=True, auto_reset=False, debug=True, sector_outputs=True)
    env.reset()
    a = torch.as_tensor(random_actions(6, 24, 3), dtype=torch.float32, device="cuda")
    checked = 0
    for t in range(6):
        env.step(a[t])
        d = env.get_attr("lidar_dist").cpu().numpy()
        smin = env.get_attr("sector_min_dist").cpu().numpy()
        sfeas = env.get_attr("sector_feasible_dist").cpu().numpy()
        for m in range(24):
            want_min, want_feas = sector_pool(d[m], 9, cfg.vessel.vessel_width, cfg.vessel.feasibility_width_multiplier)
            assert np.array_equal(smin[m], want_min.astype(np.float32))  # pooling is exact on the same ranges
            assert np.array_equal(sfeas[m], want_feas.astype(np.float32))
            checked += int((want_min < 150).sum())
    assert checked > 50
    # sector index table is the reference's, bit-exact
    g = np.load(os.path.join(GOLD, "reference_dynamics.npz"))
    assert np.array_equal(env.sector_index, g["sectors_180"].astype(np.uint8))


# ---------------------------------------------------------------------------------------
# done / auto-reset semantics
# ---------------------------------------------------------------------------------------
def test_time_limit_done_and_auto_reset():
    from gym_auv_b200.vec_env import AUVVecEnv

    cfg = lidar_config()
    cfg.episode.max_timesteps = 6
    scn = S.moving_obstacles(4, 3, 2, seed=8)
    env = AUVVecEnv(scn, 4, cfg, test_mode=False, auto_reset=True, debug=True)
    obs0 = env.reset().clone()
    st0 = env.state.clone()
    a = torch.zeros((4, 2), device="cuda")
    a[:, 0] = 1.0
    dones, last_obs = [], None
    for t in range(6):
        obs, rew, done, info = env.step(a)
        dones.append(done.cpu().numpy().copy())
    # done exactly when the pre-increment t_step reaches max_timesteps - 1 (environment.py:380)
    assert [bool(d.all()) for d in dones] == [False, False, False, False, False, True]
    assert not any(d.any() for d in dones[:5])
    assert torch.equal(env.state, st0)  # env is back in its initial state...
    assert torch.allclose(obs, obs0)  # ...and the returned obs is the reset obs
    assert (env.get_attr("t_step") == 0).all() and (env.get_attr("episode") == 2).all()
    term = info["terminal_observation"]
    assert not torch.allclose(term, obs0)
    stats = env.episode_stats(reduce=False)
    assert stats["episodes"] == 4 and stats["timesteps"] == 6.0 and stats["collision"] == 0.0


def test_test_mode_ignores_time_and_reward_limits():
    from gym_auv_b200.vec_env import AUVVecEnv

    cfg = lidar_config()
    cfg.episode.max_timesteps = 3
    cfg.episode.min_cumulative_reward = -1.0
    env = AUVVecEnv(S.empty_scenario(), 1, cfg, test_mode=True, auto_reset=False)
    env.reset()
    a = torch.zeros((1, 2), device="cuda")
    for _ in range(6):
        _, _, done, _ = env.step(a)
        assert not bool(done.item())
    env2 = AUVVecEnv(S.empty_scenario(), 1, cfg, test_mode=False, auto_reset=False)
    env2.reset()
    _, r, done, _ = env2.step(a)
    assert bool(done.item()) and float(r.item()) < -1.0  # cumulative reward limit


def test_step_host_equals_step():
    from gym_auv_b200.vec_env import AUVVecEnv

    cfg = lidar_config()
    scn = S.moving_obstacles(33, 4, 4, seed=12)  # odd batch size on purpose
    e1 = AUVVecEnv(scn, 33, cfg, auto_reset=True)
    e2 = AUVVecEnv(scn, 33, cfg, auto_reset=True, host_transfer="compact")
    e3 = AUVVecEnv(scn, 33, cfg, auto_reset=True)  # default: delta transfer
    e1.reset(), e2.reset(), e3.reset()
    acts = random_actions(10, 33, 3).astype(np.float32)
    for t in range(10):
        o1, r1, d1, _ = e1.step(torch.as_tensor(acts[t], device="cuda"))
        o2, r2, d2 = e2.step_host(acts[t])
        o3, r3, d3 = e3.step_host(acts[t])
        assert np.array_equal(o1.cpu().numpy(), o2) and np.array_equal(r1.cpu().numpy(), r2)
        assert np.array_equal(d1.cpu().numpy(), d2)
        assert np.array_equal(o2, o3) and np.array_equal(r2, r3) and np.array_equal(d2, d3)
    assert e2.h2d_bytes_per_step == 33 * 8
    assert e3.host_transfer == "delta" and 33 * 5 < e3.d2h_bytes_per_step < 33 * (186 * 4 + 5)
    nz = int((o2[:, 6:] != 0).sum())  # compact transfer: 32 B head + 6 mask words + reward + done per env, 4 B per non-zero value
    assert e2.compact_host and e2.d2h_bytes_per_step == 33 * (32 + 24 + 5) + 4 * nz < 33 * (186 * 4 + 5)


@pytest.mark.parametrize("chunks,streams", [(2, 2), (3, 2), (5, 4), (8, 1)])
def test_chunked_step_is_bit_identical(chunks, streams):
    """auv_step_chunked / auv_step_host_chunked: env ranges on the pipeline's own streams give
    exactly the results of the single-stream step (envs are independent), including the
    auto-reset bookkeeping and the episode statistics."""
    from gym_auv_b200.vec_env import AUVVecEnv

    cfg = lidar_config()
    cfg.episode.max_timesteps = 9  # force auto-resets inside the rollout
    n = 333  # several ragged 64-env ranges
    scn = S.moving_obstacles(n, 5, 5, seed=21)
    e1 = AUVVecEnv(scn, n, cfg, auto_reset=True)
    e2 = AUVVecEnv(scn, n, cfg, auto_reset=True, chunks=chunks, chunk_streams=streams)
    e3 = AUVVecEnv(scn, n, cfg, auto_reset=True, chunks=1, host_chunks=chunks + 1)
    e1.reset(), e2.reset(), e3.reset()
    acts = random_actions(24, n, 4).astype(np.float32)
    for t in range(24):
        a = torch.as_tensor(acts[t], device="cuda")
        o1, r1, d1, i1 = e1.step(a)
        o2, r2, d2, i2 = e2.step(a)
        o3, r3, d3 = e3.step_host(acts[t])
        torch.cuda.synchronize()
        assert torch.equal(o1, o2) and torch.equal(r1, r2) and torch.equal(d1, d2)
        for k in i1:
            assert torch.equal(i1[k], i2[k]), k
        assert np.array_equal(o1.cpu().numpy(), o3) and np.array_equal(r1.cpu().numpy(), r3)
        assert np.array_equal(d1.cpu().numpy(), d3)
    for k in ("state", "scn_id", "t_step", "cum_reward", "nearby_mask", "obst_steps", "prev_seg", "env_pid"):
        assert torch.equal(e1._st[k], e2._st[k]) and torch.equal(e1._st[k], e3._st[k]), k
    assert torch.equal(e1.get_attr("mov_pos"), e2.get_attr("mov_pos")) and torch.equal(e1.get_attr("mov_pos"), e3.get_attr("mov_pos"))
    assert e3.lib.auv_pipeline_graph_state(e3._pipe) == 1, "host-buffer step should replay a CUDA graph"
    s1, s2 = e1.episode_stats(reduce=False), e2.episode_stats(reduce=False)
    assert s1["episodes"] == s2["episodes"] > 0
    assert abs(s1["reward"] - s2["reward"]) <= 1e-9 * abs(s1["reward"])  # atomic order differs
    e2.close(), e3.close()


def test_counting_variant_equals_product_variant():
    """debug=True launches k_lidar<COUNT=true> (+ lidar_dist / windows outputs); the product
    launch must produce the same observations, rewards and state."""
    from gym_auv_b200.vec_env import AUVVecEnv

    cfg = lidar_config()
    n = 97
    scn = S.moving_obstacles(n, 17, 11, seed=31)
    e1 = AUVVecEnv(scn, n, cfg, auto_reset=True, debug=True)
    e2 = AUVVecEnv(scn, n, cfg, auto_reset=True, debug=False)
    assert torch.equal(e1.reset(), e2.reset())
    acts = random_actions(40, n, 6).astype(np.float32)
    for t in range(40):
        a = torch.as_tensor(acts[t], device="cuda")
        o1, r1, d1, _ = e1.step(a)
        o2, r2, d2, _ = e2.step(a)
        assert torch.equal(o1, o2) and torch.equal(r1, r2) and torch.equal(d1, d2), t
    assert torch.equal(e1.state, e2.state)
    assert int(e1._out["seg_tests"].item()) > 0


def test_step_async_two_groups_equals_sync():
    """step_async / step_wait (stable-baselines' asynchronous VecEnv interface): two env groups
    stepped alternately on their own streams give exactly the synchronous results."""
    from gym_auv_b200.vec_env import AUVVecEnv

    cfg = lidar_config()
    cfg.episode.max_timesteps = 11
    n = 130
    scn = S.moving_obstacles(2 * n, 5, 5, seed=33)
    ref = AUVVecEnv(scn, 2 * n, cfg, auto_reset=True)
    ga, gb = ref.groups(2, host_chunks=3)  # share ref's device tables (pool, path bank, reset cache)
    assert ga.num_envs == gb.num_envs == n and gb.env_offset == n
    ref.reset(), ga.reset(), gb.reset()
    acts = random_actions(20, 2 * n, 8).astype(np.float32)
    ga.step_async(acts[0, :n])
    gb.step_async(acts[0, n:])
    with pytest.raises(RuntimeError):
        ga.step_async(acts[0, :n])
    for t in range(20):
        o, r, d, _ = ref.step(torch.as_tensor(acts[t], device="cuda"))
        o, r, d = o.cpu().numpy(), r.cpu().numpy(), d.cpu().numpy()
        oa, ra, da = (x.copy() for x in ga.step_wait())
        if t + 1 < 20:
            ga.step_async(acts[t + 1, :n])  # group A's next step is in flight while B is read
        ob, rb, db = (x.copy() for x in gb.step_wait())
        if t + 1 < 20:
            gb.step_async(acts[t + 1, n:])
        # a pool of 2n scenarios shared by two groups of n: env i of group B is env n + i of ref
        # only until its first reset (ref moves on by 2n scenarios, the groups by n)
        fresh_a = ga.get_attr("episode").cpu().numpy() == 0
        fresh_b = gb.get_attr("episode").cpu().numpy() == 0
        assert np.array_equal(oa[fresh_a], o[:n][fresh_a]) and np.array_equal(ra[fresh_a], r[:n][fresh_a])
        assert np.array_equal(ob[fresh_b], o[n:][fresh_b]) and np.array_equal(db[fresh_b], d[n:][fresh_b])
    with pytest.raises(RuntimeError):
        ga.step_wait()
    assert (~fresh_a).any(), "fixture should contain resets"


# ---------------------------------------------------------------------------------------
# GPU-side scenario generation
# ---------------------------------------------------------------------------------------
def _generated_env(M=512, n_paths=8, seed=3, **kw):
    from gym_auv_b200.vec_env import AUVVecEnv

    cfg = lidar_config()
    scn = S.moving_obstacles(M, 17, 11, seed=seed, n_paths=n_paths)
    env = AUVVecEnv(scn, kw.pop("num_envs", M), cfg, auto_reset=True, **kw)
    return cfg, scn, env


def test_gpu_generated_scenarios_obey_the_reference_rules_and_distributions():
    cfg, host, env = _generated_env(M=2048)
    env.regenerate_scenarios(seed=11, epoch=1)
    g = env.pull_scenarios()
    bank = host.bank
    L = np.array([bank.tables[p].length for p in g.path_id])
    # vessel start: within +-25 m of path(0) per axis, heading anywhere (movingobstacles.py:34-38)
    p0 = np.array([bank.tables[p](0.0) for p in g.path_id])
    assert np.abs(g.vessel_init[:, :2] - p0).max() <= 25.0 + 1e-9
    assert g.vessel_init[:, 2].min() >= -np.pi and g.vessel_init[:, 2].max() < np.pi
    assert len(np.unique(g.path_id)) == len(bank.tables)  # uniform choice over the bank
    goal = np.array([bank.tables[p](bank.tables[p].length) for p in g.path_id])
    w = cfg.vessel.vessel_width
    for pos, rad in ((g.mov_start, g.mov_width), (g.st_pos, g.st_radius)):
        # acceptance rule of helpers.generate_obstacle (utils/helpers.py:28-33)
        dv = np.linalg.norm(pos - g.vessel_init[:, None, :2], axis=2) - w - rad
        dg = np.linalg.norm(pos - goal[:, None, :], axis=2) - rad
        assert np.minimum(dv, dg).min() > 0
        assert rad.min() >= 1.0 and np.all(rad == np.round(rad))  # max(1, Poisson)
    speed = np.linalg.norm(g.vel_table, axis=1)
    assert speed.min() >= 1.0 and speed.max() <= 3.0 and abs(speed.mean() - 2.0) < 0.02
    # distributions against the host generator (NumPy streams, same path bank): means within 2 %
    for a, b in ((g.mov_width, host.mov_width), (g.st_radius, host.st_radius)):
        assert abs(a.mean() - b.mean()) < 0.02 * b.mean() and abs(a.std() - b.std()) < 0.05 * b.std()
    def spread(pos, sc):
        return np.linalg.norm(pos - sc.vessel_init[:, None, :2], axis=2)
    for a, b in ((spread(g.mov_start, g), spread(host.mov_start, host)), (spread(g.st_pos, g), spread(host.st_pos, host))):
        assert abs(np.median(a) - np.median(b)) < 0.05 * np.median(b)
    heading = np.arctan2(g.vel_table[:, 1], g.vel_table[:, 0])
    assert abs(np.mean(np.cos(heading))) < 0.02 and abs(np.mean(np.sin(heading))) < 0.02


def test_gpu_generation_is_deterministic_and_epoch_dependent():
    _, _, e1 = _generated_env(M=256)
    _, _, e2 = _generated_env(M=256)
    e1.regenerate_scenarios(seed=5, epoch=1)
    e2.regenerate_scenarios(seed=5, epoch=1)
    a, b = e1.pull_scenarios(), e2.pull_scenarios()
    for k in ("path_id", "vessel_init", "mov_start", "mov_width", "vel_table", "st_pos", "st_radius"):
        assert np.array_equal(getattr(a, k), getattr(b, k)), k
    assert torch.equal(e1._pool["reset_obs"], e2._pool["reset_obs"])
    ids = torch.arange(0, 256, 2)
    e2.regenerate_scenarios(ids, seed=5, epoch=2)  # only the even slots change
    c = e2.pull_scenarios()
    assert np.array_equal(c.st_pos[1::2], a.st_pos[1::2]) and not np.array_equal(c.st_pos[0::2], a.st_pos[0::2])


def test_gpu_generated_scenarios_replay_through_the_oracle():
    """Scenarios drawn on the GPU are pulled back and injected into the CPU oracle: the full
    step parity holds on them like on host-generated ones (incl. the cached first observation)."""
    from gym_auv_b200.vec_env import AUVVecEnv

    cfg, _, env = _generated_env(M=8, n_paths=3, seed=9)
    env.regenerate_scenarios(seed=21, epoch=4)
    scn = env.pull_scenarios()
    actions = random_actions(40, 8, 91)
    ref = rollout_oracle(scn, cfg, actions)
    gpu, _ = rollout_gpu(scn, cfg, actions)
    compare(ref, gpu, cfg, "gpu-generated")
    assert np.allclose(env._pool["reset_obs"].cpu().numpy(), np.array(ref["obs0"]), atol=2e-5)
    # post-reset obstacle state written by the generator == the host rule (obstacles.py:192-193)
    pos0, disp0, counter0 = scn.initial_obstacle_state(float(cfg.simulation.t_step_size))
    assert np.abs(env._pool["mov_pos0"].cpu().numpy() - pos0).max() < 1e-9
    assert np.abs(env._pool["mov_disp0"].cpu().numpy() - disp0).max() < 1e-12
    assert np.abs(env._pool["mov_counter0"].cpu().numpy() - counter0).max() < 1e-12


def test_refresh_finished_gives_every_new_episode_a_fresh_scenario():
    from gym_auv_b200.vec_env import AUVVecEnv

    cfg = lidar_config()
    cfg.episode.max_timesteps = 6
    N = 96
    scn = S.moving_obstacles(2 * N, 6, 6, seed=2, n_paths=4)
    env = AUVVecEnv(scn, N, cfg, auto_reset=True)
    env.reset()
    seen = []
    a = torch.zeros((N, 2), device="cuda")
    for t in range(20):
        env.step(a)
        if t % 3 == 2:
            before = env._pool["st_pos"].clone()
            n = env.refresh_finished(seed=77)
            changed = (env._pool["st_pos"] != before).flatten(1).any(1).nonzero().flatten().cpu().numpy()
            active = env._st["scn_id"].cpu().numpy()
            assert len(changed) == n and not set(changed) & set(active), "a running scenario was overwritten"
            seen.append(n)
    assert sum(seen) >= N  # every env timed out at least once
    with pytest.raises(ValueError):
        AUVVecEnv(S.moving_obstacles(N, 2, 2, seed=1), N, cfg, auto_reset=True).refresh_finished()


def test_bad_shapes_raise():
    from gym_auv_b200.vec_env import AUVVecEnv

    env = AUVVecEnv(S.empty_scenario(), 1, lidar_config(), auto_reset=False)
    with pytest.raises(ValueError):
        env.step(torch.zeros((2, 2), device="cuda"))


# ---------------------------------------------------------------------------------------
# gym.Env facade: the reference's tests/test_end_to_end.py over every registered id
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("scenario_name", sorted(S.SCENARIOS))
def test_single_step_end_to_end(scenario_name):
    import gym_auv_b200

    cfg = lidar_config() if scenario_name != "PathFollowNoObstacles-v0" else None
    env = gym_auv_b200.make(scenario_name, cfg)
    first_obs = env.reset()
    obs, reward, done, info = env.step(np.array([0.5, 0.6]))
    space = env.observation_space
    assert isinstance(obs, np.ndarray) and obs.shape == space.shape
    assert np.all(space.low <= obs) and np.all(space.high >= obs)
    assert isinstance(reward, float) and isinstance(done, bool) and isinstance(info, dict)
    assert set(info) == {"collision", "reached_goal", "goal_distance", "progress"}
    assert np.any(first_obs != obs)
    assert env.t_step == 1 and env.episode == 2  # constructor reset + explicit reset
    assert env.rewarder.params["lambda"] == 0.5


def test_dict_observation_facade():
    import gym_auv_b200

    cfg = lidar_config(use_dict_observation=True, sensor_use_velocity_observations=True)
    env = gym_auv_b200.make("TestScenario3-v0", cfg)
    obs = env.reset()
    assert set(obs) == {"proprioceptive", "lidar"}
    assert obs["proprioceptive"].shape == (6,) and obs["lidar"].shape == (3, 180)
    assert env.observation_space.contains({k: v.astype(np.float32) for k, v in obs.items()})


def test_auto_reset_uses_cached_first_observation_and_matches_fresh_reset():
    """The in-step reset copies the scenario's cached first observation; it must equal what an
    explicit reset() of that scenario returns, and the following steps must be identical to
    those of a freshly reset env (nearby list, obstacle state and counters included)."""
    from gym_auv_b200.vec_env import AUVVecEnv

    cfg = lidar_config()
    cfg.episode.max_timesteps = 4
    scn = S.moving_obstacles(6, 5, 5, seed=14)
    a = torch.as_tensor(random_actions(9, 6, 2), dtype=torch.float32, device="cuda")
    env = AUVVecEnv(scn, 6, cfg, test_mode=False, auto_reset=True, debug=True)
    fresh = AUVVecEnv(scn, 6, cfg, test_mode=False, auto_reset=False, debug=True)
    obs0 = fresh.reset().clone()
    env.reset()
    for t in range(4):
        obs, _, done, _ = env.step(a[t])
    assert bool(done.all())  # time limit
    assert torch.equal(obs, obs0)  # cached reset obs == explicit reset obs, bit for bit
    for t in range(4, 7):  # second episode of env == first episode of a fresh env, same actions
        o1, r1, d1, _ = env.step(a[t])
        o2, r2, d2, _ = fresh.step(a[t])
        assert torch.equal(o1, o2) and torch.equal(r1, r2) and torch.equal(d1, d2)
        assert torch.equal(env.get_attr("mov_pos"), fresh.get_attr("mov_pos"))


def test_record_capacity_overflow_is_reported_not_silent():
    from gym_auv_b200.vec_env import AUVVecEnv

    scn = S.test_scenario3()  # 21 circles, many within sensor range
    env = AUVVecEnv(scn, 1, lidar_config(), test_mode=True, auto_reset=False, max_nearby=2)
    with pytest.raises(RuntimeError, match="max_nearby"):
        env.reset()
    ok = AUVVecEnv(scn, 1, lidar_config(), test_mode=True, auto_reset=False)  # default capacity = all slots
    ok.reset()


def test_staged_entry_points_compose_to_a_step():
    """obstacle_update + vessel_step + observe (the staged ABI mirroring _update /
    Vessel.step / observe) == one auv_step."""
    from gym_auv_b200.vec_env import AUVVecEnv

    cfg = lidar_config()
    scn = S.moving_obstacles(5, 4, 4, seed=3)
    e1 = AUVVecEnv(scn, 5, cfg, test_mode=True, auto_reset=False)
    e2 = AUVVecEnv(scn, 5, cfg, test_mode=True, auto_reset=False)
    e1.reset(), e2.reset()
    a = torch.as_tensor(random_actions(5, 5, 1), dtype=torch.float32, device="cuda")
    for t in range(5):
        o1, r1, d1, _ = e1.step(a[t])
        e2.obstacle_update()
        e2.vessel_step(a[t])
        o2 = e2.observe()
        assert torch.equal(o1, o2) and torch.equal(r1, e2._out["reward"]) and torch.equal(e1.state, e2.state)
    nav = e2.navigate()
    assert nav.shape == (5, 24) and torch.equal(nav[:, :16], e1.get_attr("nav")[:, :16])


# ---------------------------------------------------------------------------------------
# CUDA path against goldens produced by the REFERENCE'S OWN classes behind import stubs
# (tests/golden/make_reference_goldens_stubbed.py) -- no oracle in between
# ---------------------------------------------------------------------------------------
STUBBED = np.load(os.path.join(GOLD, "reference_stubbed.npz"))


def test_culling_windows_match_reference_sensor_module_on_gpu():
    """k_vessel_nav's window integers == sensor._find_limit_angle_rays / the list indices of
    find_rays_to_simulate_for_obstacles (sensor.py:41-97), run on the reference's own code."""
    from gym_auv_b200.vec_env import AUVVecEnv

    cases, lim, err = STUBBED["win_cases"], STUBBED["win_limits"], STUBBED["win_index_error"]
    M, Rn = len(cases), 180
    base = S.empty_scenario()
    scn = S.ScenarioSet(
        waypoints=base.waypoints, path_id=np.zeros(M, dtype=np.int32), vessel_init=cases[:, [0, 1, 2]].copy(),
        mov_start=np.zeros((M, 0, 2)), mov_width=np.zeros((M, 0)), mov_track=np.zeros((M, 0, 4), dtype=np.int32),
        vel_table=np.zeros((0, 2)), st_pos=cases[:, None, 3:5].copy(), st_radius=cases[:, None, 5].copy(),
        rewarder="colav", post_generate_update=False, name="windows")
    scn._bank = base.bank
    env = AUVVecEnv(scn, M, lidar_config(), auto_reset=False, debug=True)
    env.reset()
    win = env.get_attr("windows").cpu().numpy()[:, 0]
    dist = np.hypot(cases[:, 3] - cases[:, 0], cases[:, 4] - cases[:, 1])
    surely_near = np.abs(dist - cases[:, 5]) < 140.0  # ring within sensor range whatever the polygonisation
    checked = 0
    for k in range(M):
        if not surely_near[k] or err[k]:
            continue
        assert (int(win[k, 0]), int(win[k, 1])) == (int(lim[k, 0]) - 1, int(lim[k, 1]) % Rn), k
        checked += 1
    assert checked > 250


@pytest.mark.parametrize("r", range(8))
def test_vessel_and_navigation_rollouts_match_reference_vessel_class_on_gpu(r):
    """State after Vessel.step and the Vessel.navigate features, step by step, against the
    reference's Vessel / Path classes (vessel.py:226-247,461-541; path.py)."""
    from gym_auv_b200.vec_env import AUVVecEnv

    pidx = int(STUBBED["roll_path"][r])
    wp = STUBBED["path_waypoints"][pidx]
    wp = wp[:, ~np.isnan(wp[0])]
    cfg = Config()  # no LiDAR: dynamics + navigation only
    cfg.simulation.t_step_size = float(STUBBED["roll_dt"][r])
    scn = S._single(wp, vessel_init=STUBBED["roll_init"][r], rewarder="pathfollow")
    env = AUVVecEnv(scn, 1, cfg, test_mode=True, auto_reset=False)
    env.reset()
    acts = torch.as_tensor(STUBBED["roll_actions"][r], dtype=torch.float32, device="cuda")
    for t in range(acts.shape[0]):
        obs, _, _, info = env.step(acts[t][None])
        nav = env.get_attr("nav")[0].cpu().numpy()
        st = env.state[:, 0].cpu().numpy()
        assert np.abs(st - STUBBED["roll_state"][r, t]).max() <= 1e-9, t
        assert abs(nav[0] - STUBBED["roll_s"][r, t]) <= 1e-7, t  # arclength (golden: brute-force stand-in for GEOS)
        assert abs(nav[3] - STUBBED["roll_s_la"][r, t]) <= 1e-7
        assert abs(nav[2] / 100 - STUBBED["roll_cte"][r, t]) <= 1e-9
        assert abs(nav[4] - STUBBED["roll_la_err"][r, t]) <= 2e-6  # FP32 atan2 of FP64 differences
        assert abs(nav[5] - STUBBED["roll_head_err"][r, t]) <= 2e-6
        assert abs(nav[6] - STUBBED["roll_goal"][r, t]) <= 1e-9
        assert abs(nav[7] - STUBBED["roll_progress"][r, t]) <= 1e-10
        assert bool(nav[10]) == bool(STUBBED["roll_reached"][r, t]) == bool(info["reached_goal"][0].item())
        assert abs(float(env.get_attr("max_progress")[0]) - STUBBED["roll_max_progress"][r, t]) <= 1e-10
        want_obs = np.clip([*STUBBED["roll_state"][r, t, 3:6], STUBBED["roll_la_err"][r, t], STUBBED["roll_head_err"][r, t],
                            STUBBED["roll_cte"][r, t]], -1, 1)
        assert np.abs(obs[0].cpu().numpy() - want_obs).max() <= 2e-6


@pytest.mark.parametrize("k", range(3))
def test_moving_obstacle_tracks_match_reference_vessel_obstacle_on_gpu(k):
    """k_obstacle_update (and the update fused into the step) against VesselObstacle.update
    (obstacles.py:195-215) on the reference's own class: constant and table-driven tracks, wrap."""
    from gym_auv_b200.vec_env import AUVVecEnv

    tr = STUBBED["trk_traj"][k]
    tr = tr[~np.isnan(tr[:, 0])]
    traj = [(int(t), (x, y)) for t, x, y in tr]
    dt = float(STUBBED["trk_dt"][k, 0])
    cfg = lidar_config()
    cfg.simulation.t_step_size = dt
    scn = S._single(np.array([[0.0, 1000.0], [500.0, 500.0]]), moving_traj=[(5.0, traj)])
    staged = AUVVecEnv(scn, 1, cfg, test_mode=True, auto_reset=False)
    fused = AUVVecEnv(scn, 1, cfg, test_mode=True, auto_reset=False)
    staged.reset(), fused.reset()
    assert np.abs(staged.get_attr("mov_pos")[0, 0].cpu().numpy() - STUBBED["trk_pos"][k, 0]).max() <= 1e-12
    a = torch.zeros((1, 2), device="cuda")
    for t in range(60):
        staged.obstacle_update()
        fused.step(a)
        for env in (staged, fused):
            pos = env.get_attr("mov_pos")[0, 0].cpu().numpy()
            disp = env.get_attr("mov_disp")[0, 0].cpu().numpy()
            assert np.abs(pos - STUBBED["trk_pos"][k, t + 1]).max() <= 1e-9, t
            assert abs(math.atan2(disp[1], disp[0]) - STUBBED["trk_head"][k, t + 1]) <= 1e-12
            assert abs(float(env.get_attr("mov_counter")[0, 0]) - STUBBED["trk_counter"][k, t + 1]) <= 1e-12


@pytest.mark.parametrize("k", range(5))
def test_whole_episodes_match_reference_base_environment_on_gpu(k):
    """One env through auv_step against the reference's BaseEnvironment episodes (gym / renderer
    stubbed, obstacle-free): obs, reward, done by time limit / reward limit / goal, info, and -- via
    the auto-reset that follows the done -- the env.history entry of the episode."""
    from gym_auv_b200.vec_env import AUVVecEnv
    from tests.test_reference_goldens_stubbed import _env_case

    cfg, scn, test_mode, T, D = _env_case(k)
    env = AUVVecEnv(scn, 1, cfg, test_mode=test_mode, auto_reset=True)
    obs0 = env.reset().cpu().numpy()[0].copy()
    assert obs0.shape == (D,) and np.abs(obs0 - STUBBED["env_obs0"][k][:D]).max() <= 2e-6
    acts = torch.as_tensor(STUBBED["env_actions"][k], dtype=torch.float32, device="cuda")
    for t in range(T):
        obs, rew, done, info = env.step(acts[t][None])
        d = bool(done[0].item())
        assert d == bool(STUBBED["env_done"][k, t]), t
        o = (info["terminal_observation"] if d else obs)[0].cpu().numpy()
        assert np.abs(o - STUBBED["env_obs"][k, t, :D]).max() <= 2e-6, t
        r = float(rew[0].item())
        assert abs(r - STUBBED["env_reward"][k, t]) <= 1e-5 * max(1.0, abs(r)), t  # reward[] is float32
        assert bool(info["reached_goal"][0].item()) == bool(STUBBED["env_reached"][k, t])
        assert not bool(info["collision"][0].item())
        assert abs(float(info["goal_distance"][0]) - STUBBED["env_goal"][k, t]) <= 1e-4
        assert abs(float(info["progress"][0]) - STUBBED["env_progress"][k, t]) <= 1e-6
        if not d:
            cum = float(env.get_attr("cumulative_reward")[0])
            assert abs(cum - STUBBED["env_cum"][k, t]) <= 1e-5 * max(1.0, abs(cum))
    h = STUBBED["env_history"][k]
    st = env.episode_stats(reduce=False)
    if bool(STUBBED["env_done"][k, T - 1]):  # the finished episode was filed by the in-step auto-reset
        assert st["episodes"] == 1.0
        assert abs(st["cross_track_error"] - h[0]) <= 1e-6 * max(1.0, h[0])
        assert st["reached_goal"] == h[1] and st["collision"] == h[2]
        assert abs(st["reward"] - h[3]) <= 1e-5 * max(1.0, abs(h[3]))
        assert st["timesteps"] == h[4] and abs(st["duration"] - h[5]) <= 1e-12
        assert abs(st["progress"] - h[6]) <= 1e-9 and abs(st["pathlength"] - h[7]) <= 1e-9
    else:
        assert st["episodes"] == 0.0


@pytest.mark.parametrize("k", range(6))
def test_lidar_pipeline_matches_reference_classes_on_geos_lite_on_gpu(k):
    """The CUDA step against the reference's own LiDAR pipeline classes run on geos_lite
    primitives (tests/golden/make_reference_goldens_hybrid.py): ranges within the FP32 casting
    tolerance, closeness / obs, rewards, collision and done flags -- no oracle code in the loop
    except those primitives."""
    from gym_auv_b200.vec_env import AUVVecEnv
    from tests._parity import COLLISION_BAND, RANGE_ATOL, RANGE_RTOL, REWARD_ATOL, REWARD_RTOL
    from tests.test_reference_goldens_stubbed import HYB, hybrid_case

    cfg, scn, T = hybrid_case(k)
    env = AUVVecEnv(scn, 1, cfg, test_mode=True, auto_reset=False, debug=True)
    obs0 = env.reset().cpu().numpy()[0]
    assert np.abs(obs0 - HYB["obs0"][k]).max() <= 2e-5
    acts = torch.as_tensor(HYB["actions"][k], dtype=torch.float32, device="cuda")
    for t in range(T):
        obs, rew, done, info = env.step(acts[t][None])
        d_ref = HYB["dists"][k, t]
        d_gpu = env.get_attr("lidar_dist")[0].cpu().numpy()
        assert np.all(np.abs(d_gpu - d_ref) <= RANGE_ATOL + RANGE_RTOL * d_ref), (t, np.abs(d_gpu - d_ref).max())
        assert np.abs(obs[0].cpu().numpy() - HYB["obs"][k, t]).max() <= 1e-4, t
        if abs(d_ref.min() - cfg.vessel.vessel_width) > COLLISION_BAND:
            assert bool(info["collision"][0].item()) == bool(HYB["collision"][k, t]), t
            assert bool(done[0].item()) == bool(HYB["done"][k, t]), t
            r = float(rew[0].item())
            assert abs(r - HYB["reward"][k, t]) <= REWARD_ATOL + REWARD_RTOL * abs(r), t
        assert int(env._scratch["rec_cnt"][0].item()) == int(HYB["n_nearby"][k, t]), t


@pytest.mark.parametrize("k", range(3))
def test_reference_movingobstacles_class_episodes_on_gpu(k):
    """The CUDA step on scenarios generated by the reference's own MovingObstaclesNoRules._generate,
    against that class's own step() (geometry primitives from geos_lite): BASELINE config 1."""
    from gym_auv_b200.vec_env import AUVVecEnv
    from tests._parity import COLLISION_BAND, RANGE_ATOL, RANGE_RTOL, REWARD_ATOL, REWARD_RTOL
    from tests.test_reference_goldens_stubbed import HYB, movingobstacles_case

    cfg, scn, T = movingobstacles_case(k)
    env = AUVVecEnv(scn, 1, cfg, test_mode=True, auto_reset=False, debug=True)
    assert np.abs(env.reset().cpu().numpy()[0] - HYB["mo_obs0"][k]).max() <= 2e-5
    acts = torch.as_tensor(HYB["mo_actions"][k], dtype=torch.float32, device="cuda")
    for t in range(T):
        obs, rew, done, info = env.step(acts[t][None])
        d_ref = HYB["mo_dists"][k, t]
        d_gpu = env.get_attr("lidar_dist")[0].cpu().numpy()
        assert np.all(np.abs(d_gpu - d_ref) <= RANGE_ATOL + RANGE_RTOL * d_ref), (t, np.abs(d_gpu - d_ref).max())
        assert np.abs(obs[0].cpu().numpy() - HYB["mo_obs"][k, t]).max() <= 1e-4, t
        assert int(env._scratch["rec_cnt"][0].item()) == int(HYB["mo_n_nearby"][k, t]), t
        if abs(d_ref.min() - cfg.vessel.vessel_width) > COLLISION_BAND:
            assert bool(info["collision"][0].item()) == bool(HYB["mo_collision"][k, t])
            assert bool(done[0].item()) == bool(HYB["mo_done"][k, t])
            r = float(rew[0].item())
            assert abs(r - HYB["mo_reward"][k, t]) <= REWARD_ATOL + REWARD_RTOL * abs(r), t


def test_gpu_scenario_generator_matches_reference_generate_statistics():
    """auv_generate_moving_obstacles against 12 scenarios drawn by the reference's own
    MovingObstacles._generate + helpers.generate_obstacle (204 vessels, 132 circles)."""
    from tests.test_reference_goldens_stubbed import HYB

    _, _, env = _generated_env(M=2048, n_paths=64, seed=4)
    env.regenerate_scenarios(seed=99, epoch=1)
    g = env.pull_scenarios()
    v0 = g.vessel_init[:, None, :2]
    stats = {
        "mov_width": g.mov_width.ravel(), "st_radius": g.st_radius.ravel(), "speed": np.linalg.norm(g.vel_table, axis=1),
        "mov_dist": np.linalg.norm(g.mov_start - v0, axis=2).ravel(), "st_dist": np.linalg.norm(g.st_pos - v0, axis=2).ravel(),
    }
    for key, mine in stats.items():
        ref = HYB["gen_" + key]
        se = ref.std() / math.sqrt(len(ref)) + mine.std() / math.sqrt(len(mine))
        assert abs(ref.mean() - mine.mean()) <= 4.0 * se, (key, ref.mean(), mine.mean(), se)


@pytest.mark.parametrize("name", ["TestScenario1", "TestScenario3", "TestScenario4", "TestHeadOn", "TestCrossing",
                                  "TestCrossing1", "EmptyScenario", "DebugScenario"])
def test_registered_scenario_episodes_match_reference_classes_on_gpu(name):
    """25 steps of each reference test-scenario class (its own step(), geos_lite primitives) against
    the CUDA step on the PRODUCT's definition of that scenario id."""
    from gym_auv_b200.vec_env import AUVVecEnv
    from tests._parity import RANGE_ATOL, RANGE_RTOL, REWARD_ATOL, REWARD_RTOL
    from tests.test_reference_goldens_stubbed import HYB, _product_scenario

    g = lambda k: HYB["ts_" + name + "_" + k]
    cfg = lidar_config()
    env = AUVVecEnv(_product_scenario(name), 1, cfg, test_mode=True, auto_reset=False, debug=True)
    assert np.abs(env.reset().cpu().numpy()[0] - g("obs0")).max() <= 2e-5
    acts = torch.as_tensor(g("actions"), dtype=torch.float32, device="cuda")
    for t in range(len(g("obs"))):
        obs, rew, done, _ = env.step(acts[t][None])
        d_ref, d_gpu = g("dists")[t], env.get_attr("lidar_dist")[0].cpu().numpy()
        assert np.all(np.abs(d_gpu - d_ref) <= RANGE_ATOL + RANGE_RTOL * d_ref), (t, np.abs(d_gpu - d_ref).max())
        assert np.abs(obs[0].cpu().numpy() - g("obs")[t]).max() <= 1e-4, t
        r = float(rew[0].item())
        assert abs(r - g("reward")[t]) <= REWARD_ATOL + REWARD_RTOL * abs(r), t
        assert bool(done[0].item()) == bool(g("done")[t])
