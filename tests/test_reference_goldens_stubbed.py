"""The oracle (and the host-side product code) against golden vectors produced by the
REFERENCE'S OWN classes run behind import stubs (tests/golden/make_reference_goldens_stubbed.py):
culling windows incl. Python's negative-index wrap, feasibility pooling, Path /
RandomCurveThroughOrigin, Vessel.step + Vessel.navigate rollouts, VesselObstacle tracks, both
rewarders.  These pin everything on the path that does not need a real GEOS."""
import math
import os
import types

import numpy as np
import pytest

from gym_auv_b200 import Config, lidar_config, scenarios as S
from gym_auv_b200.pathbank import build_path
from oracle import sim as O
from tests._parity import oracle_cfg

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_stubbed.npz"))
R = 180
DTH = 2 * np.pi / R


def test_culling_windows_are_the_reference_integers():
    lim = GOLD["win_limits"]
    n_seam = 0
    for k, (x, y, h, cx, cy, rho) in enumerate(GOLD["win_cases"]):
        lo, hi = O.limit_angle_rays((cx, cy), rho, (x, y), h, DTH)
        assert (lo, hi) == (int(lim[k, 0]), int(lim[k, 1])), k
        n_seam += int(lim[k, 1] >= R)
    assert n_seam > 20  # the fixture exercises the seam arithmetic


def test_candidate_rays_reproduce_pythons_negative_index_wrap():
    """How often each ray lists the obstacle in find_rays_to_simulate_for_obstacles: the closed
    form used by the kernels, candidate(i) <=> a <= i < b or a <= i - R < b with a = lo - 1,
    b = hi mod R, counted with multiplicity; IndexError upstream <=> a < -R."""
    lim, counts, err = GOLD["win_limits"], GOLD["win_counts"], GOLD["win_index_error"]
    i = np.arange(R)
    for k in range(len(lim)):
        a, b = int(lim[k, 0]) - 1, int(lim[k, 1]) % R
        assert bool(err[k]) == (a < -R), k
        if err[k]:
            continue
        mine = ((a <= i) & (i < b)).astype(int) + ((a <= i - R) & (i - R < b)).astype(int)
        assert np.array_equal(mine, counts[k]), k

    class Ob:
        def __init__(self, c, r):
            self.c, self.r = c, r

        def enclosing_circle(self):
            return self.c, self.r

    for k in range(0, len(lim), 7):  # the oracle's own list-based restatement
        x, y, h, cx, cy, rho = GOLD["win_cases"][k]
        per_ray, _ = O.rays_for_obstacles([Ob((cx, cy), rho)], (x, y), h, DTH, R)
        got = np.array([len(l) for l in per_ray])
        assert np.array_equal(got, counts[k]) or err[k]


def test_feasibility_pooling_matches_reference_static_method():
    for row, want in zip(GOLD["pool_in"], GOLD["pool_out"]):
        m = row[~np.isnan(row)]
        assert O.feasibility_pooling(m, 1.255 * 5.0, DTH) == want


@pytest.mark.parametrize("pidx", range(4))
def test_paths_match_reference_path_class(pidx):
    wp = GOLD["path_waypoints"][pidx]
    wp = wp[:, ~np.isnan(wp[0])]
    L = float(GOLD["path_length"][pidx])
    Sg = GOLD["path_S"]
    for path, direction in ((O.OraclePath(wp), lambda p, s: p.direction(s)), (build_path(wp), lambda p, s: p.direction(s))):
        assert abs(path.length - L) <= 1e-9 * L
        pos = np.array([path(s * L) for s in Sg]).reshape(len(Sg), 2)
        assert np.abs(pos - GOLD["path_pos"][pidx]).max() <= 1e-8
        d = np.array([float(direction(path, s * L)) for s in Sg])
        assert np.abs(np.angle(np.exp(1j * (d - GOLD["path_dir"][pidx])))).max() <= 1e-9
    tab = build_path(wp)
    assert len(tab.poly) == int(GOLD["path_poly_n"][pidx])  # int(10 * length) vertices, path.py:38
    step = max(1, len(tab.poly) // 64)
    assert np.abs(tab.poly[::step][:64] - GOLD["path_poly_sample"][pidx]).max() <= 1e-8
    assert np.abs(tab.knots - GOLD["path_knots"][pidx]).max() <= 1e-9 * L


def test_random_curve_generator_matches_reference_class():
    from gym_auv_b200.pathbank import random_curve_waypoints

    for pidx in range(3):
        rng = np.random.RandomState(100 + pidx)
        nwp = int(np.floor(4 * rng.rand() + 2))
        wp = random_curve_waypoints(rng, nwp, length=800)
        want = GOLD["path_waypoints"][pidx]
        want = want[:, ~np.isnan(want[0])]
        assert wp.shape == want.shape and np.abs(wp - want).max() <= 1e-12


@pytest.mark.parametrize("r", range(8))
def test_vessel_step_and_navigate_rollouts_match_reference_vessel(r):
    pidx = int(GOLD["roll_path"][r])
    wp = GOLD["path_waypoints"][pidx]
    wp = wp[:, ~np.isnan(wp[0])]
    cfg = lidar_config()
    cfg.simulation.t_step_size = float(GOLD["roll_dt"][r])
    path = O.OraclePath(wp)
    v = O.OracleVessel(oracle_cfg(cfg), GOLD["roll_init"][r])
    v.navigate(path)  # the observe() of reset()
    for t, a in enumerate(GOLD["roll_actions"][r]):
        v.step(a)
        v.navigate(path)
        assert np.abs(v.state - GOLD["roll_state"][r, t]).max() <= 1e-10, t
        # the golden arclength comes from the generator's brute-force stand-in for GEOS project
        assert abs(v.nav["vessel_arclength"] - GOLD["roll_s"][r, t]) <= 1e-7, t
        assert abs(v.nav["target_arclength"] - GOLD["roll_s_la"][r, t]) <= 1e-7
        assert abs(v.nav["look_ahead_heading_error"] - GOLD["roll_la_err"][r, t]) <= 1e-8
        assert abs(v.nav["heading_error"] - GOLD["roll_head_err"][r, t]) <= 1e-8
        assert abs(v.nav["cross_track_error"] - GOLD["roll_cte"][r, t]) <= 1e-9
        assert abs(v.nav["goal_distance"] - GOLD["roll_goal"][r, t]) <= 1e-9
        assert abs(v.progress - GOLD["roll_progress"][r, t]) <= 1e-10
        assert abs(v.max_progress - GOLD["roll_max_progress"][r, t]) <= 1e-10
        assert v.reached_goal == bool(GOLD["roll_reached"][r, t])


@pytest.mark.parametrize("k", range(4))
def test_vessel_obstacle_tracks_match_reference_class(k):
    tr = GOLD["trk_traj"][k]
    tr = tr[~np.isnan(tr[:, 0])]
    traj = [(int(t), (x, y)) for t, x, y in tr]
    dt, init_update = float(GOLD["trk_dt"][k, 0]), bool(GOLD["trk_dt"][k, 1])
    o = O.OracleVesselObstacle.from_trajectory(5.0, traj, init_update=init_update)
    assert np.abs(o.position - GOLD["trk_pos"][k, 0]).max() <= 1e-12
    for t in range(60):
        o.update(dt)
        assert np.abs(o.position - GOLD["trk_pos"][k, t + 1]).max() <= 1e-9, t
        assert abs(o.heading - GOLD["trk_head"][k, t + 1]) <= 1e-12
        assert abs(o.counter - GOLD["trk_counter"][k, t + 1]) <= 1e-12
    # the host-side product code (scenario arrays + vectorised update) on the same trajectory
    scn = S._single(np.array([[0.0, 100.0], [0.0, 0.0]]), moving_traj=[(5.0, traj)])
    pos = scn.mov_start.astype(np.float64).copy()
    counter = np.zeros(scn.mov_width.shape)
    if init_update:
        pos, _, counter = S.advance_obstacles(scn, pos, counter, 0.1)
    for t in range(60):
        pos, disp, counter = S.advance_obstacles(scn, pos, counter, dt)
        assert np.abs(pos[0, 0] - GOLD["trk_pos"][k, t + 1]).max() <= 1e-9, t
        assert abs(math.atan2(disp[0, 0, 1], disp[0, 0, 0]) - GOLD["trk_head"][k, t + 1]) <= 1e-12


def test_rewarders_match_reference_classes():
    angles = np.array([-np.pi + (i + 1) * DTH for i in range(R)])
    for row, want_c, want_p in zip(GOLD["rew_in"], GOLD["rew_colav"], GOLD["rew_pathfollow"]):
        speed, yaw, cte, he, prog, maxprog, collision = row[:7]
        v = types.SimpleNamespace(
            collision=bool(collision), nav=dict(cross_track_error=cte, heading_error=he), speed=speed,
            n_sensors=R, sensor_angles=angles, dists=row[7:], speeds=np.zeros((2, R)), cfg=dict(sensor_range=150.0), progress=prog,
            max_progress=maxprog, state=np.array([0, 0, 0, 0, 0, yaw]))
        assert abs(O.colav_reward(v) - want_c) <= 1e-9 * max(1.0, abs(want_c))
        assert abs(O.pathfollow_reward(v) - want_p) <= 1e-9 * max(1.0, abs(want_p))


def _env_case(k):
    """Scenario / config of BaseEnvironment episode k of the golden file."""
    colav, use_lidar, dt, max_t, min_cum, test_mode = GOLD["env_cfg"][k]
    wp = GOLD["env_wp"][k]
    wp = wp[:, ~np.isnan(wp[0])]
    cfg = lidar_config() if use_lidar else Config()
    cfg.simulation.t_step_size = float(dt)
    cfg.episode.max_timesteps = int(max_t)
    cfg.episode.min_cumulative_reward = float(min_cum)
    scn = S._single(wp, vessel_init=GOLD["env_init"][k], rewarder="colav" if colav else "pathfollow")
    return cfg, scn, bool(test_mode), int(GOLD["env_T"][k]), (186 if use_lidar else 6)


@pytest.mark.parametrize("k", range(5))
def test_whole_episodes_match_reference_base_environment(k):
    """reset / step / observe / _isdone of the reference's BaseEnvironment (gym and the renderer
    stubbed, obstacle-free scenarios): observation vectors, rewards, done by time limit / reward
    limit / reached goal, info, cumulative reward, and the env.history entry."""
    cfg, scn, test_mode, T, D = _env_case(k)
    env = O.OracleEnv(scn.describe(0), oracle_cfg(cfg), test_mode=test_mode)
    obs0 = env.observe()
    assert np.abs(obs0 - GOLD["env_obs0"][k][:D]).max() <= 1e-9
    for t in range(T):
        obs, rew, done, info = env.step(GOLD["env_actions"][k][t])
        assert np.abs(obs - GOLD["env_obs"][k, t, :D]).max() <= 1e-9, t
        assert abs(rew - GOLD["env_reward"][k, t]) <= 1e-9 * max(1.0, abs(rew)), t
        assert done == bool(GOLD["env_done"][k, t]), t
        assert info["reached_goal"] == bool(GOLD["env_reached"][k, t]) and not info["collision"]
        assert abs(info["goal_distance"] - GOLD["env_goal"][k, t]) <= 1e-9
        assert abs(info["progress"] - GOLD["env_progress"][k, t]) <= 1e-10
        assert abs(env.cumulative_reward - GOLD["env_cum"][k, t]) <= 1e-9 * max(1.0, abs(env.cumulative_reward))
    h = GOLD["env_history"][k]  # cross_track_error, reached_goal, collision, reward, timesteps, duration, progress, pathlength
    assert abs(np.mean(env.cross_track_errors) - h[0]) <= 1e-9
    assert env.t_step == int(h[4]) and abs(env.path.length - h[7]) <= 1e-9


# ---------------------------------------------------------------------------------------
# HYBRID goldens: the reference's LiDAR pipeline classes run on oracle/geos_lite primitives
# (tests/golden/make_reference_goldens_hybrid.py) -- pins the glue, not the primitives
# ---------------------------------------------------------------------------------------
HYB = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_hybrid.npz"))


def hybrid_case(k):
    wp = HYB["scn_waypoints"][k]
    wp = wp[:, ~np.isnan(wp[0])]
    Km, Ks = HYB["scn_mov_width"][k].shape[0], HYB["scn_st_radius"][k].shape[0]
    mov_track = np.zeros((1, Km, 4), dtype=np.int32)
    mov_track[0, :, 0] = np.arange(Km)
    mov_track[0, :, 1] = S.VESSEL_TRACK_LEN
    scn = S.ScenarioSet(
        waypoints=[wp], path_id=np.zeros(1, dtype=np.int32), vessel_init=HYB["scn_vessel_init"][k][None],
        mov_start=HYB["scn_mov_start"][k][None], mov_width=HYB["scn_mov_width"][k][None], mov_track=mov_track,
        vel_table=HYB["scn_vel"][k].copy(), st_pos=HYB["scn_st_pos"][k][None], st_radius=HYB["scn_st_radius"][k][None],
        rewarder="colav", post_generate_update=True, name="hybrid")
    cfg = lidar_config()
    cfg.simulation.t_step_size = float(HYB["dt"][k])
    return cfg, scn, int(HYB["T"][k])


@pytest.mark.parametrize("k", range(6))
def test_lidar_pipeline_matches_reference_classes_on_geos_lite(k):
    """Vessel.perceive / simulate_sensor / _standardize_intersect / obstacle classes of the
    reference, run unmodified on geos_lite primitives, against the oracle's restatement of the same
    pipeline: nearby-list size and refresh, every ray's range, closeness, collision, obs, reward."""
    cfg, scn, T = hybrid_case(k)
    env = O.OracleEnv(scn.describe(0), oracle_cfg(cfg), test_mode=True)
    assert np.abs(env.observe() - HYB["obs0"][k]).max() <= 1e-9
    for t in range(T):
        obs, rew, done, info = env.step(HYB["actions"][k][t])
        assert len(env.vessel.nearby) == int(HYB["n_nearby"][k, t]), t
        assert np.abs(env.vessel.dists - HYB["dists"][k, t]).max() <= 1e-9, t
        assert np.abs(obs - HYB["obs"][k, t]).max() <= 1e-9, t
        assert abs(rew - HYB["reward"][k, t]) <= 1e-9 * max(1.0, abs(rew)), t
        assert info["collision"] == bool(HYB["collision"][k, t]) and done == bool(HYB["done"][k, t])
        assert info["reached_goal"] == bool(HYB["reached"][k, t])


def movingobstacles_case(k):
    """Scenario k generated by the reference's MovingObstaclesNoRules._generate (read back from the env)."""
    HYB2 = HYB
    wp = HYB2["mo_waypoints"][k]
    wp = wp[:, ~np.isnan(wp[0])]
    Km, Ks = HYB2["mo_mov_width"][k].shape[0], HYB2["mo_st_radius"][k].shape[0]
    mov_track = np.zeros((1, Km, 4), dtype=np.int32)
    mov_track[0, :, 0] = np.arange(Km)
    mov_track[0, :, 1] = S.VESSEL_TRACK_LEN
    scn = S.ScenarioSet(
        waypoints=[wp], path_id=np.zeros(1, dtype=np.int32), vessel_init=HYB2["mo_vessel_init"][k][None],
        mov_start=HYB2["mo_mov_start"][k][None], mov_width=HYB2["mo_mov_width"][k][None], mov_track=mov_track,
        vel_table=HYB2["mo_vel"][k].copy(), st_pos=HYB2["mo_st_pos"][k][None], st_radius=HYB2["mo_st_radius"][k][None],
        rewarder="colav", post_generate_update=True, name="MovingObstaclesNoRules-v0")
    return lidar_config(), scn, int(HYB2["mo_T"][k])


@pytest.mark.parametrize("k", range(3))
def test_reference_movingobstacles_class_episodes(k):
    """Episodes of the reference's own MovingObstaclesNoRules (17 vessels + 11 circles, its own
    _generate and step(), geometry primitives from geos_lite) replayed through the oracle."""
    cfg, scn, T = movingobstacles_case(k)
    assert scn.k_moving == 17 and scn.k_static == 11
    env = O.OracleEnv(scn.describe(0), oracle_cfg(cfg), test_mode=True)
    assert np.abs(env.observe() - HYB["mo_obs0"][k]).max() <= 1e-9
    for t in range(T):
        obs, rew, done, info = env.step(HYB["mo_actions"][k][t])
        assert len(env.vessel.nearby) == int(HYB["mo_n_nearby"][k, t]), t
        assert np.abs(env.vessel.dists - HYB["mo_dists"][k, t]).max() <= 1e-9, t
        assert np.abs(obs - HYB["mo_obs"][k, t]).max() <= 1e-9, t
        assert abs(rew - HYB["mo_reward"][k, t]) <= 1e-9 * max(1.0, abs(rew)), t
        assert info["collision"] == bool(HYB["mo_collision"][k, t]) and done == bool(HYB["mo_done"][k, t])


def test_host_scenario_generator_matches_reference_generate_statistics():
    """12 scenarios drawn by the reference's own MovingObstacles._generate + helpers.generate_obstacle
    (204 vessels, 132 circles) against the host generator's distribution (same parameters, own RNG
    streams -- the reference mixes a seeded stream with the global np.random)."""
    mine = S.moving_obstacles(400, 17, 11, seed=123)
    v0 = mine.vessel_init[:, None, :2]
    stats = {
        "mov_width": mine.mov_width.ravel(), "st_radius": mine.st_radius.ravel(),
        "speed": np.linalg.norm(mine.vel_table, axis=1),
        "mov_dist": np.linalg.norm(mine.mov_start - v0, axis=2).ravel(),
        "st_dist": np.linalg.norm(mine.st_pos - v0, axis=2).ravel(),
    }
    for key, mine_v in stats.items():
        ref = HYB["gen_" + key]
        se = ref.std() / math.sqrt(len(ref)) + mine_v.std() / math.sqrt(len(mine_v))
        assert abs(ref.mean() - mine_v.mean()) <= 4.0 * se, (key, ref.mean(), mine_v.mean(), se)
    off = HYB["gen_init_offset"]
    assert np.abs(off).max() <= 25.0 and HYB["gen_mov_width"].min() >= 1 and HYB["gen_speed"].min() >= 1.0


TS_NAMES = ["TestScenario1", "TestScenario2", "TestScenario3", "TestScenario4", "TestHeadOn", "TestCrossing",
            "TestCrossing1", "EmptyScenario", "DebugScenario"]


def _product_scenario(name):
    """The product's scenario of that id; TestHeadOn's start angle is drawn from the global `random`
    upstream (testscenario.py:145) and an explicit argument here: take it from the recorded start."""
    if name == "TestHeadOn":
        sx, sy = HYB["ts_TestHeadOn_mov_start"][0] - HYB["ts_TestHeadOn_vessel_init"][:2]
        return S.test_head_on(start_angle=math.atan2(sx, sy))
    return S.SCENARIOS[name + "-v0"]()


@pytest.mark.parametrize("name", TS_NAMES)
def test_registered_scenarios_equal_the_reference_scenario_classes(name):
    """gym_auv_b200.scenarios.SCENARIOS[<id>] against what the reference's envs/testscenario.py class
    of the same id builds (instantiated behind the stubs, scenario read back from the env): path,
    vessel start, every circle, every vessel track (start, width, velocity table)."""
    g = lambda k: HYB["ts_" + name + "_" + k]
    scn = _product_scenario(name)
    assert np.abs(scn.waypoints[0] - g("waypoints")).max() <= 1e-12
    vi = scn.vessel_init[0]
    assert np.abs(vi[:2] - g("vessel_init")[:2]).max() <= 1e-9
    assert abs(math.remainder(vi[2] - g("vessel_init")[2], 2 * math.pi)) <= 1e-12
    used = scn.st_radius[0] > 0
    assert used.sum() == len(g("st_radius"))
    if used.any():
        assert np.abs(scn.st_pos[0][used] - g("st_pos")).max() <= 1e-9
        assert np.abs(scn.st_radius[0][used] - g("st_radius")).max() <= 1e-12
    mused = scn.mov_width[0] > 0
    assert mused.sum() == len(g("mov_width"))
    for j, slot in enumerate(np.nonzero(mused)[0]):
        off, ln, stride, _ = scn.mov_track[0, slot]
        assert scn.mov_width[0, slot] == g("mov_width")[j] and ln == int(g("mov_nvel")[j])
        assert np.abs(scn.mov_start[0, slot] - g("mov_start")[j]).max() <= 1e-9
        head = scn.vel_table[off + np.arange(8) * stride]
        assert np.abs(head - g("mov_vel_head")[j]).max() <= 1e-9


@pytest.mark.parametrize("name", [n for n in TS_NAMES if n != "TestScenario2"])
def test_registered_scenario_episodes_match_reference_classes(name):
    """A 25-step episode of each reference test-scenario class (its own step(), geos_lite
    primitives) replayed through the oracle on the PRODUCT's scenario definition."""
    g = lambda k: HYB["ts_" + name + "_" + k]
    scn = _product_scenario(name)
    env = O.OracleEnv(scn.describe(0), oracle_cfg(lidar_config()), test_mode=True)
    assert np.abs(env.observe() - g("obs0")).max() <= 1e-9
    for t in range(len(g("obs"))):
        obs, rew, done, _ = env.step(g("actions")[t])
        assert np.abs(env.vessel.dists - g("dists")[t]).max() <= 1e-9, t
        assert np.abs(obs - g("obs")[t]).max() <= 1e-9 and abs(rew - g("reward")[t]) <= 1e-9 * max(1.0, abs(rew))
        assert done == bool(g("done")[t])


@pytest.mark.parametrize("k", range(3))
def test_reference_pathfollow_class_episodes(k):
    """BASELINE config 2: the reference's own PathFollowNoObstacles class (no obstacles,
    PathFollowRewarder, use_lidar=False, 6-dimensional observation) replayed through the oracle."""
    wp = HYB["pf_waypoints"][k]
    wp = wp[:, ~np.isnan(wp[0])]
    scn = S._single(wp, vessel_init=HYB["pf_vessel_init"][k], rewarder="pathfollow")
    env = O.OracleEnv(scn.describe(0), oracle_cfg(Config()), test_mode=True)
    obs0 = env.observe()
    assert obs0.shape == (6,) and np.abs(obs0 - HYB["pf_obs0"][k]).max() <= 1e-9
    for t in range(int(HYB["pf_T"][k])):
        obs, rew, done, _ = env.step(HYB["pf_actions"][k][t])
        assert np.abs(obs - HYB["pf_obs"][k, t]).max() <= 1e-9, t
        assert abs(rew - HYB["pf_reward"][k, t]) <= 1e-9 * max(1.0, abs(rew)), t
        assert done == bool(HYB["pf_done"][k, t])
