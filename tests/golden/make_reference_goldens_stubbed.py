"""Golden vectors from the REFERENCE'S OWN classes, run behind import stubs.

The reference package cannot be imported here (gym, shapely, pygame, tkinter are missing and
``config.py`` is invalid on Python 3.12 -- SURVEY.md section 8c).  Most of its hot-path classes,
however, only touch those packages at their edges.  This script builds a synthetic ``gym_auv``
package whose sub-modules are the reference's own files (found through ``__path__``, the
``__init__.py`` files are never executed) and puts minimal stand-ins for ``shapely`` / ``turtle``
into ``sys.modules``:

  * ``shapely.geometry.Point``: coordinates + ``distance`` (GEOS Coordinate::distance,
    sqrt(dx*dx + dy*dy)) + the array interface -- all that sensor.py:22-97 needs;
  * ``LineString`` / ``Polygon`` / ``affinity.rotate|translate``: inert containers (their geometry
    is never consulted by what is recorded below); ``LineString.project`` is a literal first-minimum
    brute force over segments, a STAND-IN for GEOS that only supplies the arclength to
    ``Vessel.navigate`` -- the goldens record that arclength, so everything downstream of the
    projection is the reference's arithmetic.

Recorded (unmodified reference code in every case):
  windows    sensor._find_limit_angle_rays / find_rays_to_simulate_for_obstacles
             (sensor.py:41-97): idx_min, idx_max and, per obstacle, how often each ray lists it
             (Python negative-index wrap incl. the IndexError corner)
  pooling    LidarPreprocessor._feasibility_pooling (sensor.py:251-296)
  paths      Path / RandomCurveThroughOrigin (path.py:19-120): waypoints, length, knots, samples of
             the position, direction, the 0.1 m polyline
  vessel     Vessel.step + Vessel.navigate rollouts (vessel.py:226-247,461-541) on those paths
  obstacles  VesselObstacle velocity table / update / wrap (obstacles.py:144-215)
  rewards    ColavRewarder.calculate / PathFollowRewarder.calculate (rewarder.py:78-241)
  env        BaseEnvironment.reset / step / observe / _isdone / save_latest_episode
             (environment.py:176-392,460-489) with ``gym`` and the 2-D renderer stubbed too, on
             obstacle-free scenarios (Vessel.perceive then never reaches Shapely, vessel.py:275-305):
             observation vectors, rewards, done flags, info, cumulative reward and the history entry
             of whole episodes ending by time limit, reward limit and reached goal

Usage:  python tests/golden/make_reference_goldens_stubbed.py   (writes reference_stubbed.npz)
"""
import importlib
import math
import os
import sys
import types

import numpy as np

REF_ROOT = "/root/reference"
REF = os.path.join(REF_ROOT, "gym_auv")
OUT = os.path.dirname(os.path.abspath(__file__))


# ----------------------------------------------------------------------------------- stubs
class Point:
    def __init__(self, *args):
        if len(args) == 1:
            a = args[0]
            a = a.coords[0] if isinstance(a, Point) else a
            self.x, self.y = float(a[0]), float(a[1])
        else:
            self.x, self.y = float(args[0]), float(args[1])

    @property
    def coords(self):
        return [(self.x, self.y)]

    def distance(self, other):
        dx, dy = self.x - other.x, self.y - other.y
        return math.sqrt(dx * dx + dy * dy)

    def __array__(self, dtype=None, copy=None):
        return np.array([self.x, self.y], dtype=dtype or np.float64)

    def __len__(self):
        return 2

    def __getitem__(self, i):
        return (self.x, self.y)[i]


class LineString:
    def __init__(self, coords):
        self.coords_arr = np.asarray(coords, dtype=np.float64)
        self.is_valid = True

    def project(self, point):
        """Stand-in for GEOS LengthIndexedLine.project: first strict minimum over the segments."""
        p = np.array([point.x, point.y])
        pts = self.coords_arr
        best, measure, start = math.inf, 0.0, 0.0
        for k in range(len(pts) - 1):
            a, b = pts[k], pts[k + 1]
            e = b - a
            l2 = float(e @ e)
            seglen = math.sqrt(l2)
            r = 0.0 if l2 == 0.0 else float((p - a) @ e) / l2
            rc = min(max(r, 0.0), 1.0)
            d = math.hypot(*(p - (a + rc * e)))
            if d < best:
                best, measure = d, start + rc * seglen
            start += seglen
        return measure


class Polygon:
    def __init__(self, coords):
        self.coords_arr = np.asarray(coords, dtype=np.float64)
        self.is_valid = True

    def buffer(self, *_a, **_k):
        return self


def install_stubs():
    sh = types.ModuleType("shapely")
    geo = types.ModuleType("shapely.geometry")
    geo.Point, geo.LineString, geo.Polygon = Point, LineString, Polygon
    aff = types.ModuleType("shapely.affinity")
    aff.rotate = lambda g, *_a, **_k: g
    aff.translate = lambda g, *_a, **_k: g
    sh.geometry, sh.affinity = geo, aff
    sys.modules.update({"shapely": sh, "shapely.geometry": geo, "shapely.affinity": aff})
    for name in ("errors", "strtree", "ops", "prepared"):
        m = types.ModuleType("shapely." + name)
        setattr(sh, name, m)
        sys.modules["shapely." + name] = m
    tu = types.ModuleType("turtle")
    tu.shape = None
    sys.modules["turtle"] = tu
    # synthetic package tree: the reference's files are found through __path__, its __init__.py
    # files (which import gym / pygame) are never run
    for pkg, rel in (("gym_auv", ""), ("gym_auv.utils", "utils"), ("gym_auv.objects", "objects"),
                     ("gym_auv.objects.vessel", "objects/vessel")):
        m = types.ModuleType(pkg)
        m.__path__ = [os.path.join(REF, rel)]
        sys.modules[pkg] = m
    sys.modules["gym_auv"].Config = type("Config", (), {})
    # gym 0.21 surface used by environment.py / clip_to_space.py, and the pygame-backed renderer
    gym = types.ModuleType("gym")
    spaces = types.ModuleType("gym.spaces")

    class Space:
        pass

    class Box(Space):
        def __init__(self, low, high, shape=None, dtype=np.float32):
            self.low = np.full(shape, low, dtype=np.float64) if shape is not None else np.asarray(low, dtype=np.float64)
            self.high = np.full(shape, high, dtype=np.float64) if shape is not None else np.asarray(high, dtype=np.float64)
            self.shape, self.dtype = self.low.shape, dtype

    class Dict(Space, dict):
        def __init__(self, d):
            dict.__init__(self, d)

    spaces.Space, spaces.Box, spaces.Dict = Space, Box, Dict
    gutils = types.ModuleType("gym.utils")
    seeding = types.ModuleType("gym.utils.seeding")
    seeding.np_random = lambda seed=None: (np.random.RandomState(seed), seed)
    gutils.seeding = seeding
    gym.Env, gym.spaces, gym.utils = type("Env", (), {}), spaces, gutils
    sys.modules.update({"gym": gym, "gym.spaces": spaces, "gym.utils": gutils, "gym.utils.seeding": seeding})
    r2d = types.ModuleType("gym_auv.render2d")
    r2d.__path__ = []
    st = types.ModuleType("gym_auv.render2d.state")
    st.RenderableState = type("RenderableState", (), {})
    rd = types.ModuleType("gym_auv.render2d.renderer")
    rd.Renderer2d, rd.FPS = type("Renderer2d", (), {}), 30
    r2d.state, r2d.renderer = st, rd
    sys.modules.update({"gym_auv.render2d": r2d, "gym_auv.render2d.state": st, "gym_auv.render2d.renderer": rd})


def ns(**kw):
    return types.SimpleNamespace(**kw)


def make_config(dt=1.0, n_sectors=9, per_sector=20):
    return ns(
        vessel=ns(vessel_width=1.255, n_sectors=n_sectors, n_sensors_per_sector=per_sector,
                  sensor_use_feasibility_pooling=False, sensor_log_transform=True, sensor_range=150.0,
                  look_ahead_distance=300.0, thrust_max_auv=2.0, moment_max_auv=0.15,
                  feasibility_width_multiplier=5.0),
        simulation=ns(sensor_frequency=1.0, observe_frequency=1.0, t_step_size=dt),
        episode=ns(min_goal_distance=5.0, min_path_progress=0.99),
    )


def main():
    install_stubs()
    sensor = importlib.import_module("gym_auv.objects.vessel.sensor")
    obst = importlib.import_module("gym_auv.objects.obstacles")
    pathm = importlib.import_module("gym_auv.objects.path")
    vesselm = importlib.import_module("gym_auv.objects.vessel.vessel")
    sys.modules["gym_auv.objects.vessel"].Vessel = vesselm.Vessel
    rew = importlib.import_module("gym_auv.objects.rewarder")
    out = {}

    # ---------------------------------------------------------------- culling windows
    rng = np.random.RandomState(0)
    R = 180
    dth = 2 * np.pi / R
    cases = []
    for k in range(400):
        p0 = rng.uniform(-300, 300, 2)
        heading = rng.uniform(-np.pi, np.pi)
        rho = float(rng.choice([1.0, 5.0, 11.18, 30.0, 80.0]))
        dist = rng.uniform(0.2, 4.0) * rho if k % 5 == 0 else rng.uniform(rho + 0.5, 200.0)
        ang = rng.uniform(-np.pi, np.pi) if k % 3 else heading + np.pi + rng.uniform(-0.2, 0.2)  # many dead astern
        c = p0 + dist * np.array([np.cos(ang), np.sin(ang)])
        cases.append((p0[0], p0[1], heading, c[0], c[1], rho))
    cases = np.array(cases)
    lim = np.zeros((len(cases), 2), dtype=np.int64)
    counts = np.zeros((len(cases), R), dtype=np.int8)
    index_error = np.zeros(len(cases), dtype=bool)

    class FakeObstacle:
        def __init__(self, c, r):
            self.enclosing_circle = obst.CircleParams(Point(c[0], c[1]), r)

    for k, (x, y, h, cx, cy, rho) in enumerate(cases):
        ob = FakeObstacle((cx, cy), rho)
        lim[k] = sensor._find_limit_angle_rays(ob.enclosing_circle, Point(x, y), h, dth)
        try:
            per_ray = sensor.find_rays_to_simulate_for_obstacles([ob], Point(x, y), h, dth, R)
            counts[k] = [len(lst) for lst in per_ray]
        except IndexError:
            index_error[k] = True
    out.update(win_cases=cases, win_limits=lim, win_counts=counts, win_index_error=index_error)

    # ---------------------------------------------------------------- feasibility pooling
    rng = np.random.RandomState(1)
    pool_in, pool_out = [], []
    for k in range(200):
        n = int(rng.choice([8, 9, 10, 15, 20, 51, 54]))
        m = np.full(n, 150.0)
        hit = rng.rand(n) < rng.uniform(0.1, 0.9)
        m[hit] = rng.uniform(0.0, 150.0, hit.sum())
        if k % 7 == 0:
            m[:] = rng.choice([150.0, 20.0])
        m32 = m.astype(np.float32).astype(np.float64)  # ranges are float32 on the GPU side
        pad = np.full(54, np.nan)
        pad[:n] = m32
        pool_in.append(pad)
        pool_out.append(float(sensor.LidarPreprocessor._feasibility_pooling(m32, 1.255 * 5.0, dth)))
    out.update(pool_in=np.array(pool_in), pool_out=np.array(pool_out))

    # ---------------------------------------------------------------- paths + vessel rollouts
    S = np.linspace(0.0, 1.0, 41)
    n_paths = 4
    T = 80
    path_wp, path_len, path_pos, path_dir, path_knots, path_poly_n, path_poly_head = [], [], [], [], [], [], []
    roll = {k: [] for k in ("state", "s", "s_la", "la_err", "head_err", "cte", "goal", "progress", "max_progress",
                            "reached", "actions", "init", "dt", "path")}
    for pidx in range(n_paths):
        prng = np.random.RandomState(100 + pidx)
        if pidx == n_paths - 1:  # the straight EmptyScenario path, testscenario.py:259-274
            path = pathm.Path(np.array([[25.0, 25.0], [10.0, 200.0]]))
        else:
            nwp = int(np.floor(4 * prng.rand() + 2))
            path = pathm.RandomCurveThroughOrigin(prng, nwp, 800)
        wp = np.full((2, 16), np.nan)
        wp[:, : path.init_waypoints.shape[1]] = path.init_waypoints
        path_wp.append(wp)
        path_len.append(path.length)
        path_pos.append(np.array([path(s * path.length) for s in S]))
        path_dir.append(np.array([path.get_direction(s * path.length) for s in S]))
        path_knots.append(np.array(path._arclengths))
        path_poly_n.append(len(path.points))
        path_poly_head.append(path.points[:: max(1, len(path.points) // 64)][:64])
        for rep in range(2):
            dt = [1.0, 0.5][rep]
            cfg = make_config(dt=dt)
            init = np.hstack([path(0.0) + 50 * (prng.rand(2) - 0.5),
                              path.get_direction(0.0) + 2 * np.pi * (prng.rand() - 0.5) * 0.3])
            v = vesselm.Vessel(cfg, init)
            v.navigate(path)  # BaseEnvironment.reset() observes once (environment.py:176-245)
            acts = prng.uniform([-0.2, -1.0], [1.0, 1.0], size=(T, 2)).astype(np.float32).astype(np.float64)
            rec = {k: [] for k in roll if k not in ("actions", "init", "dt", "path")}
            for a in acts:
                v.step(list(a))
                v.navigate(path)
                d = v._last_navi_state_dict
                rec["state"].append(v._state.copy())
                rec["s"].append(d["vessel_arclength"])
                rec["s_la"].append(d["target_arclength"])
                rec["la_err"].append(d["look_ahead_heading_error"])
                rec["head_err"].append(d["heading_error"])
                rec["cte"].append(d["cross_track_error"])
                rec["goal"].append(d["goal_distance"])
                rec["progress"].append(v._progress)
                rec["max_progress"].append(v._max_progress)
                rec["reached"].append(bool(v._reached_goal))
            for k in rec:
                roll[k].append(np.array(rec[k]))
            roll["actions"].append(acts)
            roll["init"].append(init)
            roll["dt"].append(dt)
            roll["path"].append(pidx)
    out.update(path_waypoints=np.array(path_wp), path_length=np.array(path_len), path_S=S, path_pos=np.array(path_pos),
               path_dir=np.array(path_dir), path_knots=np.array(path_knots), path_poly_n=np.array(path_poly_n),
               path_poly_sample=np.array(path_poly_head))
    out.update({"roll_" + k: np.array(v) for k, v in roll.items()})

    # ---------------------------------------------------------------- VesselObstacle tracks
    trk_pos, trk_head, trk_counter, trk_traj, trk_dt = [], [], [], [], []
    for k, (dt, n_pts, init_update) in enumerate([(1.0, 40, True), (0.5, 40, True), (1.0, 12, True), (0.3, 9, False)]):
        orng = np.random.RandomState(200 + k)
        if k < 2:  # movingobstacles.py:66-72: straight 10000-point track (shortened)
            start, direction, speed = orng.uniform(-100, 100, 2), orng.uniform(0, 2 * np.pi), orng.uniform(1, 3)
            traj = [[i, tuple(start + i * speed * np.array([np.cos(direction), np.sin(direction)]))] for i in range(n_pts)]
        else:  # table-driven, non-uniform time stamps (testscenario.py:302-350 style)
            t, traj = 0, []
            for i in range(n_pts):
                traj.append([t, tuple(orng.uniform(-50, 50, 2))])
                t += int(orng.randint(1, 4))
        o = obst.VesselObstacle(width=5.0, trajectory=traj, init_update=init_update)
        pos, head, cnt = [np.array(o.position, dtype=np.float64)], [float(o.heading)], [float(o.waypoint_counter)]
        for _ in range(60):
            o.update(dt)
            pos.append(np.array(o.position, dtype=np.float64))
            head.append(float(o.heading))
            cnt.append(float(o.waypoint_counter))
        tr = np.full((40, 3), np.nan)
        tr[: len(traj)] = [[tt, p[0], p[1]] for tt, p in traj]
        trk_traj.append(tr)
        trk_pos.append(np.array(pos))
        trk_head.append(np.array(head))
        trk_counter.append(np.array(cnt))
        trk_dt.append([dt, float(init_update)])
    out.update(trk_traj=np.array(trk_traj), trk_pos=np.array(trk_pos), trk_head=np.array(trk_head),
               trk_counter=np.array(trk_counter), trk_dt=np.array(trk_dt))

    # ---------------------------------------------------------------- rewarders
    cfg = make_config()
    angles = np.array([-np.pi + (i + 1) * dth for i in range(R)])
    rrng = np.random.RandomState(3)
    rin, rout_colav, rout_pf = [], [], []
    for k in range(300):
        d = np.full(R, 150.0)
        hit = rrng.rand(R) < rrng.uniform(0, 0.5)
        d[hit] = rrng.uniform(0.0, 150.0, hit.sum())
        d = d.astype(np.float32).astype(np.float64)
        speed, yaw = rrng.uniform(0, 0.6) * (rrng.rand() > 0.2), rrng.uniform(-0.2, 0.2)
        cte, he = rrng.uniform(-2, 2), rrng.uniform(-np.pi, np.pi)
        prog = rrng.uniform(0, 1)
        maxprog = prog if rrng.rand() < 0.5 else prog + rrng.uniform(0, 0.1)
        collision = bool(rrng.rand() < 0.1)
        fake = ns(
            req_latest_data=lambda d=d, cte=cte, he=he, collision=collision: {
                "navigation": {"cross_track_error": cte, "heading_error": he}, "distance_measurements": d,
                "speed_measurements": np.zeros((2, R)), "collision": collision},
            speed=speed, max_speed=2, yaw_rate=yaw, n_sensors=R, sensor_angles=angles, config=cfg, progress=prog,
            max_progress=maxprog)
        rc = rew.ColavRewarder(fake, test_mode=True)
        rp = rew.PathFollowRewarder(fake, test_mode=True)
        rin.append(np.hstack([speed, yaw, cte, he, prog, maxprog, float(collision), d]))
        rout_colav.append(float(rc.calculate()))
        rout_pf.append(float(rp.calculate()))
    out.update(rew_in=np.array(rin), rew_colav=np.array(rout_colav), rew_pathfollow=np.array(rout_pf))

    # ---------------------------------------------------------------- BaseEnvironment episodes
    import contextlib
    import io

    envm = importlib.import_module("gym_auv.environment")
    Config = sys.modules["gym_auv"].Config

    def env_config(dt, use_lidar, max_timesteps, min_cum):
        c = Config()
        base = make_config(dt=dt)
        c.vessel = base.vessel
        c.vessel.use_lidar = use_lidar
        c.vessel.dense_observation_size = 6
        c.vessel.n_lidar_observations = R
        c.vessel.use_dict_observation = False
        c.vessel.sensor_use_velocity_observations = False
        c.vessel.sensor_interval_load_obstacles = 25
        c.simulation = base.simulation
        c.episode = base.episode
        c.episode.max_timesteps = max_timesteps
        c.episode.min_cumulative_reward = min_cum
        return c

    ep = {k: [] for k in ("obs0", "obs", "reward", "done", "reached", "collision", "goal", "progress", "cum", "T",
                          "actions", "init", "wp", "cfg", "history")}
    Tmax = 120
    specs = [  # (rewarder, use_lidar, dt, max_timesteps, min_cumulative_reward, test_mode, start_s_frac, thrust)
        ("colav", True, 1.0, 40, -2000.0, False, 0.0, 0.3),      # ends by the time limit (t_step >= max - 1)
        ("colav", True, 1.0, 10000, -60.0, False, 0.0, -1.0),    # ends by the cumulative-reward limit (no thrust: slow penalty)
        ("colav", True, 1.0, 10000, -2000.0, False, 0.97, 1.0),  # reaches the goal (progress >= 0.99)
        ("pathfollow", False, 0.5, 10000, -2000.0, False, 0.90, 1.0),
        ("colav", True, 1.0, 30, -10.0, True, 0.0, 0.5),         # test_mode: neither limit ends the episode
    ]
    for k, (rname, use_lidar, dt, max_t, min_cum, test_mode, s_frac, thrust) in enumerate(specs):
        erng = np.random.RandomState(400 + k)
        wp = np.array([[0.0, 60.0, 150.0, 260.0], [0.0, 35.0, 20.0, -40.0]]) if k % 2 == 0 else np.array([[25.0, 25.0], [10.0, 200.0]])
        path0 = pathm.Path(wp)
        s0 = s_frac * path0.length
        init = np.hstack([path0(s0) + (erng.rand(2) - 0.5) * 4.0, path0.get_direction(s0) + (erng.rand() - 0.5) * 0.4])

        class Scn(envm.BaseEnvironment):
            def __init__(self, *a, **kw):
                self._rewarder_class = rew.ColavRewarder if rname == "colav" else rew.PathFollowRewarder
                self._n_moving_obst = self._n_moving_stat = 0
                super().__init__(*a, **kw)

            def _generate(self):
                self.path = pathm.Path(wp)
                self.vessel = vesselm.Vessel(self.config, init)
                self.obstacles = []
                self.rewarder = None

        cfg = env_config(dt, use_lidar, max_t, min_cum)
        with contextlib.redirect_stdout(io.StringIO()):  # the reference prints its config and episode info
            env = Scn(cfg, test_mode=test_mode, renderer=None)
            obs0 = np.array(env.observe())  # what reset() returned
            D = len(obs0)
            rec = {kk: [] for kk in ("obs", "reward", "done", "reached", "collision", "goal", "progress", "cum")}
            acts = np.stack([np.full(Tmax, thrust) + erng.uniform(-0.1, 0.1, Tmax), erng.uniform(-0.5, 0.5, Tmax)], axis=1)
            acts = acts.astype(np.float32).astype(np.float64)
            T = 0
            for a in acts:
                o, r_, d_, info = env.step(np.array(a))
                T += 1
                rec["obs"].append(np.array(o))
                rec["reward"].append(float(r_))
                rec["done"].append(bool(d_))
                rec["reached"].append(bool(info["reached_goal"]))
                rec["collision"].append(bool(info["collision"]))
                rec["goal"].append(float(info["goal_distance"]))
                rec["progress"].append(float(info["progress"]))
                rec["cum"].append(float(env.cumulative_reward))
                if d_ or T >= (60 if test_mode else Tmax):
                    break
            env.reset()  # files the finished episode under env.history (environment.py:476-489)
        h = env.history[-1]
        pad = lambda a, shape: np.concatenate([np.asarray(a, dtype=np.float64).reshape((len(a),) + shape[1:]),
                                               np.full((shape[0] - len(a),) + shape[1:], np.nan)])
        o186 = lambda o: np.concatenate([o, np.full(186 - len(o), np.nan)])
        ep["obs0"].append(o186(obs0))
        ep["obs"].append(pad([o186(o) for o in rec["obs"]], (Tmax, 186)))
        for kk in ("reward", "goal", "progress", "cum"):
            ep[kk].append(pad(rec[kk], (Tmax,)))
        for kk in ("done", "reached", "collision"):
            ep[kk].append(pad(np.array(rec[kk], dtype=np.float64), (Tmax,)))
        ep["T"].append(T)
        ep["actions"].append(acts)
        ep["init"].append(init)
        w = np.full((2, 4), np.nan)
        w[:, : wp.shape[1]] = wp
        ep["wp"].append(w)
        ep["cfg"].append([float(rname == "colav"), float(use_lidar), dt, max_t, min_cum, float(test_mode)])
        ep["history"].append([h["cross_track_error"], h["reached_goal"], h["collision"], h["reward"], h["timesteps"],
                              h["duration"], h["progress"], h["pathlength"]])
    out.update({"env_" + k: np.array(v) for k, v in ep.items()})

    np.savez_compressed(os.path.join(OUT, "reference_stubbed.npz"), **out)
    print("wrote", os.path.join(OUT, "reference_stubbed.npz"), {k: np.asarray(v).shape for k, v in out.items()})


if __name__ == "__main__":
    main()
