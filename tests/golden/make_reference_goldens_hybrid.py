"""HYBRID goldens: the reference's own LiDAR pipeline classes on top of ``oracle/geos_lite.py``.

``make_reference_goldens_stubbed.py`` pins everything that never reaches Shapely.  This script
goes one step further: it gives the stubbed ``shapely`` just enough REAL geometry -- every
primitive delegated to ``oracle/geos_lite.py`` (the restated GEOS algorithms) -- for the
reference's ``BaseEnvironment`` / ``Vessel.perceive`` / ``sensor.simulate_sensor`` /
``_standardize_intersect`` / ``CircularObstacle`` / ``VesselObstacle`` / ``PolygonObstacle`` /
``enclosing_circle_of_shape`` to run UNMODIFIED on scenarios WITH obstacles.  What these goldens
pin is therefore the reference's glue: which obstacles are nearby and when the list is refreshed,
which obstacles each ray tests, that the range is the minimum over all intersection pieces, ring
vs filled boundaries, closeness, collision, the observation vector, reward and done -- everything
except the numerical primitives themselves (buffer + simplify, rotate about the centroid,
minimum rotated rectangle, intersection, distance, project), which stay PARITY UNPINNED and are
checked separately (tests/test_oracle_geometry.py, tests/test_oracle_exact.py).

Usage:  python tests/golden/make_reference_goldens_hybrid.py   (writes reference_hybrid.npz)
"""
import contextlib
import importlib
import io
import math
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from oracle import geos_lite as G  # noqa: E402
import make_reference_goldens_stubbed as stubbed  # noqa: E402


# ------------------------------------------------------------------ geometry on geos_lite
class Geom:
    is_empty = False
    is_valid = True


class Empty(Geom):
    is_empty = True


class Point(Geom):
    def __init__(self, *args):
        if len(args) == 1:
            a = args[0]
            a = (a.x, a.y) if isinstance(a, Point) else a
            self.x, self.y = float(a[0]), float(a[1])
        else:
            self.x, self.y = float(args[0]), float(args[1])

    @property
    def coords(self):
        return [(self.x, self.y)]

    def distance(self, other):
        p = (self.x, self.y)
        if isinstance(other, Point):
            dx, dy = self.x - other.x, self.y - other.y
            return math.sqrt(dx * dx + dy * dy)
        if isinstance(other, Polygon):
            return G.point_polygon_distance(p, other.ring)
        if isinstance(other, LineString):
            return G.point_ring_distance(p, other.ring)
        raise TypeError(type(other))

    def buffer(self, radius):
        return Polygon(G.buffer_point_ring(self.x, self.y, radius))

    def __array__(self, dtype=None, copy=None):
        return np.array([self.x, self.y], dtype=dtype or np.float64)

    def __len__(self):
        return 2

    def __getitem__(self, i):
        return (self.x, self.y)[i]


def _xy(c):
    return (c.x, c.y) if isinstance(c, Point) else (float(c[0]), float(c[1]))


class LineString(Geom):
    def __init__(self, coords):
        self.ring = np.array([_xy(c) for c in coords], dtype=np.float64)

    @property
    def coords(self):
        return [tuple(p) for p in self.ring]

    def simplify(self, tol, preserve_topology=True):
        assert not preserve_topology
        return LineString(G.douglas_peucker(self.ring, tol))

    def project(self, point):
        return G.linestring_project(self.ring, (point.x, point.y))

    def intersection(self, other):
        """Only the ray (a 2-point line) is ever intersected with a boundary."""
        p0, p1 = self.ring[0], self.ring[1]
        length = math.hypot(*(p1 - p0))
        d = (p1 - p0) / length
        if isinstance(other, Polygon):  # filled: the clipped ray starts at P0 if inside, else at the entry
            t = G.ray_polygon_min_distance(p0, p1, other.ring)
            if t is None:
                return Empty()
            start = p0 + t * d
            return LineString([start, start])  # only coords[0] is consulted (sensor.py:14-15)
        hits = []
        for k in range(len(other.ring) - 1):
            hits += G._segment_hits(p0[0], p0[1], p1[0], p1[1], *other.ring[k], *other.ring[k + 1])
        if not hits:
            return Empty()
        pts = [Point(*(p0 + t * d)) for t in sorted(set(hits))]
        return pts[0] if len(pts) == 1 else MultiPoint(pts)


class MultiPoint(Geom):
    def __init__(self, pts):
        self.geoms = list(pts)


class _Exterior:
    def __init__(self, ring):
        self.coords = [tuple(p) for p in ring]


class Polygon(Geom):
    def __init__(self, coords):
        ring = np.array([_xy(c) for c in coords], dtype=np.float64)
        if not np.array_equal(ring[0], ring[-1]):
            ring = np.vstack([ring, ring[:1]])
        self.ring = ring

    @property
    def boundary(self):
        return LineString(self.ring)

    @property
    def exterior(self):
        return _Exterior(self.ring)

    @property
    def centroid(self):
        return Point(*G.polygon_centroid(self.ring))

    @property
    def minimum_rotated_rectangle(self):
        return Polygon(G.minimum_rotated_rectangle(self.ring))

    def buffer(self, *_a, **_k):
        return self


def rotate(poly, angle, use_radians=False, origin="centroid"):
    assert use_radians and origin == "centroid"
    return Polygon(G.rotate_about(poly.ring, angle, G.polygon_centroid(poly.ring)))


def translate(poly, xoff=0.0, yoff=0.0):
    return Polygon(poly.ring + np.array([xoff, yoff]))


def install():
    stubbed.install_stubs()
    geo, aff = sys.modules["shapely.geometry"], sys.modules["shapely.affinity"]
    geo.Point, geo.LineString, geo.Polygon, geo.MultiPoint = Point, LineString, Polygon, MultiPoint
    aff.rotate, aff.translate = rotate, translate


# ------------------------------------------------------------------ scenarios
def main():
    install()
    from gym_auv_b200 import scenarios as S  # host-side generator of the product (pure NumPy)

    obst = importlib.import_module("gym_auv.objects.obstacles")
    pathm = importlib.import_module("gym_auv.objects.path")
    vesselm = importlib.import_module("gym_auv.objects.vessel.vessel")
    sys.modules["gym_auv.objects.vessel"].Vessel = vesselm.Vessel
    rew = importlib.import_module("gym_auv.objects.rewarder")
    envm = importlib.import_module("gym_auv.environment")
    Config = sys.modules["gym_auv"].Config

    def env_config(dt):
        c = Config()
        base = stubbed.make_config(dt=dt)
        c.vessel, c.simulation, c.episode = base.vessel, base.simulation, base.episode
        c.vessel.use_lidar = True
        c.vessel.dense_observation_size = 6
        c.vessel.n_lidar_observations = 180
        c.vessel.use_dict_observation = False
        c.vessel.sensor_use_velocity_observations = False
        c.vessel.sensor_interval_load_obstacles = 25
        c.episode.max_timesteps = 10000
        c.episode.min_cumulative_reward = -2000.0
        return c

    T = 60
    out = {k: [] for k in ("obs0", "obs", "reward", "done", "collision", "reached", "dists", "n_nearby", "cum", "T",
                           "actions", "seed", "dt", "close")}
    cases = [(0, 1.0, False), (1, 0.5, False), (2, 1.0, True), (3, 1.0, True), (4, 0.5, True), (5, 1.0, True)]
    def run_case(seed, dt, close, attempt):
        scn = S.moving_obstacles(1, 6, 5, seed=500 + seed)
        rng = np.random.RandomState(600 + seed + 100 * attempt)
        if close:  # start next to an obstacle so that short ranges, inside-ring and collisions occur
            if seed % 2 == 0:
                j = rng.randint(5)
                ang = rng.uniform(-np.pi, np.pi)
                r = scn.st_radius[0, j] * (0.6 if seed == 4 else 1.0) + (0.0 if seed == 4 else rng.uniform(2.0, 8.0))
                scn.vessel_init[0, :2] = scn.st_pos[0, j] + r * np.array([np.cos(ang), np.sin(ang)])
                scn.vessel_init[0, 2] = ang + np.pi + rng.uniform(-0.3, 0.3)
            else:
                j = rng.randint(6)
                v = scn.vel_table[scn.mov_track[0, j, 0]]
                scn.vessel_init[0, :2] = scn.mov_start[0, j] + v * rng.uniform(10, 25) + rng.uniform(-2, 2, size=2)
                scn.vessel_init[0, 2] = rng.uniform(-np.pi, np.pi)
        d = scn.describe(0)
        acts = rng.uniform([0.0, -0.15], [1.0, 0.15], size=(T, 2)).astype(np.float32).astype(np.float64)

        class Scn(envm.BaseEnvironment):
            def __init__(self, *a, **kw):
                self._rewarder_class = rew.ColavRewarder
                self._n_moving_obst, self._n_moving_stat = 6, 5
                super().__init__(*a, **kw)

            def _generate(self):
                self.path = pathm.Path(np.array(d["waypoints"]))
                self.vessel = vesselm.Vessel(self.config, np.array(d["vessel_init"]))
                self.obstacles = []
                mov = d["moving"]
                for j in range(len(mov["width"])):  # movingobstacles.py:51-75: 10000-point linear track
                    vel = mov["vel_tables"][j][0]
                    n = len(mov["vel_tables"][j]) + 1
                    traj = [[i, tuple(mov["start"][j] + i * vel)] for i in range(n)]
                    self.obstacles.append(obst.VesselObstacle(width=float(mov["width"][j]), trajectory=traj))
                for j in range(len(d["static"]["radius"])):
                    self.obstacles.append(obst.CircularObstacle(d["static"]["pos"][j], float(d["static"]["radius"][j])))
                self.rewarder = None
                self._update()  # movingobstacles.py:95

        with contextlib.redirect_stdout(io.StringIO()):
            env = Scn(env_config(dt), test_mode=True, renderer=None)
            obs0 = np.array(env.observe())  # what reset() returned (step counter 0: same nearby list)
            rec = {k: [] for k in ("obs", "reward", "done", "collision", "reached", "dists", "n_nearby", "cum")}
            n = 0
            for a in acts:
                o, r_, d_, info = env.step(np.array(a))
                n += 1
                rec["obs"].append(np.array(o))
                rec["reward"].append(float(r_))
                rec["done"].append(bool(d_))
                rec["collision"].append(bool(info["collision"]))
                rec["reached"].append(bool(info["reached_goal"]))
                rec["dists"].append(np.array(env.vessel._last_sensor_dist_measurements, dtype=np.float64))
                rec["n_nearby"].append(len(env.vessel._nearby_obstacles))
                rec["cum"].append(float(env.cumulative_reward))
                if d_:
                    break
        return scn, rec, n, obs0, acts

    scenario_arrays = {k: [] for k in ("vessel_init", "mov_start", "mov_width", "vel", "st_pos", "st_radius", "waypoints")}
    for seed, dt, close in cases:
        for attempt in range(20):  # placements that end the episode within a few steps are redrawn
            scn, rec, n, obs0, acts = run_case(seed, dt, close, attempt)
            if n >= 12:
                break
        pad = lambda a, shape: np.concatenate([np.asarray(a, dtype=np.float64).reshape((len(a),) + shape[1:]),
                                               np.full((shape[0] - len(a),) + shape[1:], np.nan)])
        out["obs0"].append(obs0)
        out["obs"].append(pad(rec["obs"], (T, 186)))
        out["dists"].append(pad(rec["dists"], (T, 180)))
        for k in ("reward", "cum"):
            out[k].append(pad(rec[k], (T,)))
        for k in ("done", "collision", "reached", "n_nearby"):
            out[k].append(pad(np.array(rec[k], dtype=np.float64), (T,)))
        out["T"].append(n)
        out["actions"].append(acts)
        out["seed"].append(seed)
        out["dt"].append(dt)
        out["close"].append(float(close))
        # the scenario itself, so that tests do not depend on the generator's RNG stream
        scenario_arrays["vessel_init"].append(scn.vessel_init[0])
        scenario_arrays["mov_start"].append(scn.mov_start[0])
        scenario_arrays["mov_width"].append(scn.mov_width[0])
        scenario_arrays["vel"].append(scn.vel_table[scn.mov_track[0, :, 0]])
        scenario_arrays["st_pos"].append(scn.st_pos[0])
        scenario_arrays["st_radius"].append(scn.st_radius[0])
        w = np.full((2, 16), np.nan)
        w[:, : scn.waypoints[0].shape[1]] = scn.waypoints[0]
        scenario_arrays["waypoints"].append(w)
        print("case", seed, "steps", n, "collision", any(rec["collision"]), "min range %.3f" % np.min(rec["dists"]),
              "nearby", sorted(set(rec["n_nearby"])))
    out.update({"scn_" + k: v for k, v in scenario_arrays.items()})
    # ------------------------------------------------------------------ the reference's own scenario class
    # MovingObstaclesNoRules (envs/movingobstacles.py:17-104: 17 vessels + 11 circles, ColavRewarder)
    # generates its scenarios with its own RNG calls; the generated scenario is read back from the
    # env (path waypoints, vessel start, obstacle tracks / circles) so that the same episode can be
    # replayed through the oracle and the CUDA path ("the reference's own step() on identical inputs").
    envs_pkg = types.ModuleType("gym_auv.envs")
    envs_pkg.__path__ = [os.path.join(stubbed.REF, "envs")]
    sys.modules["gym_auv.envs"] = envs_pkg
    mo = importlib.import_module("gym_auv.envs.movingobstacles")
    Tm = 50
    mo_out = {k: [] for k in ("waypoints", "vessel_init", "mov_start", "mov_width", "vel", "st_pos", "st_radius", "obs0",
                              "obs", "reward", "done", "collision", "dists", "n_nearby", "actions", "T")}
    gen_stats = {k: [] for k in ("mov_width", "st_radius", "speed", "mov_dist", "st_dist", "init_offset")}
    for seed in range(12):
        with contextlib.redirect_stdout(io.StringIO()):
            env = mo.MovingObstaclesNoRules(env_config(1.0), test_mode=True, renderer=None)
            # the constructor seeds itself from the OS (environment.py:96,439-442); reseed the way
            # SURVEY 8d prescribes and generate the scenario that is recorded
            np.random.seed(seed)
            env.seed(seed)
            env.reset()
            mov = [o for o in env.obstacles if not o.static]
            sta = [o for o in env.obstacles if o.static]
            v0 = np.array(env.vessel._state[:3])
            start = np.array([o.trajectory[0][1] for o in mov])
            vel = np.array([o.trajectory_velocities[0] for o in mov])
            width = np.array([o.width for o in mov], dtype=np.float64)
            spos = np.array([o.position for o in sta])
            srad = np.array([o.radius for o in sta], dtype=np.float64)
            gen_stats["mov_width"] += list(width)
            gen_stats["st_radius"] += list(srad)
            gen_stats["speed"] += list(np.linalg.norm(vel, axis=1))
            gen_stats["mov_dist"] += list(np.linalg.norm(start - v0[:2], axis=1))
            gen_stats["st_dist"] += list(np.linalg.norm(spos - v0[:2], axis=1))
            gen_stats["init_offset"] += list(v0[:2] - env.path(0))
            if seed >= 3:
                continue  # seeds 3.. only feed the generator statistics
            obs0 = np.array(env.observe())
            arng = np.random.RandomState(700 + seed)
            acts = arng.uniform([0.0, -0.15], [1.0, 0.15], size=(Tm, 2)).astype(np.float32).astype(np.float64)
            rec = {k: [] for k in ("obs", "reward", "done", "collision", "dists", "n_nearby")}
            n = 0
            for a in acts:
                o, r_, d_, info = env.step(np.array(a))
                n += 1
                rec["obs"].append(np.array(o))
                rec["reward"].append(float(r_))
                rec["done"].append(bool(d_))
                rec["collision"].append(bool(info["collision"]))
                rec["dists"].append(np.array(env.vessel._last_sensor_dist_measurements, dtype=np.float64))
                rec["n_nearby"].append(len(env.vessel._nearby_obstacles))
                if d_:
                    break
        w = np.full((2, 16), np.nan)
        w[:, : env.path.init_waypoints.shape[1]] = env.path.init_waypoints
        mo_out["waypoints"].append(w)
        mo_out["vessel_init"].append(v0)
        mo_out["mov_start"].append(start)
        mo_out["mov_width"].append(width)
        mo_out["vel"].append(vel)
        mo_out["st_pos"].append(spos)
        mo_out["st_radius"].append(srad)
        mo_out["obs0"].append(obs0)
        mo_out["obs"].append(pad(rec["obs"], (Tm, 186)))
        mo_out["dists"].append(pad(rec["dists"], (Tm, 180)))
        mo_out["reward"].append(pad(rec["reward"], (Tm,)))
        for k in ("done", "collision", "n_nearby"):
            mo_out[k].append(pad(np.array(rec[k], dtype=np.float64), (Tm,)))
        mo_out["actions"].append(acts)
        mo_out["T"].append(n)
        print("MovingObstaclesNoRules seed", seed, "steps", n, "min range %.2f" % np.min(rec["dists"]), "nearby",
              sorted(set(rec["n_nearby"])))
    # ------------------------------------------------------------------ the deterministic test scenarios
    # envs/testscenario.py classes instantiated as they are (TestHeadOn draws its start angle from the
    # global, unseeded `random` module -- testscenario.py:145 -- so its entries differ from run to run;
    # the tests read the angle back from the fixture); the scenario each one builds is read
    # back (path, vessel start, obstacle circles / tracks with their velocity tables) and a short
    # episode of its step() is recorded.  DebugScenario draws from env.rng -> seeded.
    ts = importlib.import_module("gym_auv.envs.testscenario")
    names = ["TestScenario1", "TestScenario2", "TestScenario3", "TestScenario4", "TestHeadOn", "TestCrossing",
             "TestCrossing1", "EmptyScenario", "DebugScenario"]
    Tt = 25
    for name in names:
        with contextlib.redirect_stdout(io.StringIO()):
            # The FIRST episode (the constructor's reset()) is what is recorded: a second reset() would
            # list every circle of TestScenario1-4 twice (they never clear self.obstacles) and would
            # leave the rewarder bound to the first episode's vessel (these classes never reassign
            # self.rewarder; environment.py:222-228 only creates one when it is None).  DebugScenario
            # draws from env.rng, which the constructor seeds from the OS: seed 0 is injected instead.
            seeding = sys.modules["gym.utils.seeding"]
            keep = seeding.np_random
            seeding.np_random = lambda seed=None: (np.random.RandomState(0 if seed is None else seed), seed)
            try:
                env = getattr(ts, name)(env_config(1.0), test_mode=True, renderer=None)
            finally:
                seeding.np_random = keep
            mov = [o for o in env.obstacles if not o.static]
            sta = [o for o in env.obstacles if o.static]
            arng = np.random.RandomState(800)
            acts = arng.uniform([0.2, -0.1], [1.0, 0.1], size=(Tt, 2)).astype(np.float32).astype(np.float64)
            v0 = np.array(env.vessel._state[:3])
            obs0 = np.array(env.observe())
            obs, rews, dists, dones = [], [], [], []
            for a in acts:
                o, r_, d_, info = env.step(np.array(a))
                obs.append(np.array(o))
                rews.append(float(r_))
                dists.append(np.array(env.vessel._last_sensor_dist_measurements, dtype=np.float64))
                dones.append(bool(d_))
                if d_:
                    break
        key = "ts_" + name + "_"
        out[key + "waypoints"] = np.array(env.path.init_waypoints, dtype=np.float64)
        out[key + "vessel_init"] = v0
        out[key + "st_pos"] = np.array([o.position for o in sta], dtype=np.float64).reshape(len(sta), 2)
        out[key + "st_radius"] = np.array([o.radius for o in sta], dtype=np.float64)
        out[key + "mov_width"] = np.array([o.width for o in mov], dtype=np.float64)
        out[key + "mov_start"] = np.array([o.trajectory[0][1] for o in mov], dtype=np.float64).reshape(len(mov), 2)
        out[key + "mov_nvel"] = np.array([len(o.trajectory_velocities) for o in mov], dtype=np.int64)
        out[key + "mov_vel_head"] = (np.array([np.array(o.trajectory_velocities[:8], dtype=np.float64) for o in mov])
                                     if mov else np.zeros((0, 8, 2)))
        out[key + "obs0"], out[key + "obs"], out[key + "reward"] = obs0, np.array(obs), np.array(rews)
        out[key + "dists"], out[key + "done"], out[key + "actions"] = np.array(dists), np.array(dones), acts
        print(name, "static", len(sta), "moving", len(mov), "steps", len(obs), "min range %.2f" % np.min(dists))
    # ------------------------------------------------------------------ PathFollowNoObstacles (BASELINE config 2)
    # the reference's own class (movingobstacles.py:114-120: no obstacles, PathFollowRewarder) with
    # use_lidar=False: random curve, 6-dimensional observation
    pf = {k: [] for k in ("waypoints", "vessel_init", "obs0", "obs", "reward", "done", "actions", "T")}
    for seed in range(3):
        with contextlib.redirect_stdout(io.StringIO()):
            cfg_pf = env_config(1.0)
            cfg_pf.vessel.use_lidar = False
            env = mo.PathFollowNoObstacles(cfg_pf, test_mode=True, renderer=None)
            np.random.seed(seed)
            env.seed(seed)
            env.reset()
            assert env.obstacles == [] and isinstance(env.rewarder, rew.PathFollowRewarder)
            v0 = np.array(env.vessel._state[:3])
            obs0 = np.array(env.observe())
            arng = np.random.RandomState(900 + seed)
            acts = arng.uniform([0.0, -0.15], [1.0, 0.15], size=(Tm, 2)).astype(np.float32).astype(np.float64)
            obs, rews, dones = [], [], []
            for a in acts:
                o, r_, d_, info = env.step(np.array(a))
                obs.append(np.array(o))
                rews.append(float(r_))
                dones.append(bool(d_))
                if d_:
                    break
        w = np.full((2, 16), np.nan)
        w[:, : env.path.init_waypoints.shape[1]] = env.path.init_waypoints
        pf["waypoints"].append(w)
        pf["vessel_init"].append(v0)
        pf["obs0"].append(obs0)
        pf["obs"].append(pad(obs, (Tm, 6)))
        pf["reward"].append(pad(rews, (Tm,)))
        pf["done"].append(pad(np.array(dones, dtype=np.float64), (Tm,)))
        pf["actions"].append(acts)
        pf["T"].append(len(obs))
        print("PathFollowNoObstacles seed", seed, "steps", len(obs))
    out.update({"pf_" + k: v for k, v in pf.items()})
    out.update({"mo_" + k: v for k, v in mo_out.items()})
    out.update({"gen_" + k: np.array(v) for k, v in gen_stats.items()})

    np.savez_compressed(os.path.join(HERE, "reference_hybrid.npz"), **{k: np.asarray(v) for k, v in out.items()})
    print("wrote reference_hybrid.npz")


if __name__ == "__main__":
    main()
