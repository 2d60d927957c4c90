"""Goldens from the reference's RealWorldEnv (envs/realworld.py:24-240) on SYNTHETIC data files.

The reference's terrain / AIS files are not shipped with it, so this script
writes small synthetic stand-ins in the same formats -- an AIS csv with the columns realworld.py:34-110
reads (Vessel_Name, AIS_Timestamp, AIS_East, AIS_North, AIS_Length_Overall) and an obstacle-perimeter
.npy (object array of [n, 2] vertex lists, realworld.py:141-147) -- and runs the reference's own
``RealWorldEnv._generate`` / ``step`` on them behind the import stubs of make_reference_goldens_hybrid.py
(reference classes on oracle/geos_lite primitives; pandas is real).  Recorded: the vessel
trajectories the AIS preprocessing selects (width, name, (t, (x, y)) list), the polygons kept, and
a 40-step episode.

Usage:  python tests/golden/make_reference_goldens_realworld.py   (writes reference_realworld.npz + the data files)
"""
import contextlib
import importlib
import io
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import make_reference_goldens_hybrid as hybrid  # noqa: E402
import make_reference_goldens_stubbed as stubbed  # noqa: E402

DATA = os.path.join(HERE, "realworld_data")


def write_synthetic_data():
    import pandas as pd

    os.makedirs(DATA, exist_ok=True)
    rng = np.random.RandomState(11)
    rows = []
    t0 = pd.Timestamp("2020-01-01 00:00:00")
    for v in range(6):
        name = f"VESSEL_{v}"
        length = [180.0, 90.0, 250.0, 60.0, 11.0 * 10, 300.0][v]  # decimetres; vessel 4 has rows shorter than 12
        east, north = rng.uniform(2000, 9000), rng.uniform(101000, 109000)  # decimetres (realworld.py:56-57 divides by 10)
        if v in (0, 1):  # close to the start of the path, so that the own-ship's LiDAR sees them
            east, north = 2300.0 + 900.0 * v, 102800.0 - 500.0 * v
        heading = rng.uniform(0, 2 * np.pi)
        speed = rng.uniform(5, 15)  # dm/s -> 0.5 .. 1.5 m/s, inside VESSEL_SPEED_RANGE
        t = t0 + pd.Timedelta(seconds=int(rng.randint(0, 600)))
        for k in range(40):
            if (v in (0, 1, 2, 5) and k == 25) or (v == 1 and k == 33):  # gaps longer than cutoff_dt (0.1 day): a track is emitted
                t = t + pd.Timedelta(hours=5)
            if v == 3 and k == 12:  # a jump faster than 2 m/s: the path restarts
                east += 5000.0
            dt = int(rng.randint(20, 60))
            t = t + pd.Timedelta(seconds=dt)
            heading += rng.normal(0, 0.05)
            east += speed * dt * np.cos(heading)
            north += speed * dt * np.sin(heading)
            ln = 8.0 if (v == 4 and k % 3 == 0) else length
            rows.append(dict(Vessel_Name=name, AIS_Timestamp=t.strftime("%Y-%m-%d %H:%M:%S"), AIS_East=east, AIS_North=north,
                             AIS_Length_Overall=ln))
    pd.DataFrame(rows).to_csv(os.path.join(DATA, "vessel_data_synthetic.csv"), index=False)
    polys = []
    for k in range(9):
        c = np.array([rng.uniform(100, 900), rng.uniform(100, 900)])
        n = int(rng.choice([3, 5, 9, 40, 230]))  # 3-vertex lists are skipped (len > 3 only); 230 > one vertex stage
        if k < 4:  # around the first leg of the path (100, 150) -> (300, 400)
            c = np.array([(50.0, 215.0), (215.0, 150.0), (150.0, 300.0), (330.0, 330.0)][k])
            n = [9, 40, 230, 5][k]
        ang = np.sort(rng.uniform(0, 2 * np.pi, n))
        rad = rng.uniform(20, 70) * rng.uniform(0.6, 1.0, n)
        polys.append(c + rad[:, None] * np.stack([np.cos(ang), np.sin(ang)], axis=1))
    arr = np.empty(len(polys), dtype=object)
    for i, p in enumerate(polys):
        arr[i] = p
    np.save(os.path.join(DATA, "obstacles_synthetic.npy"), arr, allow_pickle=True)


def main():
    write_synthetic_data()
    hybrid.install()
    obst = importlib.import_module("gym_auv.objects.obstacles")
    pathm = importlib.import_module("gym_auv.objects.path")
    vesselm = importlib.import_module("gym_auv.objects.vessel.vessel")
    sys.modules["gym_auv.objects.vessel"].Vessel = vesselm.Vessel
    importlib.import_module("gym_auv.objects.rewarder")
    importlib.import_module("gym_auv.environment")
    envs_pkg = types.ModuleType("gym_auv.envs")
    envs_pkg.__path__ = [os.path.join(stubbed.REF, "envs")]
    sys.modules["gym_auv.envs"] = envs_pkg
    rw = importlib.import_module("gym_auv.envs.realworld")
    Config = sys.modules["gym_auv"].Config

    def env_config():
        c = Config()
        base = stubbed.make_config(dt=1.0)
        c.vessel, c.simulation, c.episode = base.vessel, base.simulation, base.episode
        c.vessel.use_lidar = True
        c.vessel.dense_observation_size = 6
        c.vessel.n_lidar_observations = 180
        c.vessel.use_dict_observation = False
        c.vessel.sensor_use_velocity_observations = False
        c.vessel.sensor_interval_load_obstacles = 25
        c.vessel.render_distance = 300
        c.episode.max_timesteps = 10000
        c.episode.min_cumulative_reward = -2000.0
        return c

    class Synthetic(rw.RealWorldEnv):
        def __init__(self, *a, **kw):
            self.x0, self.y0 = 0, 10000
            self.vessel_data_path = os.path.join(DATA, "vessel_data_synthetic.csv")
            self.n_vessels = 4
            super().__init__(*a, **kw)

        def _generate(self):
            self.path = pathm.Path([[100, 300, 500, 800], [150, 400, 450, 800]])
            self.obstacle_perimeters = np.load(os.path.join(DATA, "obstacles_synthetic.npy"), allow_pickle=True)
            super()._generate()

    with contextlib.redirect_stdout(io.StringIO()):
        env = Synthetic(env_config(), test_mode=True, renderer=None)
        env.seed(5)
        env.rewarder = None  # (upstream keeps the FIRST episode's rewarder -- and with it that episode's vessel -- alive
        #                       across resets, environment.py:219-224; the golden is a first episode)
        obs0 = np.array(env.reset())
    out = {}
    tr = env.other_vessels
    out["n_vessels"] = np.array(len(tr))
    out["widths"] = np.array([w for w, _, _ in tr], dtype=np.float64)
    out["names"] = np.array([n for _, _, n in tr])
    maxlen = max(len(t) for _, t, _ in tr)
    tt = np.full((len(tr), maxlen, 3), np.nan)
    for i, (_, t, _) in enumerate(tr):
        tt[i, : len(t)] = [(a, b[0], b[1]) for a, b in t]
    out["traj"] = tt
    out["n_polygons"] = np.array(sum(isinstance(o, obst.PolygonObstacle) for o in env.obstacles))
    out["n_vessel_obstacles"] = np.array(sum(isinstance(o, obst.VesselObstacle) for o in env.obstacles))
    out["vessel_init"] = np.array(env.vessel._state[:3])
    out["obs0"] = obs0
    rng = np.random.RandomState(3)
    acts = rng.uniform([0.2, -0.15], [1.0, 0.15], size=(40, 2)).astype(np.float32).astype(np.float64)
    rec = {k: [] for k in ("obs", "reward", "done", "dists", "n_nearby")}
    with contextlib.redirect_stdout(io.StringIO()):
        for a in acts:
            o, r, d, info = env.step(np.array(a))
            rec["obs"].append(np.array(o))
            rec["reward"].append(float(r))
            rec["done"].append(bool(d))
            rec["dists"].append(np.array(env.vessel._last_sensor_dist_measurements, dtype=np.float64))
            rec["n_nearby"].append(len(env.vessel._nearby_obstacles))
            if d:
                break
    out["actions"] = acts
    for k, v in rec.items():
        out[k] = np.array(v)
    np.savez_compressed(os.path.join(HERE, "reference_realworld.npz"), **out)
    print({k: getattr(v, "shape", v) for k, v in out.items()})


if __name__ == "__main__":
    main()
