"""Golden vectors for the LiDAR VELOCITY channel from the reference's own code.

At HEAD the live ``simulate_sensor`` returns (0, 0) speeds (sensor.py:140-159); the obstacle
velocity per ray is computed by ``simulate_sensor_brute_force`` (sensor.py:100-137), which the
product offers as ``velocity_mode="nearest"``.  This script runs -- unmodified, behind the same
import stubs as make_reference_goldens_hybrid.py (the reference's classes on oracle/geos_lite
primitives; Shapely itself is not installable here) --

  * ``sensor.simulate_sensor_brute_force`` on random rays against the reference's own
    ``VesselObstacle`` / ``CircularObstacle`` objects: measured distance, the relative speed vector
    Rz(-angle - pi/2) (dx, dy) of the nearest hit obstacle, blocked flag;
  * ``ColavRewarder.calculate`` (rewarder.py:167-241) with NON-ZERO speed measurements, i.e. the
    ``max(0, v_y)`` term of the closeness penalty (rewarder.py:199-206).

Usage:  python tests/golden/make_reference_goldens_velocity.py   (writes reference_velocity.npz)
"""
import importlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import make_reference_goldens_hybrid as hybrid  # noqa: E402
import make_reference_goldens_stubbed as stubbed  # noqa: E402


def main():
    hybrid.install()
    sensor = importlib.import_module("gym_auv.objects.vessel.sensor")
    obst = importlib.import_module("gym_auv.objects.obstacles")
    vesselm = importlib.import_module("gym_auv.objects.vessel.vessel")
    sys.modules["gym_auv.objects.vessel"].Vessel = vesselm.Vessel
    rew = importlib.import_module("gym_auv.objects.rewarder")
    Point = sys.modules["shapely.geometry"].Point
    rng = np.random.RandomState(7)
    out = {}

    # ---- simulate_sensor_brute_force: scenes of 3 vessels + 2 circles around the own-ship
    scenes, rays, res = [], [], []
    for s in range(40):
        p0 = rng.uniform(-50, 50, 2)
        mov = []
        objs = []
        for j in range(3):
            ang, dist = rng.uniform(0, 2 * np.pi), rng.uniform(15, 120)
            start = p0 + dist * np.array([np.cos(ang), np.sin(ang)])
            vel = rng.uniform(1, 3) * np.array([np.cos(a := rng.uniform(0, 2 * np.pi)), np.sin(a)])
            width = float(max(1, rng.poisson(10)))
            traj = [[i, tuple(start + i * vel)] for i in range(50)]
            ob = obst.VesselObstacle(width=width, trajectory=traj)
            for _ in range(int(rng.randint(0, 4))):
                ob.update(1.0)
            objs.append(ob)
            mov.append([width, start[0], start[1], vel[0], vel[1], ob.position[0], ob.position[1], ob.dx, ob.dy])
        st = []
        for j in range(2):
            ang, dist = rng.uniform(0, 2 * np.pi), rng.uniform(40, 140)
            c = p0 + dist * np.array([np.cos(ang), np.sin(ang)])
            r = float(max(1, rng.poisson(30)))
            objs.append(obst.CircularObstacle(c, r))
            st.append([c[0], c[1], r])
        order = rng.permutation(5)  # candidate-list order decides ties
        cand = [objs[k] for k in order]
        for k in range(24):
            angle = rng.uniform(-np.pi, np.pi)
            if k % 2 == 0:  # aim at an obstacle
                tgt = objs[int(rng.randint(5))]
                c = np.array(tgt.position, dtype=float).flatten()
                angle = np.arctan2(c[1] - p0[1], c[0] - p0[0]) + rng.normal(0, 0.08)
            d, v, blocked = sensor.simulate_sensor_brute_force(angle, Point(*p0), 150.0, cand)
            rays.append([s, angle])
            res.append([float(d), float(v[0]), float(v[1]), float(blocked)])
        scenes.append(np.hstack([p0, np.array(mov).ravel(), np.array(st).ravel(), order]))
    out.update(vel_scenes=np.array(scenes), vel_rays=np.array(rays), vel_res=np.array(res))

    # ---- ColavRewarder with speed measurements
    R = 180
    angles = np.array([-np.pi + (i + 1) * 2 * np.pi / R for i in range(R)])
    cfg = stubbed.make_config()
    rin, rout = [], []
    for k in range(120):
        d = np.full(R, 150.0)
        n_hit = int(rng.randint(0, 60))
        idx = rng.choice(R, n_hit, replace=False)
        d[idx] = rng.uniform(0.5, 150.0, n_hit)
        sp = np.zeros((2, R))
        sp[:, idx] = rng.uniform(-3, 3, (2, n_hit))
        speed, yaw = rng.uniform(0, 0.6), rng.uniform(-0.1, 0.1)
        cte, he = rng.uniform(-1, 1), rng.uniform(-np.pi, np.pi)
        prog = rng.uniform(0, 1)
        maxprog = max(prog, rng.uniform(0, 1)) if k % 2 else prog
        fake = stubbed.ns(
            req_latest_data=lambda d=d, sp=sp, cte=cte, he=he: {
                "navigation": {"cross_track_error": cte, "heading_error": he}, "distance_measurements": d,
                "speed_measurements": sp, "collision": False},
            speed=speed, max_speed=2, yaw_rate=yaw, n_sensors=R, sensor_angles=angles, config=cfg, progress=prog,
            max_progress=maxprog)
        rc = rew.ColavRewarder(fake, test_mode=True)
        rin.append(np.hstack([speed, yaw, cte, he, prog, maxprog, d, sp[0], sp[1]]))
        rout.append(float(rc.calculate()))
    out.update(velrew_in=np.array(rin), velrew_out=np.array(rout))
    np.savez_compressed(os.path.join(HERE, "reference_velocity.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
