"""The oracle's velocity channel (velocity_mode="nearest") against goldens produced by the
reference's own ``simulate_sensor_brute_force`` (sensor.py:100-137) and ``ColavRewarder.calculate``
(rewarder.py:167-241) -- tests/golden/make_reference_goldens_velocity.py, reference classes on the
geos_lite primitives."""
import math
import os
import types

import numpy as np

from oracle import sim as O

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_velocity.npz"))


def _scene_objects(row):
    p0 = row[:2]
    mov = row[2:29].reshape(3, 9)
    st = row[29:35].reshape(2, 3)
    order = row[35:40].astype(int)
    objs = []
    for width, sx, sy, vx, vy, px, py, dx, dy in mov:
        ob = O.OracleVesselObstacle(width, (sx, sy), np.tile([vx, vy], (49, 1)), init_update=True)
        ob.position = np.array([px, py])
        ob.dx, ob.dy = dx, dy
        ob.heading = math.atan2(dy, dx)
        ob._rebuild()
        objs.append(ob)
    for cx, cy, r in st:
        objs.append(O.OracleCircle((cx, cy), r))
    return p0, [objs[k] for k in order]


def test_brute_force_sensor_distance_and_relative_speed():
    scenes = {}
    n_moving_hits = 0
    for (s, angle), (d, vx, vy, blocked) in zip(GOLD["vel_rays"], GOLD["vel_res"]):
        s = int(s)
        if s not in scenes:
            scenes[s] = _scene_objects(GOLD["vel_scenes"][s])
        p0, cand = scenes[s]
        got_d, who = O.cast_ray(angle, (float(p0[0]), float(p0[1])), 150.0, cand, with_obstacle=True)
        gvx, gvy = O.relative_speed(angle, who)
        assert abs(got_d - d) <= 1e-9, (s, angle, got_d, d)
        assert bool(blocked) == (who is not None)
        assert abs(gvx - vx) <= 1e-12 and abs(gvy - vy) <= 1e-12, (s, angle, (gvx, gvy), (vx, vy))
        n_moving_hits += int(abs(vx) + abs(vy) > 0)
    assert n_moving_hits > 100  # the goldens do exercise moving obstacles


def test_colav_reward_with_speed_measurements():
    R = 180
    angles = np.array([-np.pi + (i + 1) * 2 * np.pi / R for i in range(R)])
    for row, want in zip(GOLD["velrew_in"], GOLD["velrew_out"]):
        speed, yaw, cte, he, prog, maxprog = row[:6]
        d, sp = row[6:6 + R], row[6 + R:].reshape(2, R)
        v = types.SimpleNamespace(
            collision=False, nav=dict(cross_track_error=cte, heading_error=he), speed=speed, n_sensors=R,
            sensor_angles=angles, dists=d, speeds=sp, cfg=dict(sensor_range=150.0), progress=prog, max_progress=maxprog,
            state=np.array([0, 0, 0, 0, 0, yaw]))
        assert abs(O.colav_reward(v) - want) <= 1e-9 * max(1.0, abs(want))
    zero = GOLD["velrew_in"][:, 6 + R:] == 0
    assert not zero.all()
