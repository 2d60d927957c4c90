"""Real-world scenarios (gym_auv/envs/realworld.py:24-337): static land from obstacle-perimeter
files, other vessels from AIS tracks, a fixed path -- as ``ScenarioSet`` pools for the batched step.

The reference's data files (``resources/obstacles_*.npy``, ``vessel_data*.csv``, ``terrain.npy``) are
not shipped with it; these loaders read the same formats:

* ``load_obstacle_perimeters``: the ``.npy`` object array of [n, 2] vertex lists that
  ``RealWorldEnv._generate`` turns into ``PolygonObstacle``s (realworld.py:141-147; lists with 3 or
  fewer vertices are dropped).  Any perimeter length is supported (k_lidar walks long perimeters
  chain by chain).
* ``vessel_trajectories_from_ais``: the AIS preprocessing of realworld.py:33-117, restated row by row
  (same draws from the env's RNG, same quirks -- a track is only emitted when a gap longer than 0.1 day
  follows it, rows shorter than 12 or faster / slower than the speed window restart the path without
  advancing the "last" row), checked against the reference's own code on synthetic files
  (tests/golden/make_reference_goldens_realworld.py).
* ``real_world_scenarios``: path + polygons + table-driven vessel tracks -> a pool of M identical
  scenarios (every env of a batch runs the same world; the reference builds one env per process).

``SCENARIOS`` holds the reference's four named worlds (paths and file names of realworld.py:243-330);
they raise ``FileNotFoundError`` until the data files are placed in ``resources_dir``.
"""
from __future__ import annotations

import os
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import scenarios as S

VESSEL_SPEED_RANGE_LOWER = 0.1  # realworld.py:21-22
VESSEL_SPEED_RANGE_UPPER = 2

Trajectory = Tuple[float, List[Tuple[int, Tuple[float, float]]], str]  # (width, [(t, (x, y)), ...], name)


def load_obstacle_perimeters(source) -> List[np.ndarray]:
    """``np.load("obstacles_*.npy")`` -> the perimeters that become PolygonObstacles (len > 3)."""
    arr = np.load(source, allow_pickle=True) if isinstance(source, (str, os.PathLike)) else source
    out = []
    for p in arr:
        p = np.asarray(p, dtype=np.float64)
        if len(p) > 3:  # realworld.py:143
            out.append(p.reshape(-1, 2))
    return out


def vessel_trajectories_from_ais(source, rng, n_vessels: int, x0: float = 0.0, y0: float = 0.0) -> List[Trajectory]:
    """realworld.py:33-124.  ``source``: csv path or DataFrame with the columns Vessel_Name,
    AIS_Timestamp, AIS_East, AIS_North, AIS_Length_Overall; ``rng``: the env's RandomState."""
    import pandas as pd

    df = pd.read_csv(source) if isinstance(source, (str, os.PathLike)) else source
    vessels = dict(tuple(df.groupby("Vessel_Name")))
    names = sorted(vessels.keys())
    trajectories: List[Trajectory] = []
    cutoff = pd.to_timedelta(0.1, unit="D")
    while len(trajectories) < n_vessels:
        if len(names) == 0:
            break
        name = names.pop(rng.randint(0, len(names)))
        v = vessels[name].copy()
        v["AIS_Timestamp"] = pd.to_datetime(v["AIS_Timestamp"])
        v["AIS_Timestamp"] -= v.iloc[0]["AIS_Timestamp"]
        start_ts = None
        last_ts = pd.to_timedelta(0, unit="D")
        last_e = last_n = None
        path: list = []
        for _, row in v.iterrows():
            east, north = row["AIS_East"] / 10.0, row["AIS_North"] / 10.0
            if row["AIS_Length_Overall"] < 12:
                continue
            if len(path) == 0:
                start_ts = row["AIS_Timestamp"]
            delta = row["AIS_Timestamp"] - last_ts
            if delta < cutoff:
                if last_e is not None:
                    dist = np.sqrt((east - last_e) ** 2 + (north - last_n) ** 2)
                    with np.errstate(divide="ignore", invalid="ignore"):
                        speed = np.float64(dist) / delta.seconds
                    if speed < VESSEL_SPEED_RANGE_LOWER or speed > VESSEL_SPEED_RANGE_UPPER:
                        path = []
                        continue
                path.append((int((row["AIS_Timestamp"] - start_ts).total_seconds()), (east - x0, north - y0)))
            else:
                if len(path) > 1 and not np.isnan(row["AIS_Length_Overall"]) and row["AIS_Length_Overall"] > 0:
                    start = rng.randint(0, len(path) - 1)
                    trajectories.append((row["AIS_Length_Overall"] / 10.0, path[start:], name))
                path = []
            last_ts, last_e, last_n = row["AIS_Timestamp"], east, north
    pick = rng.choice(list(range(len(trajectories))), min(len(trajectories), n_vessels), replace=False)
    return [trajectories[i] for i in pick]


def real_world_scenarios(path_waypoints, obstacle_perimeters: Sequence[np.ndarray] = (), vessel_trajectories: Sequence[Trajectory] = (),
                         n_scenarios: int = 1, rewarder: str = "colav", name: str = "RealWorld") -> S.ScenarioSet:
    """``RealWorldEnv._generate`` (realworld.py:119-170) as a pool of ``n_scenarios`` identical
    scenarios: vessel at path(0) heading along it, one PolygonObstacle per perimeter, one
    VesselObstacle(width=int(w)) per trajectory with more than 2 points (table-driven track), and the
    closing ``_update()``."""
    moving = [(int(w), list(traj)) for w, traj, _ in vessel_trajectories if len(traj) > 2]
    one = S._single(np.asarray(path_waypoints, dtype=np.float64), moving_traj=moving, rewarder=rewarder, post_update=True,
                    name=name)
    one.world_polygons = [np.asarray(p, dtype=np.float64) for p in obstacle_perimeters]
    if n_scenarios == 1:
        return one
    M = int(n_scenarios)
    rep = lambda a: np.repeat(a, M, axis=0)
    out = S.ScenarioSet(
        waypoints=one.waypoints, path_id=np.zeros(M, dtype=np.int32), vessel_init=rep(one.vessel_init),
        mov_start=rep(one.mov_start), mov_width=rep(one.mov_width), mov_track=rep(one.mov_track), vel_table=one.vel_table,
        st_pos=rep(one.st_pos), st_radius=rep(one.st_radius), rewarder=rewarder, post_generate_update=True, name=name,
        world_polygons=one.world_polygons, path_group=M, path_period=M)
    out._bank = one.bank
    return out


def _named(path_wp, obstacles_file, vessel_file, n_vessels, x0, y0):
    def build(resources_dir: str, rng=None, n_scenarios: int = 1) -> S.ScenarioSet:
        rng = rng or np.random.RandomState(0)
        obst = os.path.join(resources_dir, obstacles_file)
        ais = os.path.join(resources_dir, vessel_file)
        for f in (obst, ais):
            if not os.path.exists(f):
                raise FileNotFoundError(f"{f}: the reference's data files are not shipped with it (gym_auv/envs/realworld.py)")
        traj = vessel_trajectories_from_ais(ais, rng, n_vessels, x0, y0)
        return real_world_scenarios(path_wp, load_obstacle_perimeters(obst), traj, n_scenarios, name=obstacles_file)

    return build


# realworld.py:243-330 (paths in the env's local frame, file names, vessel counts, frame offsets)
SCENARIOS = {
    "Sorbuoya-v0": _named([[1000, 830, 700, 960, 1080, 1125], [910, 800, 700, 550, 750, 810]], "obstacles_sorbuoya.npy",
                          "vessel_data_local_sorbuoya.csv", 25, 0, 10000),
    "Agdenes-v0": _named([[4100 - 3121, 4247 - 3121, 4137 - 3121, 3937 - 3121, 3217 - 3121],
                          [6100 - 5890, 6100 - 5890, 6860 - 5890, 6910 - 5890, 6690 - 5890]], "obstacles_entrance.npy",
                         "vessel_data_local_agdenes.csv", 15, 3121, 5890),
    "Trondheim-v0": _named([[6945 - 5000, 6329 - 5000], [4254 - 3900, 5614 - 3900]], "obstacles_trondheim.npy",
                           "vessel_data_local_trondheim.csv", 100, 5000, 3900),
    "Trondheimsfjorden-v0": _named([[520, 1070, 4080, 5473, 10170, 12220], [3330, 5740, 7110, 4560, 7360, 11390]],
                                   "obstacles_trondheimsfjorden.npy", "vessel_data.csv", 999999, 0, 0),
}
