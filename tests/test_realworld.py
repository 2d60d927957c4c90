"""Real-world scenario loaders (gym_auv_b200/realworld.py) against goldens produced by the
reference's own RealWorldEnv (envs/realworld.py:24-240) on synthetic data files in the reference's
formats (tests/golden/make_reference_goldens_realworld.py; the real files are not shipped upstream):
the AIS preprocessing selects the same vessel tracks, the perimeter list gives the same polygons,
and the oracle replays the reference's episode on the scenario the loaders build.  The CUDA step is
compared with the same golden in tests/test_gpu_v2.py."""
import os

import numpy as np

from gym_auv_b200 import lidar_config
from gym_auv_b200 import realworld as RW
from tests._parity import oracle_cfg

HERE = os.path.dirname(os.path.abspath(__file__))
DATA = os.path.join(HERE, "golden", "realworld_data")
GOLD = np.load(os.path.join(HERE, "golden", "reference_realworld.npz"))
PATH_WP = [[100, 300, 500, 800], [150, 400, 450, 800]]


def build_scenario(n=1):
    traj = RW.vessel_trajectories_from_ais(os.path.join(DATA, "vessel_data_synthetic.csv"), np.random.RandomState(5), 4, 0, 10000)
    per = RW.load_obstacle_perimeters(os.path.join(DATA, "obstacles_synthetic.npy"))
    return RW.real_world_scenarios(PATH_WP, per, traj, n_scenarios=n), traj, per


def test_ais_preprocessing_and_perimeters_equal_the_reference():
    scn, traj, per = build_scenario()
    assert len(traj) == int(GOLD["n_vessels"])
    for i, (w, t, name) in enumerate(traj):
        ref = GOLD["traj"][i]
        ref = ref[~np.isnan(ref[:, 0])]
        got = np.array([(a, b[0], b[1]) for a, b in t])
        assert name == str(GOLD["names"][i]) and w == float(GOLD["widths"][i])
        assert got.shape == ref.shape and np.array_equal(got, ref)
    assert len(per) == int(GOLD["n_polygons"]) and max(len(p) for p in per) > 192  # longer than one vertex stage
    assert scn.k_moving == int(GOLD["n_vessel_obstacles"]) and len(scn.world_polygons) == len(per)
    assert np.abs(scn.vessel_init[0] - GOLD["vessel_init"]).max() <= 1e-12


def test_oracle_replays_the_reference_realworld_episode():
    from oracle.sim import OracleEnv

    scn, _, _ = build_scenario()
    env = OracleEnv(scn.describe(0), oracle_cfg(lidar_config()), test_mode=True)
    obs0 = env.reset()
    assert np.abs(obs0 - GOLD["obs0"]).max() <= 1e-9
    for t, a in enumerate(GOLD["actions"][: len(GOLD["obs"])]):
        obs, rew, done, info = env.step(a)
        assert len(env.vessel.nearby) == int(GOLD["n_nearby"][t])
        assert np.abs(env.vessel.dists - GOLD["dists"][t]).max() <= 1e-9, t
        assert np.abs(obs - GOLD["obs"][t]).max() <= 1e-9 and abs(rew - GOLD["reward"][t]) <= 1e-9 * max(1, abs(rew))
        assert done == bool(GOLD["done"][t])


def test_named_worlds_need_the_reference_data_files(tmp_path):
    import pytest

    assert set(RW.SCENARIOS) == {"Sorbuoya-v0", "Agdenes-v0", "Trondheim-v0", "Trondheimsfjorden-v0"}
    with pytest.raises(FileNotFoundError):
        RW.SCENARIOS["Sorbuoya-v0"](str(tmp_path))
