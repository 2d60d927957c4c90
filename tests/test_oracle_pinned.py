"""The oracle's dynamics / angle wrap / sector map against golden vectors produced by the
REFERENCE'S OWN files (tests/golden/make_reference_goldens.py).  This is the pinned
part of the oracle."""
import os

import numpy as np
import pytest

from oracle import model as M

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_dynamics.npz"))


def test_model_matrices_match_reference_constants():
    assert np.allclose(M.MASS_INV, G["M_inv"], rtol=0, atol=1e-17)
    assert np.array_equal(M.DAMP, G["D"])
    assert np.array_equal(M.nonlinear_damping(np.array([1.0, 0.0, 0.0])), G["N_u1"])


def test_rkf45_rollouts_match_reference_odesolver():
    worst = 0.0
    for init, acts, dt, traj in zip(G["inits"], G["actions"], G["dts"], G["trajs"]):
        s = np.hstack([init, np.zeros(3)])
        for t in range(len(acts)):
            s = M.vessel_step(s, acts[t], float(dt))
            worst = max(worst, float(np.abs(s - traj[t + 1]).max()))
    assert worst <= 1e-13, worst


def test_survey_probe_values():
    # SURVEY.md App. A.1 sanity values, reproduced from the reference's files
    for h, key in ((1.0, "probe_h1"), (0.5, "probe_h05")):
        s = np.array([0, 0, 0.3, 0, 0, 0], dtype=np.float64)
        for _ in range(5):
            s = M.vessel_step(s, [0.5, 0.6], h)
        assert np.abs(s - G[key]).max() <= 1e-15
    assert abs(G["probe_h1"][0] - 0.350372991) < 1e-9


def test_princip_matches_reference():
    out = np.array([M.princip(x) for x in G["princip_in"]])
    assert np.array_equal(out, G["princip_out"])
    assert M.princip(np.pi) == -np.pi  # floored modulo => [-pi, pi)


@pytest.mark.parametrize("n,s,key", [(180, 9, "sectors_180"), (128, 8, "sectors_128")])
def test_sector_map_bit_exact(n, s, key):
    assert np.array_equal(M.sector_table(n, s), G[key])


def test_sector_counts_from_survey():
    counts = np.bincount(M.sector_table(180, 9))
    assert counts.tolist() == [54, 15, 10, 8, 8, 9, 10, 15, 51]


def test_product_sector_and_ray_tables_match_oracle():
    pytest.importorskip("torch")
    from gym_auv_b200.vec_env import ray_table

    ang, cos_sin, weight, wsum, sector = ray_table(180, 9)
    assert np.array_equal(sector, G["sectors_180"].astype(np.uint8))
    ref_ang = np.array([-np.pi + (i + 1) * (2 * np.pi / 180) for i in range(180)])
    assert np.array_equal(ang, ref_ang)
    assert abs(ang[89]) < 1e-12 and abs(ang[179] - np.pi) < 1e-12  # ray 89 dead ahead, 179 astern
