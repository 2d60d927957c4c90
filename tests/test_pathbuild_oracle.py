"""oracle/pathbuild.py (the operation-by-operation restatement the CUDA path builder follows)
pinned against SciPy's PchipInterpolator as the reference's Path.__init__ uses it (path.py:19-40)
and against the reference's own Path class run behind stubs (tests/golden/reference_stubbed.npz)."""
import os

import numpy as np

from gym_auv_b200.pathbank import build_path, random_curve_waypoints
from oracle import pathbuild as PB


def _cases():
    rng = np.random.RandomState(3)
    cases = [random_curve_waypoints(rng, int(np.floor(4 * rng.rand() + 2)), 800.0) for _ in range(12)]
    cases.append(np.array([[0.0, 1100.0], [0.0, 1100.0]]))          # TestScenario1: two waypoints
    cases.append(np.array([[25.0, 25.0], [10.0, 200.0]]))            # EmptyScenario
    cases.append(np.array([[0.0, 300.0, 300.0, 900.0], [0.0, 0.0, 400.0, 400.0]]))  # right angles: zero slopes
    return cases


def test_restatement_equals_scipy_to_the_last_bits():
    for wp in _cases():
        ref = build_path(wp)  # SciPy
        got = PB.build(wp)
        assert abs(got["length"] - ref.length) <= 1e-12 * ref.length
        assert np.abs(got["knots"] - ref.knots).max() <= 1e-12 * ref.length
        scale = np.abs(ref.coef).max(axis=(0, 2))
        assert np.abs(got["cx"] - ref.coef[:, 0, :]).max() <= 1e-9 * max(1.0, scale[0])
        assert np.abs(got["cy"] - ref.coef[:, 1, :]).max() <= 1e-9 * max(1.0, scale[1])
        assert got["poly"].shape == ref.poly.shape
        assert np.abs(got["poly"] - ref.poly).max() <= 1e-10
        assert np.abs(got["cum"] - ref.cum).max() <= 1e-9


def test_restatement_equals_the_reference_path_class():
    """The reference's own Path / RandomCurveThroughOrigin classes (path.py:19-120), run behind import
    stubs by tests/golden/make_reference_goldens_stubbed.py: length, the 1000 knots, positions along
    the path and the head of the 0.1 m polyline."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_stubbed.npz"))
    for k in range(len(g["path_length"])):
        wp = g["path_waypoints"][k]
        wp = wp[:, ~np.isnan(wp[0])]
        got = PB.build(wp)
        L = float(g["path_length"][k])
        assert abs(got["length"] - L) <= 1e-12 * L
        assert np.abs(got["knots"] - g["path_knots"][k]).max() <= 1e-12 * L
        assert len(got["poly"]) == int(g["path_poly_n"][k])
        head = g["path_poly_sample"][k]  # path.points[:: max(1, len // 64)][:64]
        assert np.abs(got["poly"][:: max(1, len(got["poly"]) // 64)][:64] - head).max() <= 1e-10
        S = g["path_S"] * L if g["path_S"].max() <= 1.0 + 1e-12 else g["path_S"]
        pos = np.stack([PB.evaluate(got["knots"], got["cx"], S), PB.evaluate(got["knots"], got["cy"], S)], axis=1)
        assert np.abs(pos - g["path_pos"][k]).max() <= 1e-9
