"""Cross-checks that anchor the UNPINNED (GEOS-restated) part of the oracle."""
import math

import numpy as np
import pytest

from oracle import geos_lite as G
from oracle import sim as S


@pytest.mark.parametrize("r", [0.5, 1.0, 1.02, 1.03, 2.5, 3.9, 4.0, 10.0, 15.6, 15.7, 30.0, 62.0, 63.0, 200.0, 838.0])
def test_douglas_peucker_matches_closed_form_ngon(r):
    ring = G.circle_boundary_ring(3.0, -4.0, r)
    n = len(ring) - 1
    assert n == G.circle_ngon_sides(r)
    # regular n-gon inscribed in the circle, a vertex at angle 0
    ang = np.arctan2(ring[:-1, 1] + 4.0, ring[:-1, 0] - 3.0)
    k = ang / (2 * math.pi / n)
    assert np.abs(k - np.round(k)).max() < 1e-9
    assert np.abs(np.hypot(ring[:, 0] - 3.0, ring[:, 1] + 4.0) - r).max() < 1e-9 * max(r, 1)


def test_ngon_thresholds_from_survey_table():
    assert [G.circle_ngon_sides(r) for r in (1.024, 1.025, 3.941, 3.942, 15.61, 15.62, 62.30, 62.31)] == [
        4, 8, 8, 16, 16, 32, 32, 64]


def test_project_vectorised_equals_sequential_and_first_minimum_wins():
    rng = np.random.RandomState(0)
    pts = np.cumsum(rng.normal(size=(300, 2)), axis=0)
    for _ in range(50):
        p = rng.normal(size=2) * 10
        assert G.linestring_project(pts, p) == pytest.approx(G.linestring_project_sequential(pts, p), abs=1e-12)
    # a closed square: the centre is equidistant from all 4 sides -> first segment wins
    sq = np.array([[0, 0], [2, 0], [2, 2], [0, 2], [0, 0]], dtype=float)
    assert G.linestring_project(sq, (1.0, 1.0)) == 1.0
    assert G.linestring_project_sequential(sq, (1.0, 1.0)) == 1.0
    # beyond the ends: clamps to 0 / total length
    line = np.array([[0, 0], [1, 0], [2, 0]], dtype=float)
    assert G.linestring_project(line, (-5.0, 1.0)) == 0.0
    assert G.linestring_project(line, (9.0, 1.0)) == 2.0


def test_ray_ring_numpy_equals_scalar():
    rng = np.random.RandomState(1)
    ring = G.circle_boundary_ring(0, 0, 30.0)
    for _ in range(200):
        p0 = rng.uniform(-80, 80, size=2)
        a = rng.uniform(-math.pi, math.pi)
        p1 = (p0[0] + 150 * math.cos(a), p0[1] + 150 * math.sin(a))
        d1 = G.ray_ring_min_distance(p0, p1, ring)
        d2 = G.ray_ring_min_distance_np(p0, p1, ring)
        assert (d1 is None) == (d2 is None)
        if d1 is not None:
            assert d1 == pytest.approx(d2, abs=1e-12)


def test_polygonised_circle_range_bounded_by_analytic_circle():
    """One-sided cross-check: the n-gon lies inside the circle, edges at most
    0.3 m inside, so polygon range >= circle range, and for rays that clear the band the
    difference stays within what a 0.3 m inset can produce."""
    rng = np.random.RandomState(2)
    for r in (5.0, 30.0, 70.0):
        ring = G.circle_boundary_ring(0, 0, r)
        for _ in range(300):
            d0 = rng.uniform(r + 2, r + 120)
            phi = rng.uniform(-math.pi, math.pi)
            p0 = (d0 * math.cos(phi), d0 * math.sin(phi))
            a = phi + math.pi + rng.uniform(-0.3, 0.3)
            dirv = (math.cos(a), math.sin(a))
            p1 = (p0[0] + 150 * dirv[0], p0[1] + 150 * dirv[1])
            tc = -(p0[0] * dirv[0] + p0[1] * dirv[1])
            h2 = d0 * d0 - tc * tc
            circ = tc - math.sqrt(r * r - h2) if h2 < r * r and tc > 0 else None
            poly = G.ray_ring_min_distance(p0, p1, ring)
            if poly is not None and poly <= 150:
                assert circ is not None, "polygon hit implies circle hit (polygon is inside)"
                if circ <= 150:
                    assert poly >= circ - 1e-9
            if circ is not None and circ <= 150 and h2 < (r - 0.31) ** 2:
                assert poly is not None
                assert poly - circ <= 0.3 / max(math.sqrt(1 - h2 / (r * r)), 1e-3) + 1e-6


def test_vessel_obstacle_geometry_closed_forms():
    """rotation about the area centroid (5w/18, 0); enclosing circle of the
    min-rotated rectangle = centre c + R((w/2,0)-c) + pos, radius w*sqrt(5)/2."""
    for w, vel in ((10.0, (1.3, -2.1)), (1.0, (-0.2, 0.9)), (23.0, (0.0, -3.0))):
        ob = S.OracleVesselObstacle(w, [100.0, 50.0], np.broadcast_to(vel, (9999, 2)))
        assert np.allclose(G.polygon_centroid(ob.body), [5 * w / 18, 0.0], atol=1e-12)
        c, rad = ob.enclosing_circle()
        th = ob.heading
        pred = np.array([5 * w / 18, 0]) + (2 * w / 9) * np.array([math.cos(th), math.sin(th)]) + ob.position
        assert np.allclose(c, pred, atol=1e-9)
        assert rad == pytest.approx(w * math.sqrt(5) / 2, abs=1e-9)
        assert ob.counter == pytest.approx(0.1)
        assert np.allclose(ob.position, np.array([100.0, 50.0]) + 0.1 * np.array(vel))


def test_vessel_obstacle_track_wrap():
    vel = np.array([[1.0, 0.0], [2.0, 0.0], [3.0, 0.0], [4.0, 0.0]])
    ob = S.OracleVesselObstacle(2.0, [0.0, 0.0], vel, init_update=False)
    xs = []
    for _ in range(6):
        ob.update(1.0)
        xs.append(ob.position[0])
    # idx 1,2 then idx 3 >= len-1 -> wrap to start and advance by vel[0]
    assert xs == [2.0, 5.0, 1.0, 3.0, 6.0, 1.0]


# ---- the reference's own LiDAR test (tests/test_hierarchical_collision_detector.py) ----
def _perceive(vessel_state, obstacles, **cfg):
    c = dict(S.DEFAULT_CFG)
    c.update(cfg)
    v = S.OracleVessel(c, vessel_state)
    closeness, _ = v.perceive(obstacles)
    return closeness, v


def test_reference_hierarchical_collision_detector_case():
    closeness, v = _perceive([5, -5, np.deg2rad(45)], [S.OracleCircle([0, -9.5], 1.5)])
    assert not (0 < closeness[len(closeness) // 2] < 1)  # test_no_obst_in_front
    assert 0 < closeness[-1] < 1  # test_obst_in_last_sensor
    assert 0 < closeness[0] < 1  # test_obst_in_first_sensor
    assert v.windows == [(-9, 5)]  # SURVEY probe: rays 171..179, 0..4
    assert not v.collision


@pytest.mark.parametrize("heading_deg,seen", [(45, True), (90, True), (170, True), (-135, False), (-170, False), (0, False)])
def test_seam_bug_obstacle_dead_astern(heading_deg, seen):
    """an obstacle dead astern is tested by 14 rays for some headings and
    by none for others (range(idx_min-1, idx_max % n) is empty when the window crosses the
    seam with a positive unwrapped bearing)."""
    psi = np.deg2rad(heading_deg)
    pos = np.array([5.0, -5.0])
    behind = pos - 6.5 * np.array([math.cos(psi), math.sin(psi)])
    ob = S.OracleCircle(behind, 1.5)
    per_ray, windows = S.rays_for_obstacles([ob], pos, psi, 2 * math.pi / 180, 180)
    n = sum(len(r) for r in per_ray)
    assert (n in (14, 15)) if seen else (n == 0), windows


def test_culled_casting_equals_brute_force_away_from_seam():
    rng = np.random.RandomState(3)
    cfg = dict(S.DEFAULT_CFG)
    for trial in range(6):
        obstacles = [S.OracleCircle(rng.uniform(-120, 120, size=2), rng.uniform(2, 40)) for _ in range(5)]
        v = S.OracleVessel(cfg, [0.0, 0.0, rng.uniform(-3, 3)])
        if any(np.hypot(*o.position) < o.radius + 2 for o in obstacles):
            continue
        v.perceive(obstacles)
        angles = v.sensor_angles + v.heading
        brute = np.array([S.cast_ray(a, (0.0, 0.0), 150.0, v.nearby) for a in angles])
        per_ray, _ = S.rays_for_obstacles(v.nearby, (0.0, 0.0), v.heading, v.d_angle, 180)
        seen = {id(o) for r in per_ray for o in r}
        unseen = [o for o in v.nearby if id(o) not in seen]  # seam victims
        if not unseen:
            assert np.allclose(v.dists, brute, atol=1e-9)
        else:
            assert np.all(v.dists >= brute - 1e-9)


def test_filled_polygon_inside_gives_zero_range_ring_does_not():
    vo = S.OracleVesselObstacle(30.0, [0.0, 0.0], np.broadcast_to((0.0, 1.0), (100, 2)), init_update=False)
    closeness, v = _perceive([0.0, 1.0, 0.0], [vo])
    assert v.collision and np.nanmin(v.dists) == 0.0
    ring = S.OracleCircle([0.0, 0.0], 30.0)
    _, v2 = _perceive([0.0, 1.0, 0.0], [ring])
    assert not v2.collision and v2.dists.min() > 25.0  # inside a ring: rays measure the exit distance


def test_negative_radius_raises_like_reference():
    with pytest.raises(ValueError):
        S.OracleCircle([0, 0], -1.0)
