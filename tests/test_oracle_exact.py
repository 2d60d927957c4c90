"""Independent check of the oracle's geometric primitives against EXACT rational arithmetic
(``fractions.Fraction``) on the same definitions.  The GEOS-backed half of the oracle cannot be
pinned to the reference (Shapely/GEOS are not installable here); what can
be pinned is that the float restatement computes what its definition says: closed-segment
distance, closed-interval ray/edge intersection, crossing-number containment, first-minimum
linear referencing.  Inputs are dyadic rationals, so the float and the rational side see
exactly the same numbers."""
import math
from fractions import Fraction as F

import numpy as np
import pytest

from oracle import geos_lite as G


def dyadic(rng, shape, scale=64, span=40):
    return rng.randint(-span * scale, span * scale + 1, size=shape) / float(scale)


def exact_pt_seg_d2(p, a, b):
    px, py, ax, ay, bx, by = map(F, (*p, *a, *b))
    ex, ey = bx - ax, by - ay
    l2 = ex * ex + ey * ey
    if l2 == 0:
        return (px - ax) ** 2 + (py - ay) ** 2
    r = ((px - ax) * ex + (py - ay) * ey) / l2
    r = min(max(r, F(0)), F(1))
    cx, cy = ax + r * ex, ay + r * ey
    return (px - cx) ** 2 + (py - cy) ** 2


def exact_ray_edge_t(p0, p1, a, b):
    """Smallest parameter t in [0,1] with P0 + t (P1-P0) on the closed segment AB, or None."""
    p0x, p0y, p1x, p1y, ax, ay, bx, by = map(F, (*p0, *p1, *a, *b))
    dx, dy, ex, ey = p1x - p0x, p1y - p0y, bx - ax, by - ay
    den = dx * ey - dy * ex
    if den == 0:
        if dx * (ay - p0y) - dy * (ax - p0x) != 0:
            return None  # parallel, not collinear
        l2 = dx * dx + dy * dy
        ta = ((ax - p0x) * dx + (ay - p0y) * dy) / l2
        tb = ((bx - p0x) * dx + (by - p0y) * dy) / l2
        lo, hi = min(ta, tb), max(ta, tb)
        if hi < 0 or lo > 1:
            return None
        return max(lo, F(0))
    t = ((ax - p0x) * ey - (ay - p0y) * ex) / den
    u = ((ax - p0x) * dy - (ay - p0y) * dx) / den
    if 0 <= t <= 1 and 0 <= u <= 1:
        return t
    return None


@pytest.mark.parametrize("seed", range(4))
def test_point_segment_distance_is_exact_to_rounding(seed):
    rng = np.random.RandomState(seed)
    for _ in range(300):
        p, a, b = dyadic(rng, 2), dyadic(rng, 2), dyadic(rng, 2)
        if rng.rand() < 0.05:
            b = a.copy()  # degenerate segment
        d = G.point_segment_distance(p[0], p[1], a[0], a[1], b[0], b[1])
        want = math.sqrt(float(exact_pt_seg_d2(p, a, b)))
        assert abs(d - want) <= 1e-12 * max(1.0, want)


@pytest.mark.parametrize("seed", range(4))
def test_ray_ring_min_distance_matches_exact_intersections(seed):
    rng = np.random.RandomState(100 + seed)
    for _ in range(120):
        n = rng.randint(3, 9)
        ring = dyadic(rng, (n, 2))
        ring = np.vstack([ring, ring[:1]])
        p0, p1 = dyadic(rng, 2), dyadic(rng, 2)
        if np.array_equal(p0, p1):
            continue
        if rng.rand() < 0.2:  # aim the ray exactly through a vertex (touching counts, closed intervals)
            k = rng.randint(n)
            p1 = p0 + 2 * (ring[k] - p0)
            if np.array_equal(p0, p1):
                continue
        ts = [exact_ray_edge_t(p0, p1, ring[k], ring[k + 1]) for k in range(n)]
        ts = [t for t in ts if t is not None]
        length = math.hypot(*(p1 - p0))
        for fn in (G.ray_ring_min_distance, G.ray_ring_min_distance_np):
            got = fn(p0, p1, ring)
            if not ts:
                assert got is None
            else:
                assert got is not None and abs(got - float(min(ts)) * length) <= 1e-9 * max(1.0, length)


@pytest.mark.parametrize("seed", range(3))
def test_point_in_ring_matches_exact_crossing_number(seed):
    rng = np.random.RandomState(200 + seed)
    for _ in range(200):
        n = rng.randint(3, 10)
        ring = dyadic(rng, (n, 2), scale=4, span=12)
        ring = np.vstack([ring, ring[:1]])
        p = dyadic(rng, 2, scale=4, span=12) + 1.0 / 16  # off the dyadic grid of the vertices: never on an edge's y level
        inside = False
        for k in range(n):
            x1, y1, x2, y2, x, y = map(F, (*ring[k], *ring[k + 1], *p))
            if (y1 > y) != (y2 > y):
                xint = x1 + (y - y1) * (x2 - x1) / (y2 - y1)
                if xint > x:
                    inside = not inside
        assert G.point_in_ring(p, ring) == inside


@pytest.mark.parametrize("seed", range(3))
def test_linestring_project_picks_the_exact_first_minimum(seed):
    rng = np.random.RandomState(300 + seed)
    for _ in range(60):
        m = rng.randint(2, 40)
        pts = np.cumsum(dyadic(rng, (m, 2), scale=8, span=2), axis=0)
        p = dyadic(rng, 2, scale=8, span=30)
        d2 = [exact_pt_seg_d2(p, pts[k], pts[k + 1]) for k in range(m - 1)]
        k = d2.index(min(d2))  # first exact minimum
        start = sum(math.hypot(*(pts[j + 1] - pts[j])) for j in range(k))
        a, b = pts[k], pts[k + 1]
        e = b - a
        l2 = float(e @ e)
        r = 0.0 if l2 == 0 else min(max(float((p - a) @ e) / l2, 0.0), 1.0)
        want = start + r * math.sqrt(l2)
        ties = sum(1 for v in d2 if v == d2[k])
        got = G.linestring_project(pts, p)
        got_seq = G.linestring_project_sequential(pts, p)
        assert got == got_seq
        if ties == 1 and (len(d2) == 1 or sorted(d2)[1] - d2[k] > F(1, 10**9)):  # a unique, well separated minimum is unambiguous in FP64
            assert abs(got - want) <= 1e-9 * max(1.0, want)


@pytest.mark.parametrize("seed", range(3))
def test_minimum_rotated_rectangle_is_the_geometric_minimum(seed):
    """Shapely's minimum_rotated_rectangle enumerates the hull edges; by the rotating-calipers theorem
    the minimum-area enclosing rectangle has a side on a hull edge, so the restatement must (a) contain
    every point and (b) not be beaten by any rectangle of a fine sweep over orientations."""
    rng = np.random.RandomState(400 + seed)
    for _ in range(40):
        pts = rng.uniform(-30, 30, size=(rng.randint(3, 12), 2))
        ring = np.vstack([pts, pts[:1]])
        box = G.minimum_rotated_rectangle(ring)
        e0, e1 = box[1] - box[0], box[3] - box[0]
        area = abs(e0[0] * e1[1] - e0[1] * e1[0])
        assert abs(e0 @ e1) <= 1e-9 * max(1.0, area)  # a rectangle
        u, v = e0 / np.linalg.norm(e0), e1 / np.linalg.norm(e1)
        rel = pts - box[0]
        assert (rel @ u).min() >= -1e-9 and (rel @ u).max() <= np.linalg.norm(e0) + 1e-9
        assert (rel @ v).min() >= -1e-9 and (rel @ v).max() <= np.linalg.norm(e1) + 1e-9
        th = np.linspace(0.0, np.pi / 2, 2001)
        c, s = np.cos(th)[:, None], np.sin(th)[:, None]
        x = pts[:, 0][None] * c + pts[:, 1][None] * s
        y = -pts[:, 0][None] * s + pts[:, 1][None] * c
        sweep = ((x.max(1) - x.min(1)) * (y.max(1) - y.min(1))).min()
        assert area <= sweep * (1 + 1e-9)
        # the enclosing circle the reference derives from it (obstacles.py:235-262) contains every point
        centre, radius = G.enclosing_circle_of_ring(ring)
        assert np.linalg.norm(pts - np.asarray(centre), axis=1).max() <= radius + 1e-9


def test_vessel_pentagon_centroid_and_enclosing_circle_closed_forms():
    """The closed forms the kernels use: area centroid (5w/18, 0) of the pentagon,
    enclosing circle = centre of its 2w x w rectangle, radius w sqrt(5)/2, for any heading."""
    for w in (1.0, 6.0, 30.0):
        body = np.array([(-w / 2, -w / 2), (-w / 2, w / 2), (w / 2, w / 2), (1.5 * w, 0.0), (w / 2, -w / 2), (-w / 2, -w / 2)])
        c = G.polygon_centroid(body)
        assert np.abs(c - np.array([5 * w / 18, 0.0])).max() <= 1e-12 * w
        for th in (0.0, 0.3, 2.0, -2.5):
            ring = G.rotate_about(body, th, c) + np.array([10.0, -4.0])
            centre, radius = G.enclosing_circle_of_ring(ring)
            want = c + np.array([np.cos(th), np.sin(th)]) * (w / 2 - 5 * w / 18) + np.array([10.0, -4.0])
            assert np.abs(np.asarray(centre) - want).max() <= 1e-9 * w and abs(radius - w * math.sqrt(5) / 2) <= 1e-9 * w
