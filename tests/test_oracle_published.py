"""The oracle's Shapely/GEOS restatement (oracle/geos_lite.py) pinned to PUBLISHED third-party
vectors.

The reference's LiDAR path computes inside Shapely 1.7.0 / GEOS (requirements.txt:5): it is not
vendored under /root/reference and not installable here (no wheel, no libgeos, no network), and
the reference's own tests hold no numeric vectors for these primitives.  What exists are the
worked examples of the Shapely User Manual (https://shapely.readthedocs.io/en/stable/manual.html,
sections cited per test; the 1.7 and 1.8 manuals print the same numbers) and the assertions of
Shapely's own test-suite.  They are restated here from the published text -- this container has no
network, so the numbers were typed in, not fetched -- and every one of them is an output of the
real library for the exact call the reference makes:

  Point.buffer(r)                      -> CircularObstacle._calculate_boundary   obstacles.py:101-106
  .simplify(tol, preserve_topology=False)                                         obstacles.py:104-106
  LineString.project / interpolate     -> Path.get_closest_arclength             path.py:84-93
  Point.distance(geometry)             -> nearby list, sensor ranges             vessel.py:269, sensor.py:152
  affinity.rotate(origin=...)          -> VesselObstacle._calculate_boundary     obstacles.py:217-228
  minimum_rotated_rectangle            -> enclosing_circle_of_shape              obstacles.py:235-262
  LineString.intersection              -> simulate_sensor                        sensor.py:140-159
"""
import math

import numpy as np

from oracle import geos_lite as G


def _ring_area(ring):
    r = np.asarray(ring)
    return 0.5 * abs(float(np.sum(r[:-1, 0] * r[1:, 1] - r[1:, 0] * r[:-1, 1])))


# ---------------------------------------------------------------------------------------
# Manual, "object.buffer(distance, resolution=16, ...)":
#   >>> p = Point(0, 0).buffer(10.0)
#   >>> len(p.exterior.coords)
#   66
#   >>> p.area
#   313.65484905459385
#   >>> q = Point(0, 0).buffer(10.0, 3)
#   >>> len(q.exterior.coords)
#   14
#   >>> q.area
#   300.0   (printed as 299.99999999999994 / 300.00000000000006 depending on the GEOS build)
#   and "a resolution of 1 gives a square patch": 5 coords, area 200.0
# (GEOS >= 3.8 no longer repeats the start vertex: 65 / 13 / 5 coordinates for the same rings.)
# ---------------------------------------------------------------------------------------
def test_buffer_manual_example_areas_and_vertex_counts():
    p = G.buffer_point_ring(0.0, 0.0, 10.0)
    assert len(p) in (65, 66)
    assert abs(_ring_area(p) - 313.65484905459385) <= 1e-11
    q = G.buffer_point_ring(0.0, 0.0, 10.0, quadsegs=3)
    assert len(q) in (13, 14)
    assert abs(_ring_area(q) - 300.0) <= 1e-10
    sq = G.buffer_point_ring(0.0, 0.0, 10.0, quadsegs=1)
    assert len(sq) == 5 and abs(_ring_area(sq) - 200.0) <= 1e-10
    # GEOS starts at angle 0 and sweeps CLOCKWISE: the WKT of a unit buffer begins
    #   POLYGON ((1 0, 0.9951847266721969 -0.0980171403295606, ...     (Shapely default, resolution 16)
    #   POLYGON ((1 0, 0.9807852804032304 -0.1950903220161282, ...     (PostGIS ST_Buffer default, quad_segs 8)
    u = G.buffer_point_ring(0.0, 0.0, 1.0)
    assert tuple(u[0]) == (1.0, 0.0)
    assert abs(u[1, 0] - 0.9951847266721969) <= 1e-15 and abs(u[1, 1] + 0.0980171403295606) <= 1e-15
    u8 = G.buffer_point_ring(0.0, 0.0, 1.0, quadsegs=8)
    assert abs(u8[1, 0] - 0.9807852804032304) <= 1e-15 and abs(u8[1, 1] + 0.1950903220161282) <= 1e-15


# ---------------------------------------------------------------------------------------
# Manual, "object.simplify(tolerance, preserve_topology=True)":
#   >>> p = Point(0.0, 0.0)
#   >>> x = p.buffer(1.0)
#   >>> x.area
#   3.1365484905459389
#   >>> len(x.exterior.coords)
#   66
#   >>> s = x.simplify(0.05, preserve_topology=False)
#   >>> s.area
#   3.0614674589207187
#   >>> len(s.exterior.coords)
#   17
# ---------------------------------------------------------------------------------------
def test_simplify_manual_example():
    x = G.buffer_point_ring(0.0, 0.0, 1.0)
    assert abs(_ring_area(x) - 3.1365484905459389) <= 1e-15
    s = G.douglas_peucker(x, 0.05)
    assert len(s) == 17
    assert abs(_ring_area(s) - 3.0614674589207187) <= 1e-15
    # Douglas-Peucker is scale-invariant: buffer(6).simplify(0.3) is the same 16-gon scaled by 6 --
    # which is what CircularObstacle._calculate_boundary builds for a radius-6 obstacle and what the
    # closed form used by the kernels (ngon_sides) must agree with
    ring6 = G.circle_boundary_ring(0.0, 0.0, 6.0)
    assert len(ring6) == 17 and G.circle_ngon_sides(6.0) == 16
    assert abs(_ring_area(ring6) - 36.0 * 3.0614674589207187) <= 1e-12
    ang = np.sort(np.mod(np.arctan2(ring6[:-1, 1], ring6[:-1, 0]), 2 * np.pi))
    assert np.abs(ang - np.arange(16) * (2 * np.pi / 16)).max() <= 1e-12  # vertices at k 2 pi / 16 from +x


# ---------------------------------------------------------------------------------------
# Manual, "Linear Referencing Methods":
#   >>> ip = LineString([(0, 0), (0, 1), (1, 1)]).interpolate(1.5)
#   >>> ip.wkt
#   'POINT (0.5000000000000000 1.0000000000000000)'
#   >>> LineString([(0, 0), (0, 1), (1, 1)]).project(ip)
#   1.5
#   >>> LineString([(0, 0), (0, 1), (1, 1)]).project(ip, normalized=True)
#   0.75
# ---------------------------------------------------------------------------------------
def test_project_manual_example():
    line = np.array([(0.0, 0.0), (0.0, 1.0), (1.0, 1.0)])
    assert G.linestring_project(line, (0.5, 1.0)) == 1.5
    assert G.linestring_project_sequential(line, (0.5, 1.0)) == 1.5
    assert G.linestring_project(line, (0.5, 1.0)) / 2.0 == 0.75
    # the same example off the line: the nearest point of (0.5, 3) is still (0.5, 1)
    assert G.linestring_project(line, (0.5, 3.0)) == 1.5
    # ties go to the FIRST segment (LengthIndexOfPoint keeps a strictly smaller distance only):
    # (1, 0) is at distance 1 from both legs' nearest points (0, 0)..(0, 1) and (1, 1)
    assert G.linestring_project(line, (1.0, 0.0)) == 0.0


# ---------------------------------------------------------------------------------------
# Manual, "object.distance(other)":      >>> Point(0,0).distance(Point(1,1))   1.4142135623730951
# Manual, "object.hausdorff_distance":   >>> point = Point(1, 1); line = LineString([(2, 0), (2, 4), (3, 4)])
#                                        >>> point.distance(line)               1.0
# Manual, "shapely.ops.nearest_points":  triangle = Polygon([(0, 0), (1, 0), (0.5, 1), (0, 0)]),
#                                        square = Polygon([(0, 2), (1, 2), (1, 3), (0, 3), (0, 2)])
#                                        -> ['POINT (0.5 1)', 'POINT (0.5 2)']   (distance 1.0)
# Manual, "Polygons":                    >>> Polygon([(0, 0), (1, 1), (1, 0)]).area  0.5;  .length  3.4142135623730949
# ---------------------------------------------------------------------------------------
def test_distance_manual_examples():
    assert G.point_segment_distance(0.0, 0.0, 1.0, 1.0, 1.0, 1.0) == 1.4142135623730951
    line = np.array([(2.0, 0.0), (2.0, 4.0), (3.0, 4.0)])
    assert min(G.point_segment_distance(1.0, 1.0, *line[k], *line[k + 1]) for k in range(2)) == 1.0
    square = np.array([(0.0, 2.0), (1.0, 2.0), (1.0, 3.0), (0.0, 3.0), (0.0, 2.0)])
    assert G.point_ring_distance((0.5, 1.0), square) == 1.0
    assert G.point_polygon_distance((0.5, 1.0), square) == 1.0
    assert G.point_polygon_distance((0.5, 2.5), square) == 0.0  # inside a filled polygon
    assert G.point_ring_distance((0.5, 2.5), square) == 0.5     # its boundary is a ring
    tri = np.array([(0.0, 0.0), (1.0, 1.0), (1.0, 0.0), (0.0, 0.0)])
    assert _ring_area(tri) == 0.5
    per = sum(math.hypot(*(tri[k + 1] - tri[k])) for k in range(3))
    assert abs(per - 3.4142135623730949) <= 1e-15
    assert np.abs(G.polygon_centroid(tri) - np.array([2.0 / 3.0, 1.0 / 3.0])).max() <= 1e-15


# ---------------------------------------------------------------------------------------
# Manual, "Affine Transformations", shapely.affinity.rotate(geom, angle, origin='center'):
#   "The affine transformation matrix for 2D rotation is:
#        / cos(r) -sin(r) xoff \        xoff = x0 - x0 cos(r) + y0 sin(r)
#        | sin(r)  cos(r) yoff |        yoff = y0 - x0 sin(r) - y0 cos(r)
#        \   0       0      1  /        where (x0, y0) is the origin"
#   worked example: line = LineString([(1, 3), (1, 1), (4, 1)]); rotate(line, 90, origin='centroid')
#   (the figure's centroid of that line is the length-weighted mean (1.9, 1.4)).
# ---------------------------------------------------------------------------------------
def test_rotate_follows_the_published_matrix():
    line = np.array([(1.0, 3.0), (1.0, 1.0), (4.0, 1.0)])
    c = (1.9, 1.4)
    out = G.rotate_about(line, math.pi / 2, c)
    # 90 degrees about (x0, y0): (x, y) -> (x0 - (y - y0), y0 + (x - x0)), exactly (cos snapped to 0)
    want = np.array([(c[0] - (y - c[1]), c[1] + (x - c[0])) for x, y in line])
    assert np.abs(out - want).max() <= 1e-15
    r = 0.7
    x0, y0 = -2.0, 5.0
    m = np.array([[math.cos(r), -math.sin(r), x0 - x0 * math.cos(r) + y0 * math.sin(r)],
                  [math.sin(r), math.cos(r), y0 - x0 * math.sin(r) - y0 * math.cos(r)]])
    got = G.rotate_about(line, r, (x0, y0))
    want = (m[:, :2] @ line.T).T + m[:, 2]
    assert np.abs(got - want).max() <= 1e-15
    # the vessel pentagon of obstacles.py:175-181 rotates about its AREA centroid (5w/18, 0)
    w = 9.0
    body = np.array([(-w / 2, -w / 2), (-w / 2, w / 2), (w / 2, w / 2), (3 * w / 2, 0.0), (w / 2, -w / 2), (-w / 2, -w / 2)])
    assert np.abs(G.polygon_centroid(body) - np.array([5 * w / 18, 0.0])).max() <= 1e-14


# ---------------------------------------------------------------------------------------
# Shapely test-suite, tests/test_minimum_rotated_rectangle.py (1.7):
#   poly = Polygon([(0,1), (1, 2), (2, 1), (1, 0), (0, 1)]); rect = poly.minimum_rotated_rectangle
#   assert rect.area - poly.area < 0.1;  assert len(rect.exterior.coords) == 5
#   ls = LineString([(0,1), (1, 2), (2, 1), (1, 0)]); rect = ls.minimum_rotated_rectangle
#   assert rect.area - ls.convex_hull.area < 0.1;  assert len(rect.exterior.coords) == 5
# Manual, "object.minimum_rotated_rectangle": "Returns the general minimum bounding rectangle that
#   contains the object. Unlike envelope this rectangle is not constrained to be parallel to the
#   coordinate axes."
# ---------------------------------------------------------------------------------------
def test_minimum_rotated_rectangle_published_assertions():
    diamond = np.array([(0.0, 1.0), (1.0, 2.0), (2.0, 1.0), (1.0, 0.0), (0.0, 1.0)])
    rect = G.minimum_rotated_rectangle(diamond)
    assert len(rect) == 5
    assert abs(_ring_area(rect) - _ring_area(diamond)) < 1e-12  # the diamond is its own rectangle (area 2)
    assert abs(_ring_area(rect) - 2.0) < 1e-12
    # axis-parallel envelope of the same diamond has area 4: the rotated rectangle must beat it
    centre, radius = G.enclosing_circle_of_ring(diamond)
    assert np.abs(centre - np.array([1.0, 1.0])).max() <= 1e-12 and abs(radius - 1.0) <= 1e-12
    # an axis-parallel rectangle is its own minimum rotated rectangle
    box = np.array([(0.0, 0.0), (4.0, 0.0), (4.0, 2.0), (0.0, 2.0), (0.0, 0.0)])
    assert abs(_ring_area(G.minimum_rotated_rectangle(box)) - 8.0) < 1e-12
    c, r = G.enclosing_circle_of_ring(box)
    assert np.abs(c - np.array([2.0, 1.0])).max() <= 1e-12 and abs(r - math.sqrt(5.0)) <= 1e-12


# ---------------------------------------------------------------------------------------
# Manual, "Binary Predicates" / "object.intersection(other)":
#   >>> LineString([(0, 0), (1, 1)]).crosses(LineString([(0, 1), (1, 0)]))      True
#   >>> LineString([(0, 0), (1, 1)]).intersection(LineString([(0, 1), (1, 0)])).wkt
#   'POINT (0.5 0.5)'
#   >>> LineString([(0, 0), (1, 1)]).touches(LineString([(1, 1), (2, 2)])) ... closed-interval semantics:
#   a segment END on the other line counts as an intersection.
# ---------------------------------------------------------------------------------------
def test_ray_intersection_published_examples():
    ring = np.array([(0.0, 1.0), (1.0, 0.0), (0.0, 1.0)])  # the second line as a degenerate closed chain
    d = G.ray_ring_min_distance((0.0, 0.0), (1.0, 1.0), ring)
    assert abs(d - math.hypot(0.5, 0.5)) <= 1e-15
    # touching at an end point counts (closed segments): ray (0,0)->(1,1) against (1,1)-(2,0)-(1,1)
    touch = np.array([(1.0, 1.0), (2.0, 0.0), (1.0, 1.0)])
    assert abs(G.ray_ring_min_distance((0.0, 0.0), (1.0, 1.0), touch) - math.sqrt(2.0)) <= 1e-15
    # the vectorised variant the oracle's sensor loop uses agrees
    assert abs(G.ray_ring_min_distance_np((0.0, 0.0), (1.0, 1.0), ring) - math.hypot(0.5, 0.5)) <= 1e-15
    # a ray against the manual's simplified unit buffer (regular 16-gon): straight at vertex 8 (-1, 0)
    s = G.douglas_peucker(G.buffer_point_ring(0.0, 0.0, 1.0), 0.05)
    assert abs(G.ray_ring_min_distance((-3.0, 0.0), (0.0, 0.0), s) - 2.0) <= 1e-15
    # ... and at the middle of an edge: apothem cos(pi/16)
    a = math.pi / 16
    d = G.ray_ring_min_distance((3.0 * math.cos(a), 3.0 * math.sin(a)), (0.0, 0.0), s)
    assert abs(d - (3.0 - math.cos(a))) <= 1e-14
