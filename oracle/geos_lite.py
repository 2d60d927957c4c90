"""FP64 restatement of the Shapely 1.7.0 / GEOS operations the reference calls
on its hot path.  TEST INFRASTRUCTURE (see ``oracle/__init__.py``).

Third-party dependency being restated: Shapely 1.7.0 (``requirements.txt:5``), which
bundles GEOS 3.8.  Neither is present in ``/root/reference`` nor installable here,
so each function below restates the *published* GEOS/JTS algorithm and names the
reference call site it stands in for.  PARITY UNPINNED for this file.

Conventions: points are ``(x, y)`` float64; rings are ``(n+1, 2)`` arrays whose last
vertex repeats the first.
"""
from __future__ import annotations

import math

import numpy as np

# --------------------------------------------------------------------------------------
# Distance primitives  (GEOS algorithm::Distance::pointToSegment)
# stands in for Point.distance(...)  -- reference sensor.py:26,152; vessel.py:269
# --------------------------------------------------------------------------------------


def point_segment_distance(px, py, ax, ay, bx, by):
    """GEOS ``Distance::pointToSegment``: distance from P to the closed segment AB."""
    if ax == bx and ay == by:
        return math.hypot(px - ax, py - ay)
    ex = bx - ax
    ey = by - ay
    len2 = ex * ex + ey * ey
    r = ((px - ax) * ex + (py - ay) * ey) / len2
    if r <= 0.0:
        return math.hypot(px - ax, py - ay)
    if r >= 1.0:
        return math.hypot(px - bx, py - by)
    s = ((ay - py) * ex - (ax - px) * ey) / len2
    return abs(s) * math.sqrt(len2)


def point_ring_distance(p, ring):
    """``Point.distance(LineString ring)``: min over the ring's segments."""
    ring = np.asarray(ring, dtype=np.float64)
    best = math.inf
    for k in range(len(ring) - 1):
        d = point_segment_distance(p[0], p[1], ring[k, 0], ring[k, 1], ring[k + 1, 0], ring[k + 1, 1])
        if d < best:
            best = d
    return best


def point_in_ring(p, ring):
    """Crossing-number point-in-polygon (GEOS RayCrossingCounter semantics for a
    point strictly inside / outside; boundary points count as inside)."""
    ring = np.asarray(ring, dtype=np.float64)
    x, y = float(p[0]), float(p[1])
    inside = False
    n = len(ring) - 1
    for k in range(n):
        x1, y1 = ring[k]
        x2, y2 = ring[k + 1]
        if (y1 > y) != (y2 > y):
            xint = x1 + (y - y1) * (x2 - x1) / (y2 - y1)
            if xint == x:
                return True
            if xint > x:
                inside = not inside
    return inside


def point_polygon_distance(p, ring):
    """``Point.distance(Polygon)``: 0 inside the filled polygon, else ring distance."""
    if point_in_ring(p, ring):
        return 0.0
    return point_ring_distance(p, ring)


# --------------------------------------------------------------------------------------
# Linear referencing (GEOS linearref::LengthIndexOfPoint::indexOfFromStart)
# stands in for LineString.project  -- reference path.py:93
# --------------------------------------------------------------------------------------


def linestring_project(points, p):
    """Arc-length (chord-sum measure) of the point of polyline ``points`` nearest to
    ``p``.  Segments are visited in order; a later segment replaces the current best
    only on a strictly smaller distance, so the FIRST minimum wins.  Vectorised but
    arithmetically identical to the sequential loop (np.argmin returns the first
    minimum)."""
    pts = np.asarray(points, dtype=np.float64)
    a = pts[:-1]
    b = pts[1:]
    e = b - a
    len2 = e[:, 0] ** 2 + e[:, 1] ** 2
    seglen = np.sqrt(len2)
    start = np.concatenate([[0.0], np.cumsum(seglen)[:-1]])
    wx = p[0] - a[:, 0]
    wy = p[1] - a[:, 1]
    with np.errstate(divide="ignore", invalid="ignore"):
        r = (wx * e[:, 0] + wy * e[:, 1]) / len2
        s = ((-wy) * e[:, 0] - (-wx) * e[:, 1]) / len2
    d_a = np.hypot(wx, wy)
    d_b = np.hypot(p[0] - b[:, 0], p[1] - b[:, 1])
    d_perp = np.abs(s) * seglen
    dist = np.where(r <= 0.0, d_a, np.where(r >= 1.0, d_b, d_perp))
    degenerate = len2 == 0.0
    dist = np.where(degenerate, d_a, dist)
    i = int(np.argmin(dist))
    # segmentNearestMeasure
    if degenerate[i] or r[i] <= 0.0:
        return float(start[i])
    if r[i] <= 1.0:
        return float(start[i] + r[i] * seglen[i])
    return float(start[i] + seglen[i])


def linestring_project_sequential(points, p):
    """Literal sequential form of the loop above (used by tests to pin the
    vectorised form, small inputs only)."""
    pts = np.asarray(points, dtype=np.float64)
    best = math.inf
    measure = -1.0
    start = 0.0
    for k in range(len(pts) - 1):
        ax, ay = pts[k]
        bx, by = pts[k + 1]
        d = point_segment_distance(p[0], p[1], ax, ay, bx, by)
        seglen = math.hypot(bx - ax, by - ay)
        if ax == bx and ay == by:
            m = start
        else:
            r = ((p[0] - ax) * (bx - ax) + (p[1] - ay) * (by - ay)) / ((bx - ax) ** 2 + (by - ay) ** 2)
            if r <= 0.0:
                m = start
            elif r <= 1.0:
                m = start + r * seglen
            else:
                m = start + seglen
        if d < best and m > -1.0:
            best = d
            measure = m
        start += seglen
    return measure


# --------------------------------------------------------------------------------------
# buffer(r).boundary.simplify(0.3)   -- reference obstacles.py:101-106
# --------------------------------------------------------------------------------------

QUADSEGS = 16  # Shapely default resolution for Point.buffer


def buffer_point_ring(cx, cy, radius, quadsegs=QUADSEGS):
    """GEOS ``OffsetSegmentGenerator::createCircle``: start at angle 0 and sweep
    CLOCKWISE in steps of pi/(2*quadsegs); closed ring (last == first)."""
    quantum = (math.pi / 2.0) / quadsegs
    total = 2.0 * math.pi
    nsegs = int(total / quantum + 0.5)
    inc = total / nsegs
    pts = [(cx + radius, cy)]
    k = 0
    # angle accumulates k*inc; the first iteration re-emits angle 0 which the
    # segment list de-duplicates, so start from k = 1
    k = 1
    while k < nsegs:
        ang = -(k * inc)
        pts.append((cx + radius * math.cos(ang), cy + radius * math.sin(ang)))
        k += 1
    pts.append(pts[0])
    return np.array(pts, dtype=np.float64)


def douglas_peucker(points, tol):
    """GEOS ``DouglasPeuckerLineSimplifier`` (first farthest point wins, keep iff
    farthest distance > tol)."""
    pts = np.asarray(points, dtype=np.float64)
    n = len(pts)
    keep = np.ones(n, dtype=bool)
    stack = [(0, n - 1)]
    while stack:
        i, j = stack.pop()
        if i + 1 >= j:
            continue
        maxd = -1.0
        maxk = i
        for k in range(i + 1, j):
            d = point_segment_distance(pts[k, 0], pts[k, 1], pts[i, 0], pts[i, 1], pts[j, 0], pts[j, 1])
            if d > maxd:
                maxd = d
                maxk = k
        if maxd <= tol:
            keep[i + 1 : j] = False
        else:
            stack.append((i, maxk))
            stack.append((maxk, j))
    return pts[keep]


def circle_boundary_ring(cx, cy, radius, tol=0.3):
    """The ring the reference's ``CircularObstacle._calculate_boundary`` produces:
    64-gon buffer, Douglas-Peucker with tolerance 0.3 (a regular n-gon,
    n in {4,8,16,32,64}; see ``circle_ngon_sides``)."""
    return douglas_peucker(buffer_point_ring(cx, cy, radius), tol)


def circle_ngon_sides(radius, tol=0.3):
    """Closed form for the number of sides kept by ``circle_boundary_ring``:
    n = 64/m, m = largest power of two <= 32 with r*(1-cos(m*pi/64)) <= tol."""
    m = 32
    while m >= 1:
        if radius * (1.0 - math.cos(m * math.pi / 64.0)) <= tol:
            return 64 // m
        m //= 2
    return 64


# --------------------------------------------------------------------------------------
# affinity.rotate(origin="centroid") + translate, centroid
# reference obstacles.py:217-228
# --------------------------------------------------------------------------------------


def polygon_centroid(ring):
    """Area centroid of a simple polygon ring (GEOS Centroid, shoelace form)."""
    r = np.asarray(ring, dtype=np.float64)
    x0, y0 = r[:-1, 0], r[:-1, 1]
    x1, y1 = r[1:, 0], r[1:, 1]
    cr = x0 * y1 - x1 * y0
    a = cr.sum() / 2.0
    cx = ((x0 + x1) * cr).sum() / (6.0 * a)
    cy = ((y0 + y1) * cr).sum() / (6.0 * a)
    return np.array([cx, cy])


def rotate_about(ring, angle, origin):
    """``shapely.affinity.rotate(geom, angle, use_radians=True, origin=origin)``;
    Shapely snaps |cos|,|sin| < 2.5e-16 to 0."""
    cosp = math.cos(angle)
    sinp = math.sin(angle)
    if abs(cosp) < 2.5e-16:
        cosp = 0.0
    if abs(sinp) < 2.5e-16:
        sinp = 0.0
    x0, y0 = origin
    r = np.asarray(ring, dtype=np.float64)
    xoff = x0 - x0 * cosp + y0 * sinp
    yoff = y0 - x0 * sinp - y0 * cosp
    out = np.empty_like(r)
    out[:, 0] = cosp * r[:, 0] - sinp * r[:, 1] + xoff
    out[:, 1] = sinp * r[:, 0] + cosp * r[:, 1] + yoff
    return out


# --------------------------------------------------------------------------------------
# minimum_rotated_rectangle (pure Python in Shapely 1.7.0) -> enclosing circle
# reference obstacles.py:235-262
# --------------------------------------------------------------------------------------


def convex_hull_ring(points):
    """Andrew monotone chain; returns a closed CCW ring."""
    pts = sorted(set((float(x), float(y)) for x, y in np.asarray(points)[:, :2]))
    if len(pts) <= 2:
        return np.array(pts + pts[:1], dtype=np.float64)

    def cross(o, a, b):
        return (a[0] - o[0]) * (b[1] - o[1]) - (a[1] - o[1]) * (b[0] - o[0])

    lower = []
    for p in pts:
        while len(lower) >= 2 and cross(lower[-2], lower[-1], p) <= 0:
            lower.pop()
        lower.append(p)
    upper = []
    for p in reversed(pts):
        while len(upper) >= 2 and cross(upper[-2], upper[-1], p) <= 0:
            upper.pop()
        upper.append(p)
    hull = lower[:-1] + upper[:-1]
    return np.array(hull + hull[:1], dtype=np.float64)


def minimum_rotated_rectangle(ring):
    """Shapely 1.7.0 ``BaseGeometry.minimum_rotated_rectangle``: for every edge of
    the convex hull, the axis-parallel bounding box in the edge frame; smallest area
    (first minimum) wins.  Returns the 4 corners (closed ring, 5 rows)."""
    hull = convex_hull_ring(np.asarray(ring)[:-1])
    best = None
    for k in range(len(hull) - 1):
        dx = hull[k + 1, 0] - hull[k, 0]
        dy = hull[k + 1, 1] - hull[k, 1]
        length = math.sqrt(dx * dx + dy * dy)
        if length == 0.0:
            continue
        ux, uy = dx / length, dy / length
        vx, vy = -uy, ux
        tx = hull[:, 0] * ux + hull[:, 1] * uy
        ty = hull[:, 0] * vx + hull[:, 1] * vy
        x0, x1, y0, y1 = tx.min(), tx.max(), ty.min(), ty.max()
        area = (x1 - x0) * (y1 - y0)
        if best is None or area < best[0]:
            box = np.array([[x0, y0], [x1, y0], [x1, y1], [x0, y1], [x0, y0]])
            wx = box[:, 0] * ux + box[:, 1] * vx
            wy = box[:, 0] * uy + box[:, 1] * vy
            best = (area, np.stack([wx, wy], axis=1))
    return best[1]


def enclosing_circle_of_ring(ring):
    """``enclosing_circle_of_shape``: centre = centroid of the MRR, radius = max
    distance from it to an MRR corner (obstacles.py:235-262)."""
    mrr = minimum_rotated_rectangle(ring)
    centre = polygon_centroid(mrr)
    radius = max(math.hypot(cx - centre[0], cy - centre[1]) for cx, cy in mrr)
    return centre, radius


# --------------------------------------------------------------------------------------
# LineString(ray).intersection(boundary) + _standardize_intersect + min distance
# reference sensor.py:11-19,140-159
# --------------------------------------------------------------------------------------


def _segment_hits(p0x, p0y, p1x, p1y, ax, ay, bx, by):
    """Intersection points (closed intervals) of segments P0P1 and AB, returned as
    distances from P0 along P0P1 (metres).  Collinear overlap returns the overlap's
    end nearest to P0."""
    dx, dy = p1x - p0x, p1y - p0y
    ex, ey = bx - ax, by - ay
    d1 = dx * (ay - p0y) - dy * (ax - p0x)  # orient(P0,P1,A)
    d2 = dx * (by - p0y) - dy * (bx - p0x)  # orient(P0,P1,B)
    d3 = ex * (p0y - ay) - ey * (p0x - ax)  # orient(A,B,P0)
    d4 = ex * (p1y - ay) - ey * (p1x - ax)  # orient(A,B,P1)
    if d1 == 0.0 and d2 == 0.0:
        # collinear: project A, B on the ray parameter
        l2 = dx * dx + dy * dy
        ta = ((ax - p0x) * dx + (ay - p0y) * dy) / l2
        tb = ((bx - p0x) * dx + (by - p0y) * dy) / l2
        lo, hi = min(ta, tb), max(ta, tb)
        if hi < 0.0 or lo > 1.0:
            return []
        return [max(lo, 0.0) * math.sqrt(l2)]
    if (d1 > 0.0 and d2 > 0.0) or (d1 < 0.0 and d2 < 0.0):
        return []
    if (d3 > 0.0 and d4 > 0.0) or (d3 < 0.0 and d4 < 0.0):
        return []
    u = d1 / (d1 - d2)
    ix = ax + u * ex
    iy = ay + u * ey
    return [math.hypot(ix - p0x, iy - p0y)]


def ray_ring_min_distance(p0, p1, ring):
    """min distance from P0 to ``LineString([P0,P1]).intersection(ring)`` or None."""
    ring = np.asarray(ring, dtype=np.float64)
    best = None
    for k in range(len(ring) - 1):
        for t in _segment_hits(p0[0], p0[1], p1[0], p1[1], ring[k, 0], ring[k, 1], ring[k + 1, 0], ring[k + 1, 1]):
            if best is None or t < best:
                best = t
    return best


def ray_polygon_min_distance(p0, p1, ring):
    """Same for a FILLED polygon: the clipped ray starts at P0 if P0 is inside
    (distance 0), else at the first boundary crossing."""
    if point_in_ring(p0, ring):
        return 0.0
    return ray_ring_min_distance(p0, p1, ring)


def ray_ring_min_distance_np(p0, p1, ring):
    """NumPy-over-edges form of ``ray_ring_min_distance`` (same arithmetic per edge;
    tests pin it against the scalar loop).  Used for speed by the step oracle."""
    ring = np.asarray(ring, dtype=np.float64)
    ax, ay = ring[:-1, 0], ring[:-1, 1]
    bx, by = ring[1:, 0], ring[1:, 1]
    p0x, p0y = float(p0[0]), float(p0[1])
    dx, dy = p1[0] - p0x, p1[1] - p0y
    ex, ey = bx - ax, by - ay
    d1 = dx * (ay - p0y) - dy * (ax - p0x)
    d2 = dx * (by - p0y) - dy * (bx - p0x)
    d3 = ex * (p0y - ay) - ey * (p0x - ax)
    d4 = ex * (p1[1] - ay) - ey * (p1[0] - ax)
    coll = (d1 == 0.0) & (d2 == 0.0)
    hit = ~(((d1 > 0) & (d2 > 0)) | ((d1 < 0) & (d2 < 0))) & ~(((d3 > 0) & (d4 > 0)) | ((d3 < 0) & (d4 < 0))) & ~coll
    best = None
    if hit.any():
        with np.errstate(divide="ignore", invalid="ignore"):
            u = d1[hit] / (d1[hit] - d2[hit])
        ix = ax[hit] + u * ex[hit]
        iy = ay[hit] + u * ey[hit]
        best = float(np.min(np.hypot(ix - p0x, iy - p0y)))
    if coll.any():
        for k in np.nonzero(coll)[0]:
            for t in _segment_hits(p0x, p0y, p1[0], p1[1], ax[k], ay[k], bx[k], by[k]):
                if best is None or t < best:
                    best = t
    return best
