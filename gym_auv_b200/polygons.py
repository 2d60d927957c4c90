"""Static land polygons (``PolygonObstacle``, gym_auv/objects/obstacles.py:116-127) as a
world shared by every scenario of a pool: flat vertex arrays + the cached enclosing
circle of ``enclosing_circle_of_shape`` (obstacles.py:235-262: centre and half-diagonal
of the minimum rotated rectangle, which Shapely 1.7 computes in pure Python by trying
every convex-hull edge direction and keeping the first smallest-area box)."""
from __future__ import annotations

import math
from typing import List, Sequence, Tuple

import numpy as np


def _convex_hull(pts: np.ndarray) -> np.ndarray:
    p = sorted(set(map(tuple, np.asarray(pts, dtype=np.float64))))
    if len(p) <= 2:
        return np.array(p)

    def cross(o, a, b):
        return (a[0] - o[0]) * (b[1] - o[1]) - (a[1] - o[1]) * (b[0] - o[0])

    lo: List[Tuple[float, float]] = []
    for q in p:
        while len(lo) >= 2 and cross(lo[-2], lo[-1], q) <= 0:
            lo.pop()
        lo.append(q)
    up: List[Tuple[float, float]] = []
    for q in reversed(p):
        while len(up) >= 2 and cross(up[-2], up[-1], q) <= 0:
            up.pop()
        up.append(q)
    return np.array(lo[:-1] + up[:-1])


def enclosing_circle(ring: np.ndarray) -> Tuple[np.ndarray, float]:
    """(centre, radius) of the reference's enclosing circle for a polygon ring."""
    hull = _convex_hull(ring)
    if len(hull) < 3:
        raise ValueError("degenerate polygon")
    best_area, best_box = None, None
    nh = len(hull)
    for k in range(nh):
        d = hull[(k + 1) % nh] - hull[k]
        length = math.hypot(d[0], d[1])
        if length == 0.0:
            continue
        u = d / length
        v = np.array([-u[1], u[0]])
        tx, ty = hull @ u, hull @ v
        x0, x1, y0, y1 = tx.min(), tx.max(), ty.min(), ty.max()
        area = (x1 - x0) * (y1 - y0)
        if best_area is None or area < best_area:
            corners = np.array([[x0, y0], [x1, y0], [x1, y1], [x0, y1]])
            best_area, best_box = area, corners[:, :1] * u[None, :] + corners[:, 1:] * v[None, :]
    centre = best_box.mean(axis=0)  # centroid of a rectangle
    radius = float(np.max(np.hypot(best_box[:, 0] - centre[0], best_box[:, 1] - centre[1])))
    return centre, radius


class World:
    """P static polygons -> the flat arrays of AuvScenarioPool.world_*."""

    def __init__(self, polygons: Sequence[np.ndarray]):
        rings = []
        for poly in polygons:
            r = np.asarray(poly, dtype=np.float64)
            if r.ndim != 2 or r.shape[1] != 2 or len(r) < 3:
                raise ValueError("a polygon needs >= 3 (x, y) vertices")
            if not np.array_equal(r[0], r[-1]):
                r = np.vstack([r, r[:1]])
            # (any vertex count: the casting stage walks perimeters longer than its vertex stage chain by chain)
            rings.append(r)
        self.rings = rings
        self.n = len(rings)
        self.voff = np.zeros(self.n + 1, dtype=np.int32)
        for i, r in enumerate(rings):
            self.voff[i + 1] = self.voff[i] + len(r)
        self.verts = np.concatenate(rings, axis=0) if rings else np.zeros((0, 2))
        circ = np.zeros((self.n, 3))
        for i, r in enumerate(rings):
            c, rad = enclosing_circle(r)
            circ[i] = (c[0], c[1], rad)
        self.circle = circ
        self._grid = None

    def grid(self, cell: float = 160.0):
        """Uniform grid over the enclosing circles (AuvScenarioPool.world_cell_*): CSR lists of the
        polygons whose circle overlaps each cell.  Returns dict(x0, y0, cell, nx, ny, off, items)."""
        if self._grid is not None and self._grid["cell"] == cell:
            return self._grid
        if self.n == 0:
            return None
        c = self.circle
        x0, y0 = float((c[:, 0] - c[:, 2]).min()), float((c[:, 1] - c[:, 2]).min())
        x1, y1 = float((c[:, 0] + c[:, 2]).max()), float((c[:, 1] + c[:, 2]).max())
        nx, ny = int(max(1, math.ceil((x1 - x0) / cell))), int(max(1, math.ceil((y1 - y0) / cell)))
        cells = [[] for _ in range(nx * ny)]
        for j in range(self.n):
            ix0 = int(max(0, math.floor((c[j, 0] - c[j, 2] - x0) / cell)))
            ix1 = int(min(nx - 1, math.floor((c[j, 0] + c[j, 2] - x0) / cell)))
            iy0 = int(max(0, math.floor((c[j, 1] - c[j, 2] - y0) / cell)))
            iy1 = int(min(ny - 1, math.floor((c[j, 1] + c[j, 2] - y0) / cell)))
            for iy in range(iy0, iy1 + 1):
                for ix in range(ix0, ix1 + 1):
                    cells[iy * nx + ix].append(j)
        off = np.zeros(nx * ny + 1, dtype=np.int32)
        off[1:] = np.cumsum([len(v) for v in cells])
        items = np.array([j for v in cells for j in v], dtype=np.int32) if off[-1] else np.zeros(1, dtype=np.int32)
        self._grid = dict(x0=x0, y0=y0, cell=float(cell), nx=nx, ny=ny, off=off, items=items)
        return self._grid


def star_polygon(rng, centre, circumradius: float, n_vertices: int) -> np.ndarray:
    """Star-shaped (about its centre) simple polygon: sorted random angles, radii in
    [0.45, 1] x circumradius."""
    ang = np.sort(rng.uniform(0, 2 * np.pi, n_vertices))
    rad = circumradius * rng.uniform(0.45, 1.0, n_vertices)
    return np.asarray(centre)[None, :] + rad[:, None] * np.stack([np.cos(ang), np.sin(ang)], axis=1)


def random_land(n_polygons: int = 512, extent: float = 6000.0, seed: int = 0, keep_clear=None, clear_margin: float = 30.0):
    """Synthetic land (BASELINE config 4; the reference's terrain files are not shipped):
    star-shaped polygons, 8..64 vertices, circum-radius LogNormal(ln 60, 0.5) m, uniformly
    placed in an extent x extent box centred on the origin, rejected when they come within
    `clear_margin` of any point of `keep_clear` ([n,2] array, e.g. vessel start positions)."""
    rng = np.random.RandomState(seed)
    polys = []
    tries = 0
    while len(polys) < n_polygons and tries < 50 * n_polygons:
        tries += 1
        c = rng.uniform(-extent / 2, extent / 2, size=2)
        r = float(np.exp(rng.normal(np.log(60.0), 0.5)))
        if keep_clear is not None and len(keep_clear):
            if np.min(np.hypot(keep_clear[:, 0] - c[0], keep_clear[:, 1] - c[1])) < r + clear_margin:
                continue
        polys.append(star_polygon(rng, c, r, int(rng.randint(8, 65))))
    return polys
