"""CPU restatement of the device path builder's arithmetic (csrc/auv_pathbuild.cuh).  TEST
INFRASTRUCTURE.

``gym_auv.objects.path.Path.__init__`` (path.py:19-40) delegates to SciPy (unpinned in the
reference's requirements): ``scipy.interpolate.pchip`` = PchipInterpolator -> CubicHermiteSpline
-> PPoly.  This module restates the published SciPy algorithm operation by operation --
``PchipInterpolator._find_derivatives`` / ``_edge_case`` (Fritsch-Carlson weighted harmonic
mean), ``CubicHermiteSpline.__init__``'s coefficients, ``_ppoly.evaluate_poly1``'s power sum,
``numpy.linspace`` -- in plain NumPy, and is pinned against SciPy itself in
tests/test_pathbuild_oracle.py (SciPy IS installed here).  The CUDA kernel follows the same
order with explicit round-to-nearest intrinsics and is compared with SciPy's output directly on
the GPU (tests/test_gpu_v2.py)."""
from __future__ import annotations

import numpy as np

N_KNOTS = 1000


def _edge_case(h0, h1, m0, m1):
    d = ((2 * h0 + h1) * m0 - h0 * m1) / (h0 + h1)
    if np.sign(d) != np.sign(m0):
        return 0.0
    if np.sign(m0) != np.sign(m1) and abs(d) > 3.0 * abs(m0):
        return 3.0 * m0
    return d


def pchip_derivatives(x, y):
    n = len(x)
    hk = x[1:] - x[:-1]
    mk = (y[1:] - y[:-1]) / hk
    if n == 2:
        return np.array([mk[0], mk[0]])
    d = np.zeros(n)
    for k in range(1, n - 1):
        mm, m = mk[k - 1], mk[k]
        if np.sign(m) != np.sign(mm) or m == 0.0 or mm == 0.0:
            d[k] = 0.0
        else:
            w1 = 2 * hk[k] + hk[k - 1]
            w2 = hk[k] + 2 * hk[k - 1]
            d[k] = 1.0 / ((w1 / mm + w2 / m) / (w1 + w2))
    d[0] = _edge_case(hk[0], hk[1], mk[0], mk[1])
    d[-1] = _edge_case(hk[-1], hk[-2], mk[-1], mk[-2])
    return d


def hermite_coefficients(x, y, d):
    dxr = x[1:] - x[:-1]
    slope = (y[1:] - y[:-1]) / dxr
    t = (d[:-1] + d[1:] - 2 * slope) / dxr
    return np.stack([t / dxr, (slope - d[:-1]) / dxr - t, d[:-1], y[:-1]], axis=1)  # [n-1, 4]: c0..c3


def linspace(stop, num):
    step = stop / (num - 1)
    out = np.arange(num) * step
    out[-1] = stop
    return out


def evaluate(x, c, xq):
    j = np.clip(np.searchsorted(x, xq, side="right") - 1, 0, len(x) - 2)
    s = xq - x[j]
    res = c[j, 3] * 1.0
    z = s
    res = res + c[j, 2] * z
    z = z * s
    res = res + c[j, 1] * z
    z = z * s
    res = res + c[j, 0] * z
    return res


def build(waypoints):
    """-> dict(knots[1000], cx[999,4], cy[999,4], length, poly[n,2], cum[n])."""
    pts = np.array(waypoints, dtype=np.float64)
    for rnd in range(3):
        seg = np.sqrt((pts[0, 1:] - pts[0, :-1]) ** 2 + (pts[1, 1:] - pts[1, :-1]) ** 2)
        arc = np.concatenate([[0.0], np.cumsum(seg)])
        cx = hermite_coefficients(arc, pts[0], pchip_derivatives(arc, pts[0]))
        cy = hermite_coefficients(arc, pts[1], pchip_derivatives(arc, pts[1]))
        if rnd == 2:
            break
        xq = linspace(arc[-1], N_KNOTS)
        pts = np.stack([evaluate(arc, cx, xq), evaluate(arc, cy, xq)])
    length = float(arc[-1])
    n_poly = int(10 * length)
    xq = linspace(length, n_poly)
    poly = np.stack([evaluate(arc, cx, xq), evaluate(arc, cy, xq)], axis=1)
    seglen = np.sqrt((poly[1:, 0] - poly[:-1, 0]) ** 2 + (poly[1:, 1] - poly[:-1, 1]) ** 2)
    return dict(knots=arc, cx=cx, cy=cy, length=length, poly=poly, cum=np.concatenate([[0.0], np.cumsum(seglen)]))
