// auv_geometry.cuh -- obstacle geometry of the LiDAR path: vessel pentagon, polygonised
// circles, world polygons, culling windows, sector pooling.
#pragma once
#include "auv_device.cuh"
#include "../../include/auv_b200.h"

namespace auv {

#define OFLAG_FILLED 1
#define OFLAG_INSIDE 2
#define OFLAG_ALLRAYS 4
#define OFLAG_PENTAGON 8
#define OFLAG_WORLD 16

// body-frame pentagon of VesselObstacle relative to its area centroid (5w/18, 0), in
// units of w     obstacles.py:175-181
__constant__ double c_pent[5][2] = {{-7.0 / 9.0, -0.5}, {-7.0 / 9.0, 0.5}, {2.0 / 9.0, 0.5},
                                    {11.0 / 9.0, 0.0},  {2.0 / 9.0, -0.5}};

// culling window (sensor.py:22-97):  a = floor((pi+beta-delta)/dth) - 1,
// b = ceil((pi+beta+delta)/dth) mod R (Python modulo); candidate(i) <=> a<=i<b or a<=i-R<b.
// lo/hi are the raw floor/ceil values.
__device__ __forceinline__ void window_from_bounds(int lo, int hi, int R, int mode, int& a, int& b,
                                                   bool& allrays) {
  allrays = false;
  if (mode == AUV_CULL_EXACT) {
    // every ray whose index is in [lo-1, hi) modulo R
    if (hi - (lo - 1) >= R) {
      allrays = true;
      a = 0;
      b = R;
    } else {
      int s = (lo - 1) % R;
      if (s < 0) s += R;
      a = s;
      b = s + (hi - (lo - 1));
      if (b > R) {  // wraps: express as a negative start
        a -= R;
        b -= R;
      }
    }
    return;
  }
  a = lo - 1;
  int m = hi % R;
  if (m < 0) m += R;
  b = m;
  if (a < -R) allrays = true;  // IndexError in the reference: defined as all rays
}

__device__ __noinline__ void cull_bounds_f64(double cx, double cy, double rho, double psi, double dth,
                                             int& lo, int& hi) {
  const double dist = fmax(1e-8, sqrt(cx * cx + cy * cy));
  const double ratio = rho / dist;
  const double delta = ratio <= 1.0 ? asin(ratio) : AUV_PI;  // np.arcsin -> nan -> pi
  const double beta = atan2(cy, cx) - psi;
  lo = (int)floor((AUV_PI + (beta - delta)) / dth);
  hi = (int)ceil((AUV_PI + (beta + delta)) / dth);
}

// FP32 fast path; falls back to FP64 whenever a quotient is within GUARD of an integer or
// the asin argument is near 1, so the integers are always those of the FP64 formula.
__device__ __forceinline__ void cull_bounds(double cx, double cy, double rho, double psi, int R,
                                            int& lo, int& hi) {
  const double dth = 2.0 * AUV_PI / (double)R;
  const float fx = (float)cx, fy = (float)cy, fr = (float)rho;
  const float dist = fmaxf(1e-8f, sqrtf(fx * fx + fy * fy));
  const float ratio = fr / dist;
  bool exact = ratio > 0.98f;  // asin' blows up near 1 and the <=1 decision itself is a threshold
  if (!exact) {
    const float inv = (float)(1.0 / dth);
    const float delta = asinf(ratio);
    const float beta = atan2f(fy, fx) - (float)psi;
    const float qlo = ((float)AUV_PI + (beta - delta)) * inv;
    const float qhi = ((float)AUV_PI + (beta + delta)) * inv;
    const float flo = floorf(qlo), chi = ceilf(qhi);
    const float GUARD = 4e-4f;  // >> atan2f/asinf error (~1e-6 rad) / dth + ulp(360)
    exact = (qlo - flo < GUARD) || (flo + 1.f - qlo < GUARD) || (chi - qhi < GUARD) || (qhi - (chi - 1.f) < GUARD);
    lo = (int)flo;
    hi = (int)chi;
  }
  if (exact) cull_bounds_f64(cx, cy, rho, psi, dth, lo, hi);
}

// LidarPreprocessor._feasibility_pooling (sensor.py:251-296) for one sector: the largest
// range d such that no opening wider than `width` exists among the rays that see farther
// than d + width.  m[0..n) are the sector's ranges (shared memory), FP64 arithmetic on the
// FP32 ranges.  Candidates are visited in increasing range order (np.argsort; ties are
// equal values so their order cannot change the result).
__device__ __forceinline__ float feasibility_pooling(const float* m, int n, double width, double theta) {
  // every FP64 operation is an explicit round-to-nearest intrinsic: the algorithm is a chain
  // of threshold tests on accumulated sums, so FMA contraction would change its decisions
  const double span = __dmul_rn(theta, (double)(n - 1));
  const double half = span / 2.0, quarter = span / 4.0;
  float prev = -1.f;
  int prev_cnt = 0;  // how many rays with value == prev have been consumed already
  float maxv = 0.f;
  for (int i = 0; i < n; ++i) maxv = fmaxf(maxv, m[i]);
  for (int it = 0; it < n; ++it) {
    // next value in sorted order: smallest > prev, or another copy of prev
    int same = 0;
    float next = INFINITY;
    for (int i = 0; i < n; ++i) {
      const float v = m[i];
      if (v == prev) ++same;
      else if (v > prev) next = fminf(next, v);
    }
    float cur;
    if (prev_cnt < same) {
      cur = prev;
      ++prev_cnt;
    } else {
      cur = next;
      prev = next;
      prev_cnt = 1;
    }
    const double dcur = (double)cur;
    const double d = __dmul_rn(dcur, theta), hd = __dmul_rn(0.5, d), ht = __dmul_rn(0.5, theta);
    const double thr = __dadd_rn(dcur, width);
    double ow = 0.0, os = 0.0, ostart = -half;
    bool found = false;
    for (int i = 0; i < n; ++i) {
      if ((double)m[i] > thr) {
        ow = __dadd_rn(ow, d);
        os = __dadd_rn(os, theta);
        if (ow > width && fabs(__dadd_rn(ostart, os / 2.0)) < quarter) found = true;
      } else {
        ow = __dadd_rn(ow, hd);
        os = __dadd_rn(os, ht);
        if (ow > width && fabs(__dadd_rn(ostart, os / 2.0)) < quarter) found = true;
        ow = 0.0;
        os = 0.0;
        ostart = __dadd_rn(-half, __dmul_rn((double)i, theta));
      }
    }
    if (!found) return fmaxf(0.f, cur);
  }
  return fmaxf(0.f, maxv);
}

// pentagon vertex k of a vessel obstacle, relative to the own-ship: base (bx,by) is the
// rotation centre (area centroid) in vessel-relative coordinates
__device__ __forceinline__ void pent_vertex(int k, double bx, double by, double w, double hx, double hy,
                                            double& vx, double& vy) {
  vx = bx + w * (hx * c_pent[k][0] - hy * c_pent[k][1]);
  vy = by + w * (hy * c_pent[k][0] + hx * c_pent[k][1]);
}

// ---- shared static world polygons (PolygonObstacle, obstacles.py:116-127): FILLED.
// distance from the own-ship to the filled polygon / crossing-number inside test, on
// vessel-relative FP32 vertices formed in FP64.  Cold (nearby refresh / own-ship within
// the enclosing circle only).
__device__ __noinline__ double world_polygon_distance(const double2* __restrict__ v, int nv, double px,
                                                         double py, bool& inside) {
  float dmin = INFINITY;
  bool in = false;
  float ax = (float)(v[0].x - px), ay = (float)(v[0].y - py);
  for (int k = 1; k < nv; ++k) {
    const float bx = (float)(v[k].x - px), by = (float)(v[k].y - py);
    dmin = fminf(dmin, pt_seg_dist_f(0.f, 0.f, ax, ay, bx, by));
    if ((ay > 0.f) != (by > 0.f)) {  // edge straddles the +x axis through the own-ship
      const float xint = ax + (0.f - ay) * (bx - ax) / (by - ay);
      if (xint > 0.f) in = !in;
    }
    ax = bx;
    ay = by;
  }
  inside = in;
  return in ? 0.0 : (double)dmin;
}

// Point.distance(obstacle.boundary) from the own-ship (vessel.py:269): min over the edges
// of the polygonised circle (ring) or 0 / min over edges for the filled vessel pentagon.
// FP32 on vessel-relative vertices formed in FP64.  Cold (only on nearby-list refresh and
// only for obstacles whose enclosing circle straddles the range limit).
__device__ __noinline__ double boundary_distance(bool pent, double cx, double cy, double bx0, double by0,
                                                 double geo, double hx, double hy, int nv_cnt,
                                                 const double2* __restrict__ unit) {
  const int ne = nv_cnt - 1;
  if (!pent) {
    // ring of a regular n-gon (n a power of two): the edge nearest to a point is the one whose
    // angular sector (seen from the centre) contains the point; its two neighbours are tested as
    // well so that rounding of the angle cannot matter -- 3 edges instead of n
    const float th = atan2f((float)-cy, (float)-cx);  // direction centre -> own-ship
    const int k0 = (int)floorf(th * ((float)ne * 0.15915494309189535f)) & (ne - 1);
    const int sh = 6 - (31 - __clz(ne));  // unit table stride 64 / ne
    float dmin = INFINITY;
    double2 un = __ldg(&unit[((k0 - 1) & (ne - 1)) << sh]);
    float pxv = (float)(cx + geo * un.x), pyv = (float)(cy + geo * un.y);
#pragma unroll
    for (int t = 0; t < 3; ++t) {
      un = __ldg(&unit[((k0 + t) & (ne - 1)) << sh]);
      const float qx = (float)(cx + geo * un.x), qy = (float)(cy + geo * un.y);
      dmin = fminf(dmin, pt_seg_dist_f(0.f, 0.f, pxv, pyv, qx, qy));
      pxv = qx;
      pyv = qy;
    }
    return (double)dmin;
  }
  float dmin = INFINITY;
  double vx, vy;
  pent_vertex(0, bx0, by0, geo, hx, hy, vx, vy);
  float pxv = (float)vx, pyv = (float)vy;
  bool allpos = true, allneg = true;
  for (int k = 1; k <= ne; ++k) {
    pent_vertex(k == ne ? 0 : k, bx0, by0, geo, hx, hy, vx, vy);
    const float qx = (float)vx, qy = (float)vy;
    dmin = fminf(dmin, pt_seg_dist_f(0.f, 0.f, pxv, pyv, qx, qy));
    const float cr = pxv * qy - pyv * qx;  // cross(prev, cur) about the vessel
    allpos = allpos && (cr >= 0.f);
    allneg = allneg && (cr <= 0.f);
    pxv = qx;
    pyv = qy;
  }
  return (allpos || allneg) ? 0.0 : (double)dmin;  // filled: own-ship inside => distance 0
}

// is the own-ship (origin) inside the convex vessel pentagon?  FP64.
__device__ __forceinline__ bool vessel_inside_pentagon(double bx0, double by0, double geo, double hx, double hy) {
  bool allpos = true, allneg = true;
  double pvx, pvy, vx, vy;
  pent_vertex(4, bx0, by0, geo, hx, hy, pvx, pvy);
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    pent_vertex(k, bx0, by0, geo, hx, hy, vx, vy);
    const double cr = pvx * vy - pvy * vx;
    allpos = allpos && (cr >= 0.0);
    allneg = allneg && (cr <= 0.0);
    pvx = vx;
    pvy = vy;
  }
  return allpos || allneg;
}

}  // namespace auv
