// auv_device.cuh -- small device helpers shared by the sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define AUV_PI 3.14159265358979323846
#define AUV_FULL 0xffffffffu

// -DAUV_DEBUG_BOUNDS: every shared-memory / scratch index of the step kernels is checked and a violation
// raises AUV_STATUS_BOUNDS in AuvBatch.status (compute-sanitizer is not available on every pool; this build
// is run once by the GPU tests).  Without the flag the checks compile to nothing.
#ifdef AUV_DEBUG_BOUNDS
#define AUV_CHECK(status, cond)                                                              \
  do {                                                                                       \
    if (!(cond) && (status) != nullptr) atomicOr((status), 16 | ((__LINE__ & 0x3fff) << 8)); \
  } while (0)
#else
#define AUV_CHECK(status, cond) \
  do {                          \
  } while (0)
#endif

namespace auv {

// geomutils.py:4-5  princip(angle) = ((angle + pi) % (2 pi)) - pi  with Python's floored
// modulo (result of % has the sign of the divisor) => value in [-pi, pi).
__device__ __forceinline__ double princip(double a) {
  // x - q * 2pi with q = floor(x / 2pi) by ONE fused multiply-add: the exact remainder is a
  // double, so the FMA returns it exactly -- the same value as fmod's, without its software loop
  // (angles here are a few radians; the guards only act when x / 2pi rounds across an integer)
  const double two_pi = 2.0 * AUV_PI;
  const double x = a + AUV_PI;
  double m = fma(-two_pi, floor(x * (1.0 / two_pi)), x);
  if (m < 0.0) m += two_pi;
  if (m >= two_pi) m -= two_pi;
  return m - AUV_PI;
}

__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(AUV_FULL, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(AUV_FULL, v, o);
  return v;
}
__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(AUV_FULL, v, o);
    if (lane >= o) v += t;
  }
  return v;
}
// lexicographic (d2, idx) arg-min over the warp; result valid in every lane
__device__ __forceinline__ void warp_argmin(double& d2, int& idx) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    double od = __shfl_xor_sync(AUV_FULL, d2, o);
    int oi = __shfl_xor_sync(AUV_FULL, idx, o);
    if (od < d2 || (od == d2 && oi < idx)) {
      d2 = od;
      idx = oi;
    }
  }
}

// distance from the origin-relative point (qx,qy) to segment (ax,ay)-(bx,by), FP32
__device__ __forceinline__ float pt_seg_dist_f(float qx, float qy, float ax, float ay, float bx,
                                               float by) {
  float ex = bx - ax, ey = by - ay;
  float wx = qx - ax, wy = qy - ay;
  float len2 = ex * ex + ey * ey;
  float t = wx * ex + wy * ey;
  t = len2 > 0.f ? fminf(fmaxf(t / len2, 0.f), 1.f) : 0.f;
  float dx = wx - t * ex, dy = wy - t * ey;
  return sqrtf(dx * dx + dy * dy);
}

// squared distance from the origin-relative point (qx,qy) to the segment a + t e, t in [0,1],
// with the precomputed inv = 1/|e|^2 (0 for a degenerate segment): no division, no sqrt
__device__ __forceinline__ float pt_chord_d2_f(float qx, float qy, float4 ch, float inv) {
  const float wx = qx - ch.x, wy = qy - ch.y;
  const float t = fminf(fmaxf((wx * ch.z + wy * ch.w) * inv, 0.f), 1.f);
  const float dx = wx - t * ch.z, dy = wy - t * ch.w;
  return dx * dx + dy * dy;
}

// circle -> regular n-gon side count of buffer(r).boundary.simplify(0.3)
// (obstacles.py:101-106; closed form SURVEY.md App. A.5, pinned against a literal
// Douglas-Peucker in oracle/geos_lite.py): n = 64/m, m the largest power of two <= 32
// with r (1 - cos(m pi / 64)) <= 0.3.
__device__ __forceinline__ int ngon_sides(double r) {
  // 1 - cos(m*pi/64) for m = 32, 16, 8, 4, 2
  if (r * 1.0 <= 0.3) return 2;
  if (r * 0.29289321881345248 <= 0.3) return 4;
  if (r * 0.07612046748871326 <= 0.3) return 8;
  if (r * 0.01921471959676957 <= 0.3) return 16;
  if (r * 0.00481527332780311 <= 0.3) return 32;
  return 64;
}

}  // namespace auv
