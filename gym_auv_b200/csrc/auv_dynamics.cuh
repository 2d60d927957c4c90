// auv_dynamics.cuh -- 3-DOF vessel model + Fehlberg step (FP64, registers).
#pragma once
#include "auv_device.cuh"
#include "../../include/auv_b200.h"

namespace auv {

// ------------------------------------------------------------------------------------
// k_vessel_step     vessel.py:226-247,561-578; odesolver.py:2-47; constants.py:33-72
// ------------------------------------------------------------------------------------
struct S6 {
  double x, y, psi, u, v, r;
};

__device__ __forceinline__ S6 state_dot(const S6& s, double tau_u, double tau_r) {
  // M = [[25.8,0,0],[0,33.8,1.0948],[0,1.0948,2.76]]   (constants.py:33-36)
  constexpr double m11 = 33.8, m12 = 23.8 * 0.046, m22 = 2.76;
  constexpr double det = m11 * m22 - m12 * m12;
  constexpr double i00 = 1.0 / 25.8, i11 = m22 / det, i12 = -m12 / det, i22 = m11 / det;
  double sp, cp;
  sincos(princip(s.psi), &sp, &cp);
  S6 d;
  d.x = cp * s.u - sp * s.v;
  d.y = sp * s.u + cp * s.v;
  d.psi = s.r;
  // tau - D nu - N(nu) nu   (constants.py:39-43, 63-72)
  const double f1 = (tau_u - 2.0 * s.u) - 2.0 * s.u;
  const double f2 = (0.0 - (7.0 * s.v - 2.5425 * s.r)) - (7.0 * s.v + (23.8 * s.u + 0.1) * s.r);
  const double f3 = (tau_r - (-2.5425 * s.v + 1.422 * s.r)) - (0.1 * s.v + (23.8 * 0.046 * s.u + 0.5) * s.r);
  d.u = i00 * f1;
  d.v = i11 * f2 + i12 * f3;
  d.r = i12 * f2 + i22 * f3;
  return d;
}

#define S6_AXPY(out, y, EXPR)          \
  out.x = y.x + (EXPR(x));             \
  out.y = y.y + (EXPR(y));             \
  out.psi = y.psi + (EXPR(psi));       \
  out.u = y.u + (EXPR(u));             \
  out.v = y.v + (EXPR(v));             \
  out.r = y.r + (EXPR(r));

// one Fehlberg step for one env (vessel.py:226-247): returns the 5th-order solution q
__device__ __forceinline__ S6 vessel_rk_step(const AuvConfig& cfg, const S6& y, float2 a) {
  if (isnan(a.x) || isnan(a.y)) a = make_float2(0.f, 0.f);  // environment.py:314-315
  const double tau_u = fmin(fmax((double)a.x, 0.0), 1.0) * cfg.thrust_max_auv;
  const double tau_r = fmin(fmax((double)a.y, -1.0), 1.0) * cfg.moment_max_auv;
  const double h = cfg.t_step_size;
  S6 t, k1, k2, k3, k4, k5, k6, q;
  k1 = state_dot(y, tau_u, tau_r);
#define E2(c) h * k1.c / 4.0
  S6_AXPY(t, y, E2)
  k2 = state_dot(t, tau_u, tau_r);
#define E3(c) 3.0 * h * k1.c / 32.0 + 9.0 * h * k2.c / 32.0
  S6_AXPY(t, y, E3)
  k3 = state_dot(t, tau_u, tau_r);
#define E4(c) 1932.0 * h * k1.c / 2197.0 - 7200.0 * h * k2.c / 2197.0 + 7296.0 * h * k3.c / 2197.0
  S6_AXPY(t, y, E4)
  k4 = state_dot(t, tau_u, tau_r);
#define E5(c) 439.0 * h * k1.c / 216.0 - 8.0 * h * k2.c + 3680.0 * h * k3.c / 513.0 - 845.0 * h * k4.c / 4104.0
  S6_AXPY(t, y, E5)
  k5 = state_dot(t, tau_u, tau_r);
#define E6(c)                                                                               \
  -8.0 * h * k1.c / 27.0 + 2 * h * k2.c - 3544.0 * h * k3.c / 2565 + 1859.0 * h * k4.c / 4104.0 - \
      11.0 * h * k5.c / 40.0
  S6_AXPY(t, y, E6)
  k6 = state_dot(t, tau_u, tau_r);
#define EQ(c)                                                                                   \
  h*(16.0 * k1.c / 135.0 + 6656.0 * k3.c / 12825.0 + 28561.0 * k4.c / 56430.0 - 9.0 * k5.c / 50.0 + \
     2.0 * k6.c / 55.0)
  S6_AXPY(q, y, EQ)
  q.psi = princip(q.psi);
  return q;
}

__device__ __forceinline__ S6 load_state(const double* st, int n, int e) {
  S6 y;
  y.x = st[e];
  y.y = st[n + e];
  y.psi = st[2 * n + e];
  y.u = st[3 * n + e];
  y.v = st[4 * n + e];
  y.r = st[5 * n + e];
  return y;
}
__device__ __forceinline__ void store_state(double* st, int n, int e, const S6& q) {
  st[e] = q.x;
  st[n + e] = q.y;
  st[2 * n + e] = q.psi;
  st[3 * n + e] = q.u;
  st[4 * n + e] = q.v;
  st[5 * n + e] = q.r;
}

}  // namespace auv
