// auv_kernels.cu -- hand-written sm_100a kernels for the gym-auv step path.
//
// Kernel inventory (reference file:line each one replaces is in include/auv_b200.h):
//   k_obstacle_update  one thread per (env, moving-obstacle slot)            HBM-bound
//   k_vessel_step      one thread per env, FP64 RKF45 state in registers     HBM-bound
//   k_observe          one WARP per env: path projection (two-level exact search),
//                      PCHIP evaluation, nearby filter, reference-exact culling windows,
//                      obstacle vertices staged in shared memory, lanes-over-rays
//                      ray/segment casting, closeness, collision, reward (warp reduce),
//                      done, episode stats and in-kernel auto-reset.        FP32-pipe bound
//   k_reset            explicit host-requested reset of flagged envs
//   k_fma_probe        FP32 FMA peak micro-benchmark (roofline denominator)
//
// Precision plan: everything that is cheap and threshold-sensitive is FP64 (vessel
// state, RK step, obstacle positions, vessel-relative obstacle centres, culling-window
// integers, path projection refine, navigation scalars, reward scalars); the O(rays x
// segments) ray casting is FP32 on vessel-relative vertices that were formed in FP64.
#include "auv_device.cuh"
#include "../../include/auv_b200.h"

#include <math.h>

namespace auv {

// ------------------------------------------------------------------------------------
// k_obstacle_update     obstacles.py:195-215
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_obstacle_update(AuvConfig cfg, AuvScenarioPool pool,
                                                          AuvBatch batch) {
  const int km = pool.k_moving;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (long long)batch.n_envs * km) return;
  const int e = (int)(gid / km);
  const int j = (int)(gid - (long long)e * km);
  const long long ps = (long long)batch.scn_id[e] * km + j;
  const double w = pool.mov_width[ps];
  if (!(w > 0.0)) return;
  const double dt = cfg.t_step_size;
  double counter = batch.mov_counter[gid] + dt;
  int index = (int)floor(counter);
  const int4 tr = reinterpret_cast<const int4*>(pool.mov_track)[ps];  // off, len, stride
  double2 pos = reinterpret_cast<double2*>(batch.mov_pos)[gid];
  if (index >= tr.y - 1) {
    counter = 0.0;
    index = 0;
    pos = reinterpret_cast<const double2*>(pool.mov_start)[ps];
  }
  const double2 v = reinterpret_cast<const double2*>(pool.vel_table)[tr.x + (long long)index * tr.z];
  const double dx = dt * v.x, dy = dt * v.y;
  pos.x += dx;
  pos.y += dy;
  reinterpret_cast<double2*>(batch.mov_pos)[gid] = pos;
  reinterpret_cast<double2*>(batch.mov_disp)[gid] = make_double2(dx, dy);
  batch.mov_counter[gid] = counter;
}

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

__global__ void __launch_bounds__(128) k_vessel_step(AuvConfig cfg, AuvBatch batch,
                                                      const float* __restrict__ actions) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = batch.n_envs;
  if (e >= n) return;
  float2 a = reinterpret_cast<const float2*>(actions)[e];
  if (isnan(a.x) || isnan(a.y)) a = make_float2(0.f, 0.f);  // environment.py:314-315
  const double tau_u = fmin(fmax((double)a.x, 0.0), 1.0) * cfg.thrust_max_auv;
  const double tau_r = fmin(fmax((double)a.y, -1.0), 1.0) * cfg.moment_max_auv;
  const double h = cfg.t_step_size;
  double* st = batch.state;
  S6 y;
  y.x = st[e];
  y.y = st[n + e];
  y.psi = st[2 * n + e];
  y.u = st[3 * n + e];
  y.v = st[4 * n + e];
  y.r = st[5 * n + e];
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
  st[e] = q.x;
  st[n + e] = q.y;
  st[2 * n + e] = q.psi;
  st[3 * n + e] = q.u;
  st[4 * n + e] = q.v;
  st[5 * n + e] = q.r;
  batch.step_counter[e] += 1;
}

// ------------------------------------------------------------------------------------
// reset of one env, executed by one warp    environment.py:202-212, vessel.py:189-224
// ------------------------------------------------------------------------------------
__device__ __forceinline__ void reset_env_warp(const AuvScenarioPool& pool, AuvBatch& batch, int e,
                                               int scn, int lane) {
  const int n = batch.n_envs;
  const int km = pool.k_moving;
  if (lane == 0) {
    batch.scn_id[e] = scn;
    batch.episode[e] += 1;
    const double* vi = pool.vessel_init + 3ll * scn;
    batch.state[e] = vi[0];
    batch.state[n + e] = vi[1];
    batch.state[2 * n + e] = vi[2];
    batch.state[3 * n + e] = 0.0;
    batch.state[4 * n + e] = 0.0;
    batch.state[5 * n + e] = 0.0;
    batch.step_counter[e] = 0;
    batch.t_step[e] = 0;
    batch.cum_reward[e] = 0.0;
    batch.max_progress[e] = 0.0;
    batch.cte_sum[e] = 0.0;
  }
  for (int j = lane; j < km; j += 32) {
    const long long ps = (long long)scn * km + j, pe = (long long)e * km + j;
    reinterpret_cast<double2*>(batch.mov_pos)[pe] = reinterpret_cast<const double2*>(pool.mov_pos0)[ps];
    reinterpret_cast<double2*>(batch.mov_disp)[pe] = reinterpret_cast<const double2*>(pool.mov_disp0)[ps];
    batch.mov_counter[pe] = pool.mov_counter0[ps];
  }
  for (int w = lane; w < batch.mask_words; w += 32) batch.nearby_mask[(long long)e * batch.mask_words + w] = 0u;
  __syncwarp();
}

__global__ void __launch_bounds__(256) k_reset(AuvScenarioPool pool, AuvBatch batch,
                                               const uint8_t* __restrict__ reset_mask) {
  const int lane = threadIdx.x & 31;
  const int e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (e >= batch.n_envs) return;
  if (reset_mask != nullptr && reset_mask[e] == 0) return;
  reset_env_warp(pool, batch, e, batch.scn_id[e], lane);
}

// ------------------------------------------------------------------------------------
// Path projection + navigation features (warp-cooperative)
//   path.py:61-93 (PCHIP eval, LineString.project), vessel.py:461-541 (navigate)
// ------------------------------------------------------------------------------------
struct Nav {
  double s, chi, y_e, s_la, la_err, head_err, goal_dist, progress;
  bool reached;
};

// scipy PPoly evaluation (extrapolate=True): interval j with x[j] <= s < x[j+1], clamped.
__device__ __forceinline__ void pchip_eval(const AuvPathBank& pb, int pid, double s, double& px,
                                           double& py, double& dx, double& dy) {
  const int nk = pb.n_knots;
  const double* kn = pb.knots + (long long)pid * nk;
  const double L = kn[nk - 1];
  int j = (int)((s / L) * (nk - 1));
  j = max(0, min(nk - 2, j));
  while (j > 0 && s < kn[j]) --j;
  while (j < nk - 2 && s >= kn[j + 1]) ++j;
  const double t = s - kn[j];
  const double* c = pb.coef + ((long long)pid * (nk - 1) + j) * 8;
  px = ((c[0] * t + c[1]) * t + c[2]) * t + c[3];
  py = ((c[4] * t + c[5]) * t + c[6]) * t + c[7];
  dx = (3.0 * c[0] * t + 2.0 * c[1]) * t + c[2];
  dy = (3.0 * c[4] * t + 2.0 * c[5]) * t + c[6];
}

// GEOS LengthIndexOfPoint::indexOf restated as an exact two-level search: a block of 32
// consecutive 0.1 m segments lies inside the capsule (chord, dev); blocks whose capsule
// lower bound exceeds the best capsule upper bound cannot hold the minimum.  Candidate
// blocks are refined in FP64 in path order with a strict '<' so the FIRST minimum wins.
__device__ __forceinline__ double project_warp(const AuvPathBank& pb, int pid, double px, double py,
                                               int lane) {
  const int v0 = pb.poly_off[pid];
  const int nseg = pb.poly_off[pid + 1] - v0 - 1;
  const int b0 = pb.blk_off[pid];
  const int nblk = pb.blk_off[pid + 1] - b0;
  const double ox = pb.origin[2 * pid], oy = pb.origin[2 * pid + 1];
  const float qx = (float)(px - ox), qy = (float)(py - oy);
  const float pad = 1e-6f * (fabsf(qx) + fabsf(qy)) + 1e-6f;
  const float4* chord = reinterpret_cast<const float4*>(pb.blk_chord) + b0;
  const float* dev = pb.blk_dev + b0;
  float ub = INFINITY;
  for (int b = lane; b < nblk; b += 32) {
    const float4 ch = chord[b];
    const float dc = pt_seg_dist_f(qx, qy, ch.x, ch.y, ch.z, ch.w);
    ub = fminf(ub, dc * (1.f + 4e-6f) + dev[b] + pad);
  }
  ub = warp_min(ub);
  double best_d2 = INFINITY;
  int best_seg = 0;
  const double2* poly = reinterpret_cast<const double2*>(pb.poly_xy) + v0;
  for (int base = 0; base < nblk; base += 32) {
    const int b = base + lane;
    bool cand = false;
    if (b < nblk) {
      const float4 ch = chord[b];
      const float dc = pt_seg_dist_f(qx, qy, ch.x, ch.y, ch.z, ch.w);
      cand = dc * (1.f - 4e-6f) - dev[b] - pad <= ub;
    }
    unsigned m = __ballot_sync(AUV_FULL, cand);
    while (m) {
      const int bb = base + __ffs(m) - 1;
      m &= m - 1;
      const int seg = bb * AUV_PATH_BLOCK + lane;
      double d2 = INFINITY;
      if (seg < nseg) {
        const double2 A = poly[seg], B = poly[seg + 1];
        const double ex = B.x - A.x, ey = B.y - A.y;
        const double wx = px - A.x, wy = py - A.y;
        const double len2 = ex * ex + ey * ey;
        const double num = wx * ex + wy * ey;
        if (len2 == 0.0 || num <= 0.0) {
          d2 = wx * wx + wy * wy;
        } else if (num >= len2) {
          const double zx = px - B.x, zy = py - B.y;
          d2 = zx * zx + zy * zy;
        } else {
          const double cr = wx * ey - wy * ex;
          d2 = cr * cr / len2;
        }
      }
      int idx = seg;
      warp_argmin(d2, idx);
      if (d2 < best_d2) {
        best_d2 = d2;
        best_seg = idx;
      }
    }
  }
  // segmentNearestMeasure of the winning segment (uniform across the warp)
  const double2 A = poly[best_seg], B = poly[best_seg + 1];
  const double start = pb.poly_cum[v0 + best_seg];
  const double ex = B.x - A.x, ey = B.y - A.y;
  const double len2 = ex * ex + ey * ey;
  if (len2 == 0.0) return start;
  const double r = ((px - A.x) * ex + (py - A.y) * ey) / len2;
  if (r <= 0.0) return start;
  const double seglen = sqrt(len2);
  if (r <= 1.0) return start + r * seglen;
  return start + seglen;
}

__device__ __forceinline__ Nav navigate_warp(const AuvConfig& cfg, const AuvPathBank& pb, int pid,
                                             double px, double py, double psi, int lane) {
  Nav nv;
  nv.s = project_warp(pb, pid, px, py, lane);
  const double L = pb.length[pid];
  nv.s_la = fmin(L, nv.s + cfg.look_ahead_distance);
  // lanes 0/1 evaluate the spline at s / s_la in parallel, then broadcast
  double ex = 0, ey = 0, dx = 0, dy = 0;
  if (lane < 2) pchip_eval(pb, pid, lane == 0 ? nv.s : nv.s_la, ex, ey, dx, dy);
  const double p_x = __shfl_sync(AUV_FULL, ex, 0), p_y = __shfl_sync(AUV_FULL, ey, 0);
  const double d_x = __shfl_sync(AUV_FULL, dx, 0), d_y = __shfl_sync(AUV_FULL, dy, 0);
  const double l_x = __shfl_sync(AUV_FULL, ex, 1), l_y = __shfl_sync(AUV_FULL, ey, 1);
  const double ldx = __shfl_sync(AUV_FULL, dx, 1), ldy = __shfl_sync(AUV_FULL, dy, 1);
  nv.chi = atan2(d_y, d_x);
  double sc, cc;
  sincos(nv.chi, &sc, &cc);
  nv.y_e = -sc * (p_x - px) + cc * (p_y - py);
  nv.la_err = princip(atan2(ldy, ldx) - psi);
  nv.head_err = princip(atan2(l_y - py, l_x - px) - psi);
  nv.progress = nv.s / L;
  const double gx = pb.end_xy[2 * pid] - px, gy = pb.end_xy[2 * pid + 1] - py;
  nv.goal_dist = sqrt(gx * gx + gy * gy);
  nv.reached = (nv.goal_dist <= cfg.min_goal_distance) || (nv.progress >= cfg.min_path_progress);
  return nv;
}

// ------------------------------------------------------------------------------------
// LiDAR
// ------------------------------------------------------------------------------------
constexpr int VMAX = 448;          // staged vertices per warp per batch (float2)
constexpr int WARPS_PER_BLOCK = 8;

struct __align__(16) WarpScratch {
  float2 verts[VMAX];
  double ocxd[32], ocyd[32];  // enclosing-circle centre, vessel-relative (FP64)
  double ogeo[32];            // radius (circle) or width (vessel obstacle)
  double ohx[32], ohy[32];    // unit heading of a vessel obstacle
  float ocx[32], ocy[32], orho[32];
  int oa[32], ob[32], ovoff[32], onv[32], oflag[32];
};
#define OFLAG_FILLED 1
#define OFLAG_INSIDE 2
#define OFLAG_ALLRAYS 4
#define OFLAG_PENTAGON 8

// body-frame pentagon of VesselObstacle relative to its area centroid (5w/18, 0), in
// units of w     obstacles.py:175-181
__constant__ double c_pent[5][2] = {{-7.0 / 9.0, -0.5}, {-7.0 / 9.0, 0.5}, {2.0 / 9.0, 0.5},
                                    {11.0 / 9.0, 0.0},  {2.0 / 9.0, -0.5}};

// culling window (sensor.py:22-97), FP64:  a = floor((pi+beta-delta)/dth) - 1,
// b = ceil((pi+beta+delta)/dth) mod R (Python modulo); candidate(i) <=> a<=i<b or a<=i-R<b
__device__ __forceinline__ void cull_window(double cx, double cy, double rho, double psi, int R,
                                            int mode, int& a, int& b, bool& allrays) {
  const double dth = 2.0 * AUV_PI / (double)R;
  const double dist = fmax(1e-8, sqrt(cx * cx + cy * cy));
  const double ratio = rho / dist;
  const double delta = ratio <= 1.0 ? asin(ratio) : AUV_PI;
  const double beta = atan2(cy, cx) - psi;
  const int lo = (int)floor((AUV_PI + (beta - delta)) / dth);
  const int hi = (int)ceil((AUV_PI + (beta + delta)) / dth);
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

struct ObserveArgs {
  AuvConfig cfg;
  AuvRayTable rays;
  AuvPathBank paths;
  AuvScenarioPool pool;
  AuvBatch batch;
  AuvStepOut out;
  int mode;
  int obs_dim;
};

__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
    k_observe(const __grid_constant__ ObserveArgs A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double s_unit[64][2];  // cos/sin(2 pi k / 64)
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const int R = A.cfg.n_sensors;
  const int rpad = (R + 31) & ~31;
  const size_t per_warp = sizeof(WarpScratch) + sizeof(float) * rpad;
  WarpScratch& W = *reinterpret_cast<WarpScratch*>(smem_raw + per_warp * wib);
  float* sdist = reinterpret_cast<float*>(smem_raw + per_warp * wib + sizeof(WarpScratch));
  if (threadIdx.x < 64) {
    double s, c;
    sincospi((double)threadIdx.x / 32.0, &s, &c);
    s_unit[threadIdx.x][0] = c;
    s_unit[threadIdx.x][1] = s;
  }
  __syncthreads();
  const int e = blockIdx.x * WARPS_PER_BLOCK + wib;
  const int n = A.batch.n_envs;
  if (e >= n) return;
  AuvBatch batch = A.batch;  // local copy (reset helper takes a non-const ref)
  const AuvScenarioPool& pool = A.pool;
  const AuvConfig& cfg = A.cfg;
  const int km = pool.k_moving, ks = pool.k_static, K = km + ks;
  const double range = cfg.sensor_range;
  const float rangef = (float)range;
  const float widthf = (float)cfg.vessel_width;
  int mode = A.mode;

  for (int pass = 0; pass < 2; ++pass) {
    const int scn = batch.scn_id[e];
    const double px = batch.state[e], py = batch.state[n + e], psi = batch.state[2 * n + e];
    const double vu = batch.state[3 * n + e], vv = batch.state[4 * n + e], vr = batch.state[5 * n + e];
    const int step_counter = batch.step_counter[e];

    // ---------------- navigate ----------------
    const Nav nv = navigate_warp(cfg, A.paths, pool.path_id[scn], px, py, psi, lane);
    const double maxprog_prev = batch.max_progress[e];
    const double maxprog = fmax(nv.progress, maxprog_prev);

    float* obs = A.out.obs + (long long)e * A.obs_dim;
    bool collision = false;
    float pen_sum = 0.f;       // sum_i w_i * range * exp(-0.1 d_i)
    unsigned long long ntests = 0;

    if (cfg.use_lidar) {
      // ---------------- perceive ----------------
      for (int i = lane; i < rpad; i += 32) sdist[i] = rangef;
      const bool refresh = (step_counter % cfg.sensor_interval_load_obstacles) == 0;
      double spsi, cpsi;
      sincos(psi, &spsi, &cpsi);
      bool any_active = false;
      for (int base = 0; base < K; base += 32) {
        const int j = base + lane;
        bool valid = false;
        bool pent = false;
        double cx = 0, cy = 0, rho = 0, geo = 0, hx = 1.0, hy = 0.0;
        int nv_cnt = 0;  // vertices incl. closing one
        if (j < km) {
          const long long ps = (long long)scn * km + j, pe = (long long)e * km + j;
          const double w = pool.mov_width[ps];
          if (w > 0.0) {
            valid = true;
            pent = true;
            geo = w;
            const double2 pos = reinterpret_cast<const double2*>(batch.mov_pos)[pe];
            const double2 dsp = reinterpret_cast<const double2*>(batch.mov_disp)[pe];
            const double dl = sqrt(dsp.x * dsp.x + dsp.y * dsp.y);
            if (dl > 0.0) {
              hx = dsp.x / dl;
              hy = dsp.y / dl;
            }
            // enclosing circle of the min-rotated rectangle (obstacles.py:230-262; App. A.3)
            cx = (pos.x - px) + 5.0 * w / 18.0 + (2.0 * w / 9.0) * hx;
            cy = (pos.y - py) + (2.0 * w / 9.0) * hy;
            rho = w * 1.1180339887498949;  // sqrt(5)/2
            nv_cnt = 6;
          }
        } else if (j < K) {
          const long long ps = (long long)scn * ks + (j - km);
          const double r = pool.st_radius[ps];
          if (r > 0.0) {
            valid = true;
            geo = r;
            const double2 c = reinterpret_cast<const double2*>(pool.st_pos)[ps];
            cx = c.x - px;
            cy = c.y - py;
            rho = r;
            nv_cnt = ngon_sides(r) + 1;
          }
        }
        // pentagon vertices relative to the vessel (needed for inside / nearby tests)
        // V_k = centroid + R(heading) (P_k - centroid) + pos - p0
        const double bx0 = cx - (2.0 * geo / 9.0) * hx, by0 = cy - (2.0 * geo / 9.0) * hy;

        // ---- nearby list: refreshed every sensor_interval_load_obstacles vessel steps
        unsigned word;
        if (refresh) {
          bool near = false;
          if (valid) {
            float dmin = INFINITY;
            bool inside = false;
            const int ne = nv_cnt - 1;
            float pxv, pyv;
            {  // vertex 0
              double vx, vy;
              if (pent) {
                vx = bx0 + geo * (hx * c_pent[0][0] - hy * c_pent[0][1]);
                vy = by0 + geo * (hy * c_pent[0][0] + hx * c_pent[0][1]);
              } else {
                vx = cx + geo;
                vy = cy;
              }
              pxv = (float)vx;
              pyv = (float)vy;
            }
            bool allpos = true, allneg = true;
            for (int k = 1; k <= ne; ++k) {
              const int kk = (k == ne) ? 0 : k;
              double vx, vy;
              if (pent) {
                vx = bx0 + geo * (hx * c_pent[kk][0] - hy * c_pent[kk][1]);
                vy = by0 + geo * (hy * c_pent[kk][0] + hx * c_pent[kk][1]);
              } else {
                const int ui = kk * (64 / ne);
                vx = cx + geo * s_unit[ui][0];
                vy = cy + geo * s_unit[ui][1];
              }
              const float qx = (float)vx, qy = (float)vy;
              dmin = fminf(dmin, pt_seg_dist_f(0.f, 0.f, pxv, pyv, qx, qy));
              const float cr = pxv * qy - pyv * qx;  // cross(prev, cur) about the vessel
              allpos = allpos && (cr >= 0.f);
              allneg = allneg && (cr <= 0.f);
              pxv = qx;
              pyv = qy;
            }
            inside = pent && (allpos || allneg);
            const double dist = inside ? 0.0 : (double)dmin;
            near = (dist - cfg.vessel_width) < range;  // vessel.py:269-270
          }
          word = __ballot_sync(AUV_FULL, near);
          if (lane == 0) batch.nearby_mask[(long long)e * batch.mask_words + (base >> 5)] = word;
        } else {
          word = batch.nearby_mask[(long long)e * batch.mask_words + (base >> 5)];
        }
        const bool active = valid && ((word >> lane) & 1u);

        int wa = 0, wb = 0;
        bool allrays = false, inside = false;
        if (active) {
          cull_window(cx, cy, rho, psi, R, cfg.cull_mode, wa, wb, allrays);
          if (pent) {  // filled polygon: is the vessel inside?  (range 0)
            bool allpos = true, allneg = true;
            double pvx = bx0 + geo * (hx * c_pent[4][0] - hy * c_pent[4][1]);
            double pvy = by0 + geo * (hy * c_pent[4][0] + hx * c_pent[4][1]);
#pragma unroll
            for (int k = 0; k < 5; ++k) {
              const double vx = bx0 + geo * (hx * c_pent[k][0] - hy * c_pent[k][1]);
              const double vy = by0 + geo * (hy * c_pent[k][0] + hx * c_pent[k][1]);
              const double cr = pvx * vy - pvy * vx;
              allpos = allpos && (cr >= 0.0);
              allneg = allneg && (cr <= 0.0);
              pvx = vx;
              pvy = vy;
            }
            inside = allpos || allneg;
          }
        }
        if (A.out.windows != nullptr && j < K) {
          int2 wv = active ? make_int2(wa, wb) : make_int2(0, 0);
          reinterpret_cast<int2*>(A.out.windows)[(long long)e * K + j] = wv;
        }

        // ---- stage active obstacles in batches bounded by the vertex budget
        unsigned rem = __ballot_sync(AUV_FULL, active);
        while (rem) {
          any_active = true;
          const bool mine = (rem >> lane) & 1u;
          const int cnt = mine ? nv_cnt : 0;
          const int incl = warp_incl_scan(cnt, lane);
          const bool take = mine && incl <= VMAX;
          const unsigned tk = __ballot_sync(AUV_FULL, take);
          const int nact = __popc(tk);
          if (take) {
            const int ci = __popc(tk & ((1u << lane) - 1u));
            W.ocxd[ci] = cx;
            W.ocyd[ci] = cy;
            W.ogeo[ci] = geo;
            W.ohx[ci] = hx;
            W.ohy[ci] = hy;
            W.ocx[ci] = (float)cx;
            W.ocy[ci] = (float)cy;
            W.orho[ci] = (float)rho;
            W.oa[ci] = wa;
            W.ob[ci] = wb;
            W.ovoff[ci] = incl - cnt;
            W.onv[ci] = nv_cnt;
            W.oflag[ci] = (pent ? (OFLAG_FILLED | OFLAG_PENTAGON) : 0) | (inside ? OFLAG_INSIDE : 0) |
                          (allrays ? OFLAG_ALLRAYS : 0);
          }
          rem &= ~tk;
          __syncwarp();
          // vertices: vessel-relative, formed in FP64, stored FP32
          for (int i = 0; i < nact; ++i) {
            const int nvv = W.onv[i], off = W.ovoff[i];
            const double g = W.ogeo[i];
            if (W.oflag[i] & OFLAG_PENTAGON) {
              if (lane < 6) {
                const int kk = lane == 5 ? 0 : lane;
                const double h_x = W.ohx[i], h_y = W.ohy[i];
                const double b_x = W.ocxd[i] - (2.0 * g / 9.0) * h_x;
                const double b_y = W.ocyd[i] - (2.0 * g / 9.0) * h_y;
                W.verts[off + lane] =
                    make_float2((float)(b_x + g * (h_x * c_pent[kk][0] - h_y * c_pent[kk][1])),
                                (float)(b_y + g * (h_y * c_pent[kk][0] + h_x * c_pent[kk][1])));
              }
            } else {
              const int ne = nvv - 1;
              const int stride = 64 / ne;
              for (int k = lane; k < nvv; k += 32) {
                const int ui = (k == ne ? 0 : k) * stride;
                W.verts[off + k] = make_float2((float)(W.ocxd[i] + g * s_unit[ui][0]),
                                               (float)(W.ocyd[i] + g * s_unit[ui][1]));
              }
            }
          }
          __syncwarp();
          // ---- rays: lanes over rays, loop over staged obstacles
          for (int i0 = 0; i0 < R; i0 += 32) {
            const int i = i0 + lane;
            if (i < R) {
              const double2 cs = reinterpret_cast<const double2*>(A.rays.cos_sin)[i];
              const float c = (float)(cs.x * cpsi - cs.y * spsi);
              const float s = (float)(cs.y * cpsi + cs.x * spsi);
              float best = sdist[i];
              for (int o = 0; o < nact; ++o) {
                const int oa = W.oa[o], ob = W.ob[o], fl = W.oflag[o];
                // candidate(i) <=> a<=i<b or a<=i-R<b (Python negative-index wrap); an obstacle
                // whose window is wider than R is listed -- and tested -- twice by the reference
                const int hits = ((oa <= i && i < ob) ? 1 : 0) + ((oa <= i - R && i - R < ob) ? 1 : 0);
                const bool inwin = (fl & OFLAG_ALLRAYS) || hits > 0;
                if (!inwin) continue;
                const int nvv = W.onv[o];
                ntests += (unsigned)((nvv - 1) * max(hits, 1));
                if (fl & OFLAG_INSIDE) {
                  best = 0.f;
                  continue;
                }
                const float ocx = W.ocx[o], ocy = W.ocy[o], rho = W.orho[o];
                const float tc = ocx * c + ocy * s;
                const float hc = ocy * c - ocx * s;
                const float slack = rho * 1e-5f + 1e-4f;
                if (fabsf(hc) > rho + slack || tc + rho + slack < 0.f || tc - rho - slack > best) continue;
                const float2* vp = W.verts + W.ovoff[o];
                float2 v = vp[0];
                float xp = v.x * c + v.y * s;
                float yp = v.y * c - v.x * s;
                for (int k = 1; k < nvv; ++k) {
                  v = vp[k];
                  const float xc = v.x * c + v.y * s;
                  const float yc = v.y * c - v.x * s;
                  if ((yp <= 0.f && yc >= 0.f) || (yp >= 0.f && yc <= 0.f)) {
                    const float t = xp + (xc - xp) * (yp / (yp - yc));
                    if (t >= 0.f && t <= rangef) best = fminf(best, t);
                  }
                  xp = xc;
                  yp = yc;
                }
              }
              sdist[i] = best;
            }
          }
          __syncwarp();
        }
      }
      // ---- closeness / collision / penalty  (vessel.py:88-95,356-359; rewarder.py:199-214)
      const float inv_log = 1.f / log1pf(rangef);
      for (int i0 = 0; i0 < R; i0 += 32) {
        const int i = i0 + lane;
        if (i < R) {
          const float d = any_active ? sdist[i] : rangef;
          float cl;
          if (!any_active) {
            cl = 0.f;  // vessel.py:275-305: no nearby obstacles => closeness 0
          } else if (d >= rangef) {
            cl = 0.f;  // 1 - log(1+range)/log(1+range) is exactly 0 in the reference
          } else if (cfg.sensor_log_transform) {
            cl = 1.f - fminf(fmaxf(log1pf(d) * inv_log, 0.f), 1.f);
          } else {
            cl = 1.f - fminf(fmaxf(d / rangef, 0.f), 1.f);
          }
          obs[6 + i] = fminf(fmaxf(cl, -1.f), 1.f);
          if (cfg.sensor_use_velocity_observations) {
            obs[6 + R + i] = 0.f;  // sensor.py:159: speed channel is (0,0) at HEAD
            obs[6 + 2 * R + i] = 0.f;
          }
          if (A.out.lidar_dist != nullptr) A.out.lidar_dist[(long long)e * R + i] = d;
          collision = collision || (any_active && d < widthf);
          pen_sum += A.rays.weight[i] * rangef * __expf(-0.1f * d);
        }
      }
      collision = __any_sync(AUV_FULL, collision);
      pen_sum = warp_sum(pen_sum);
    }

    // ---------------- navigation part of the observation (vessel.py:518-539) ----------
    if (lane == 0) {
      obs[0] = (float)fmin(fmax(vu, -1.0), 1.0);
      obs[1] = (float)fmin(fmax(vv, -1.0), 1.0);
      obs[2] = (float)fmin(fmax(vr, -1.0), 1.0);
      obs[3] = (float)fmin(fmax(nv.la_err, -1.0), 1.0);
      obs[4] = (float)fmin(fmax(nv.head_err, -1.0), 1.0);
      obs[5] = (float)fmin(fmax(nv.y_e / 100.0, -1.0), 1.0);
      batch.max_progress[e] = maxprog;
      if (A.out.nav != nullptr) {
        double* o = A.out.nav + 8ll * e;
        o[0] = nv.s;
        o[1] = nv.chi;
        o[2] = nv.y_e;
        o[3] = nv.s_la;
        o[4] = nv.la_err;
        o[5] = nv.head_err;
        o[6] = nv.goal_dist;
        o[7] = nv.progress;
      }
    }
    if (A.out.seg_tests != nullptr) {
      unsigned long long t = ntests;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(AUV_FULL, t, o);
      if (lane == 0 && t) atomicAdd(A.out.seg_tests, t);
    }
    if (mode == AUV_OBSERVE_RESET) {
      if (lane == 0 && pass == 0) {  // explicit reset observe: info mirrors a fresh env
        if (A.out.collision) A.out.collision[e] = collision;
        if (A.out.reached_goal) A.out.reached_goal[e] = nv.reached;
        if (A.out.goal_distance) A.out.goal_distance[e] = (float)nv.goal_dist;
        if (A.out.progress) A.out.progress[e] = (float)nv.progress;
      }
      return;
    }

    // ---------------- reward (rewarder.py) + done (environment.py:375-384) ------------
    const double speed = sqrt(vu * vu + vv * vv);
    double reward;
    if (collision) {
      reward = -10000.0 * (1.0 - 0.5);
    } else {
      const double cte = nv.y_e / 100.0;
      double path_reward = (1.0 + cos(nv.head_err) * speed / 2.0) * (1.0 + exp(-5.0 * fabs(cte))) - 1.0;
      const double living = 0.5 * (2.0 * 0.05 + 1.0);
      if (cfg.rewarder == AUV_REWARDER_COLAV) {
        // without LiDAR every ray keeps its reset value sensor_range (vessel.py:206-208)
        const double closeness_reward =
            cfg.use_lidar ? -(double)pen_sum / A.rays.weight_sum : -range * exp(-0.1 * range);
        if (nv.progress < maxprog) path_reward = fmin(path_reward, 0.0);
        const double slow = speed < 0.04 ? -2.0 : 0.0;
        reward = 0.5 * path_reward + 0.5 * closeness_reward - living - 10.0 * fabs(vr) + slow;
        if (reward < 0.0) reward *= 2.0;
      } else {
        const double slow = speed < 0.1 ? -2.0 : 0.0;
        reward = path_reward - living - 10.0 * fabs(vr) + slow;
      }
    }
    const double cum = batch.cum_reward[e] + reward;
    const int t_step = batch.t_step[e];
    const bool done = collision || nv.reached ||
                      (!cfg.test_mode && t_step >= cfg.max_timesteps - 1) ||
                      (!cfg.test_mode && cum < cfg.min_cumulative_reward);
    const double cte_sum = batch.cte_sum[e] + fabs(nv.y_e);
    if (lane == 0) {
      batch.cum_reward[e] = cum;
      batch.t_step[e] = t_step + 1;
      batch.cte_sum[e] = cte_sum;
      A.out.reward[e] = (float)reward;
      A.out.done[e] = done;
      if (A.out.collision) A.out.collision[e] = collision;
      if (A.out.reached_goal) A.out.reached_goal[e] = nv.reached;
      if (A.out.goal_distance) A.out.goal_distance[e] = (float)nv.goal_dist;
      if (A.out.progress) A.out.progress[e] = (float)nv.progress;
    }
    if (!(done && cfg.auto_reset)) return;

    // ---------------- auto-reset (VecEnv semantics) ------------------------------------
    __syncwarp();
    if (A.out.terminal_obs != nullptr) {
      float* to = A.out.terminal_obs + (long long)e * A.obs_dim;
      for (int i = lane; i < A.obs_dim; i += 32) to[i] = obs[i];
    }
    if (lane == 0 && A.out.stats != nullptr) {  // env.history entry, environment.py:476-489
      double* st = A.out.stats;
      atomicAdd(st + AUV_STAT_EPISODES, 1.0);
      atomicAdd(st + AUV_STAT_REWARD, cum);
      atomicAdd(st + AUV_STAT_REWARD_SQ, cum * cum);
      atomicAdd(st + AUV_STAT_PROGRESS, nv.progress);
      atomicAdd(st + AUV_STAT_COLLISIONS, collision ? 1.0 : 0.0);
      atomicAdd(st + AUV_STAT_REACHED_GOAL, nv.reached ? 1.0 : 0.0);
      atomicAdd(st + AUV_STAT_TIMESTEPS, (double)(t_step + 1));
      atomicAdd(st + AUV_STAT_CROSS_TRACK, cte_sum / (double)(t_step + 1));
      atomicAdd(st + AUV_STAT_PATHLENGTH, A.paths.length[pool.path_id[scn]]);
    }
    __syncwarp();
    const int next = (int)(((long long)scn + n) % pool.n_scenarios);
    reset_env_warp(pool, batch, e, next, lane);
    mode = AUV_OBSERVE_RESET;
  }
}

// ------------------------------------------------------------------------------------
// FP32 FMA peak probe: 8 independent FMA chains per thread
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_fma_probe(float* sink, int iters) {
  float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f;
  float a4 = a0 + 4.f, a5 = a0 + 5.f, a6 = a0 + 6.f, a7 = a0 + 7.f;
  const float m = 0.999f, c = 1e-3f;
  for (int i = 0; i < iters; ++i) {
    a0 = fmaf(a0, m, c);
    a1 = fmaf(a1, m, c);
    a2 = fmaf(a2, m, c);
    a3 = fmaf(a3, m, c);
    a4 = fmaf(a4, m, c);
    a5 = fmaf(a5, m, c);
    a6 = fmaf(a6, m, c);
    a7 = fmaf(a7, m, c);
  }
  const float r = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  if (r == 12345.678f) sink[0] = r;  // never true; keeps the chains alive
}

}  // namespace auv

// ======================================================================================
// C ABI
// ======================================================================================
#include <stdio.h>
#include <string.h>

static thread_local char g_err[512] = "";

static int set_err(int code, const char* msg) {
  snprintf(g_err, sizeof(g_err), "%s", msg);
  return code;
}
static int cuda_check(cudaError_t e, const char* where) {
  if (e == cudaSuccess) return 0;
  snprintf(g_err, sizeof(g_err), "%s: %s", where, cudaGetErrorString(e));
  return (int)e;
}

extern "C" {

int auv_abi_version(void) { return AUV_ABI_VERSION; }
int auv_sizeof(int which) {
  switch (which) {
    case 0: return (int)sizeof(AuvConfig);
    case 1: return (int)sizeof(AuvRayTable);
    case 2: return (int)sizeof(AuvPathBank);
    case 3: return (int)sizeof(AuvScenarioPool);
    case 4: return (int)sizeof(AuvBatch);
    case 5: return (int)sizeof(AuvStepOut);
    default: return AUV_EINVAL;
  }
}
const char* auv_last_error(void) { return g_err; }

int auv_obs_dim(const AuvConfig* cfg) {
  if (!cfg) return AUV_EINVAL;
  int d = 6;
  if (cfg->use_lidar) {
    d += cfg->n_sensors;
    if (cfg->sensor_use_velocity_observations) d += 2 * cfg->n_sensors;
  }
  return d;
}

static int check_cfg(const AuvConfig* cfg) {
  if (!cfg) return set_err(AUV_EINVAL, "cfg is NULL");
  if (cfg->use_lidar && (cfg->n_sensors <= 0 || cfg->n_sensors > AUV_MAX_RAYS))
    return set_err(AUV_EINVAL, "n_sensors out of range");
  if (!(cfg->t_step_size > 0.0)) return set_err(AUV_EINVAL, "t_step_size must be > 0");
  if (cfg->sensor_interval_load_obstacles <= 0)
    return set_err(AUV_EINVAL, "sensor_interval_load_obstacles must be > 0");
  return 0;
}

int auv_obstacle_update(const AuvConfig* cfg, const AuvScenarioPool* pool, AuvBatch* batch,
                        void* stream) {
  if (int rc = check_cfg(cfg)) return rc;
  if (!pool || !batch) return set_err(AUV_EINVAL, "pool/batch is NULL");
  if (batch->n_envs <= 0) return set_err(AUV_EINVAL, "n_envs must be > 0");
  const long long total = (long long)batch->n_envs * pool->k_moving;
  if (total == 0) return 0;
  const int threads = 256;
  const long long blocks = (total + threads - 1) / threads;
  auv::k_obstacle_update<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(*cfg, *pool, *batch);
  return cuda_check(cudaGetLastError(), "k_obstacle_update");
}

int auv_vessel_step(const AuvConfig* cfg, AuvBatch* batch, const float* actions, void* stream) {
  if (int rc = check_cfg(cfg)) return rc;
  if (!batch || !actions) return set_err(AUV_EINVAL, "batch/actions is NULL");
  if (batch->n_envs <= 0) return set_err(AUV_EINVAL, "n_envs must be > 0");
  const int threads = 128;
  const int blocks = (batch->n_envs + threads - 1) / threads;
  auv::k_vessel_step<<<blocks, threads, 0, (cudaStream_t)stream>>>(*cfg, *batch, actions);
  return cuda_check(cudaGetLastError(), "k_vessel_step");
}

int auv_reset(const AuvConfig* cfg, const AuvScenarioPool* pool, AuvBatch* batch,
              const uint8_t* reset_mask, void* stream) {
  if (int rc = check_cfg(cfg)) return rc;
  if (!pool || !batch) return set_err(AUV_EINVAL, "pool/batch is NULL");
  if (batch->n_envs <= 0) return set_err(AUV_EINVAL, "n_envs must be > 0");
  const int threads = 256;
  const long long blocks = ((long long)batch->n_envs * 32 + threads - 1) / threads;
  auv::k_reset<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(*pool, *batch, reset_mask);
  return cuda_check(cudaGetLastError(), "k_reset");
}

int auv_observe(const AuvConfig* cfg, const AuvRayTable* rays, const AuvPathBank* paths,
                const AuvScenarioPool* pool, AuvBatch* batch, AuvStepOut* out, int mode,
                void* stream) {
  if (int rc = check_cfg(cfg)) return rc;
  if (!paths || !pool || !batch || !out) return set_err(AUV_EINVAL, "NULL argument");
  if (cfg->use_lidar && !rays) return set_err(AUV_EINVAL, "rays is NULL with use_lidar");
  if (!out->obs) return set_err(AUV_EINVAL, "out.obs is NULL");
  if (mode == AUV_OBSERVE_STEP && (!out->reward || !out->done))
    return set_err(AUV_EINVAL, "out.reward/out.done is NULL");
  if (batch->n_envs <= 0) return set_err(AUV_EINVAL, "n_envs must be > 0");
  if (pool->k_moving + pool->k_static > AUV_MAX_OBSTACLES)
    return set_err(AUV_EINVAL, "too many obstacle slots");
  if (batch->mask_words * 32 < pool->k_moving + pool->k_static)
    return set_err(AUV_EINVAL, "mask_words too small");
  auv::ObserveArgs args;
  args.cfg = *cfg;
  if (rays) args.rays = *rays; else memset(&args.rays, 0, sizeof(args.rays));
  args.paths = *paths;
  args.pool = *pool;
  args.batch = *batch;
  args.out = *out;
  args.mode = mode;
  args.obs_dim = auv_obs_dim(cfg);
  const int rpad = cfg->use_lidar ? ((cfg->n_sensors + 31) & ~31) : 32;
  const size_t smem = (sizeof(auv::WarpScratch) + sizeof(float) * rpad) * auv::WARPS_PER_BLOCK;
  static size_t configured = 0;
  if (smem > configured) {
    if (int rc = cuda_check(cudaFuncSetAttribute(auv::k_observe, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 (int)smem),
                            "cudaFuncSetAttribute(k_observe)"))
      return rc;
    configured = smem;
  }
  const int blocks = (batch->n_envs + auv::WARPS_PER_BLOCK - 1) / auv::WARPS_PER_BLOCK;
  auv::k_observe<<<blocks, auv::WARPS_PER_BLOCK * 32, smem, (cudaStream_t)stream>>>(args);
  return cuda_check(cudaGetLastError(), "k_observe");
}

int auv_step(const AuvConfig* cfg, const AuvRayTable* rays, const AuvPathBank* paths,
             const AuvScenarioPool* pool, AuvBatch* batch, const float* actions, AuvStepOut* out,
             void* stream) {
  if (int rc = auv_obstacle_update(cfg, pool, batch, stream)) return rc;
  if (int rc = auv_vessel_step(cfg, batch, actions, stream)) return rc;
  return auv_observe(cfg, rays, paths, pool, batch, out, AUV_OBSERVE_STEP, stream);
}

int auv_step_host(const AuvConfig* cfg, const AuvRayTable* rays, const AuvPathBank* paths,
                  const AuvScenarioPool* pool, AuvBatch* batch, const float* actions_host,
                  float* actions_dev, AuvStepOut* out, float* obs_host, float* reward_host,
                  uint8_t* done_host, void* stream) {
  if (!cfg || !batch || !out || !actions_host || !actions_dev || !obs_host || !reward_host || !done_host)
    return set_err(AUV_EINVAL, "NULL argument");
  cudaStream_t s = (cudaStream_t)stream;
  const size_t n = (size_t)batch->n_envs;
  if (int rc = cuda_check(cudaMemcpyAsync(actions_dev, actions_host, n * 2 * sizeof(float),
                                          cudaMemcpyHostToDevice, s), "H2D actions"))
    return rc;
  if (int rc = auv_step(cfg, rays, paths, pool, batch, actions_dev, out, stream)) return rc;
  const size_t od = (size_t)auv_obs_dim(cfg);
  if (int rc = cuda_check(cudaMemcpyAsync(obs_host, out->obs, n * od * sizeof(float),
                                          cudaMemcpyDeviceToHost, s), "D2H obs"))
    return rc;
  if (int rc = cuda_check(cudaMemcpyAsync(reward_host, out->reward, n * sizeof(float),
                                          cudaMemcpyDeviceToHost, s), "D2H reward"))
    return rc;
  if (int rc = cuda_check(cudaMemcpyAsync(done_host, out->done, n, cudaMemcpyDeviceToHost, s), "D2H done"))
    return rc;
  return cuda_check(cudaStreamSynchronize(s), "sync");
}

int auv_fma_probe(float* sink, int blocks, int threads, int iters, void* stream, double* flops_out) {
  if (!sink || blocks <= 0 || threads <= 0 || threads > 256 || iters <= 0)
    return set_err(AUV_EINVAL, "bad probe arguments");
  auv::k_fma_probe<<<blocks, threads, 0, (cudaStream_t)stream>>>(sink, iters);
  if (flops_out) *flops_out = 2.0 * 8.0 * (double)iters * (double)blocks * (double)threads;
  return cuda_check(cudaGetLastError(), "k_fma_probe");
}

}  // extern "C"
