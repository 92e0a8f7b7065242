// Device-side bodies of the batched 8-state Kalman filter (fp64 state), shared by the stand-alone
// kernels (kalman.cu) and the per-frame fused launches (frame_kernels.cu).  Reference arithmetic:
// KalmanFilter of demo:118-336 (demo = /root/reference/demo_bottrack_onnx_tflite.py).
//
// Layout: AoS per track, mean[t][8] and cov[t][8][8] float64, contiguous per track.  Eight lanes own
// one track (lane r <-> state row r): they read one 64 B mean line and eight 64 B covariance rows =
// one contiguous 576 B span; 4 tracks per warp; cross-row terms move by warp shuffle.  Every function
// here must be called by all 32 lanes of the warp (inactive groups pass active = false).
#pragma once
#include "common.cuh"

__device__ __forceinline__ double btd_shfl(double v, int src_lane) { return __shfl_sync(0xffffffffu, v, src_lane); }

// Writes the cached tlbr of a track from its (new) mean held one component per lane.
// STrack.tlwh / .tlbr, demo:624-648: x1 = cx - w/2, x2 = w + x1 (same op order, fp64).
__device__ __forceinline__ void btd_store_tlbr(double m, int lane, int r, bool active, size_t t,
                                               double* __restrict__ tlbr, float* __restrict__ tlbr_f32) {
  const int base = lane & ~7;
  const double c = btd_shfl(m, base + (r & 1));
  const double wh = btd_shfl(m, base + 2 + (r & 1));
  const double lo = c - wh / 2;
  const double val = (r < 2) ? lo : (wh + lo);
  if (active && r < 4) {
    if (tlbr) tlbr[t * 4 + r] = val;
    // conservative fp32 interval for the fast no-overlap test of the association epilogue
    if (tlbr_f32) tlbr_f32[t * 4 + r] = (r < 2) ? __double2float_rd(val) : __double2float_ru(val);
  }
}

// KalmanFilter.multi_predict (demo:265-302) + STrack.multi_predict's velocity reset (demo:529-532) of
// track t (reset: state != Tracked).
__device__ __forceinline__ void btd_predict(double* __restrict__ mean, double* __restrict__ cov,
                                            double* __restrict__ tlbr, float* __restrict__ tlbr_f32, size_t t,
                                            bool active, bool reset_vel, int noise_f32,
                                            uint8_t* __restrict__ slot_f32, int lane) {
  const int r = lane & 7;
  const int base = lane & ~7;
  double m = 0.0;
  double c[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) c[j] = 0.0;
  if (active) {
    m = mean[t * 8 + r];
    const double2* row = reinterpret_cast<const double2*>(cov + t * 64 + r * 8);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      double2 v = row[j];
      c[2 * j] = v.x;
      c[2 * j + 1] = v.y;
    }
    if (reset_vel && r >= 6) m = 0.0;  // demo:529-532
  }
  // process noise from the PRE-predict w,h (demo:281-291)
  const double w = btd_shfl(m, base + 2);
  const double h = btd_shfl(m, base + 3);
  const double wh = (r & 1) ? h : w;
  double q;
  if (noise_f32) {
    // NumPy evaluates std and its square in float32 when every pooled mean is float32
    const float wt = (r < 4) ? (float)BT_STD_POS : (float)BT_STD_VEL;
    const float s = __fmul_rn(wt, (float)wh);
    q = (double)__fmul_rn(s, s);
  } else {
    const double s = ((r < 4) ? BT_STD_POS : BT_STD_VEL) * wh;
    q = s * s;
  }
  // mean <- mean F^T : positions += velocities
  const double m_hi = btd_shfl(m, base + ((r + 4) & 7));
  if (r < 4) m = m + m_hi;
  // P <- F P F^T: rows 0..3 += rows 4..7, then cols 0..3 += cols 4..7 (same association order
  // as the two np.dot calls of demo:299-300)
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const double o = btd_shfl(c[j], base + ((r + 4) & 7));
    if (r < 4) c[j] = c[j] + o;
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) c[j] = c[j] + c[j + 4];
#pragma unroll
  for (int j = 0; j < 8; ++j)
    if (j == r) c[j] += q;

  if (active) {
    mean[t * 8 + r] = m;
    double2* row = reinterpret_cast<double2*>(cov + t * 64 + r * 8);
#pragma unroll
    for (int j = 0; j < 4; ++j) row[j] = make_double2(c[2 * j], c[2 * j + 1]);
    if (slot_f32 && r == 0) slot_f32[t] = 0;   // the state is float64 from now on
  }
  btd_store_tlbr(m, lane, r, active, t, tlbr, tlbr_f32);
}

// KalmanFilter.update (demo:304-336; project = demo:236-263) of track t with measurement z = meas[zi].
// f32_noise: the projection noise is evaluated in float32 (never-predicted float32 state, NumPy >= 2).
__device__ __forceinline__ void btd_update(double* __restrict__ mean, double* __restrict__ cov,
                                           double* __restrict__ tlbr, float* __restrict__ tlbr_f32,
                                           const double* __restrict__ meas, size_t t, size_t zi, bool active,
                                           bool f32_noise, double* __restrict__ res_tlbr, size_t res_t, int lane) {
  const int r = lane & 7;
  const int base = lane & ~7;
  double m = 0.0;
  double c[8];
  double S[4][4];
  double z[4] = {0, 0, 0, 0};
#pragma unroll
  for (int j = 0; j < 8; ++j) c[j] = 0.0;
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) S[a][b] = (a == b) ? 1.0 : 0.0;
  if (active) {
    m = mean[t * 8 + r];
    const double* P = cov + t * 64;
    const double2* row = reinterpret_cast<const double2*>(P + r * 8);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      double2 v = row[j];
      c[2 * j] = v.x;
      c[2 * j + 1] = v.y;
    }
    // every lane reads the 4x4 block H P H^T itself (128 B, L1-resident after the row loads)
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const double2* pr = reinterpret_cast<const double2*>(P + a * 8);
      double2 v0 = pr[0], v1 = pr[1];
      S[a][0] = v0.x; S[a][1] = v0.y; S[a][2] = v1.x; S[a][3] = v1.y;
    }
    const double2* zp = reinterpret_cast<const double2*>(meas + zi * 4);
    double2 z0 = zp[0], z1 = zp[1];
    z[0] = z0.x; z[1] = z0.y; z[2] = z1.x; z[3] = z1.y;
  }
  const double w = btd_shfl(m, base + 2);
  const double h = btd_shfl(m, base + 3);
  // innovation covariance noise, demo:253-258 (w,h of the predicted mean)
  double nw, nh;
  if (active && f32_noise) {
    const float sw = __fmul_rn((float)BT_STD_POS, (float)w), sh = __fmul_rn((float)BT_STD_POS, (float)h);
    nw = (double)__fmul_rn(sw, sw);
    nh = (double)__fmul_rn(sh, sh);
  } else {
    const double sw = BT_STD_POS * w, sh = BT_STD_POS * h;
    nw = sw * sw;
    nh = sh * sh;
  }
  S[0][0] += nw; S[1][1] += nh; S[2][2] += nw; S[3][3] += nh;

  // lower Cholesky factor of S (scipy.linalg.cho_factor(lower=True) reads the lower triangle)
  const double l00 = sqrt(S[0][0]);
  const double l10 = S[1][0] / l00, l20 = S[2][0] / l00, l30 = S[3][0] / l00;
  const double l11 = sqrt(S[1][1] - l10 * l10);
  const double l21 = (S[2][1] - l20 * l10) / l11, l31 = (S[3][1] - l30 * l10) / l11;
  const double l22 = sqrt(S[2][2] - l20 * l20 - l21 * l21);
  const double l32 = (S[3][2] - l30 * l20 - l31 * l21) / l22;
  const double l33 = sqrt(S[3][3] - l30 * l30 - l31 * l31 - l32 * l32);
  // K[r,:] = S^-1 (P H^T)[r,:]  (cho_solve, demo:328-330): forward then backward substitution
  const double y0 = c[0] / l00;
  const double y1 = (c[1] - l10 * y0) / l11;
  const double y2 = (c[2] - l20 * y0 - l21 * y1) / l22;
  const double y3 = (c[3] - l30 * y0 - l31 * y1 - l32 * y2) / l33;
  double kr[4];
  kr[3] = y3 / l33;
  kr[2] = (y2 - l32 * kr[3]) / l22;
  kr[1] = (y1 - l21 * kr[2] - l31 * kr[3]) / l11;
  kr[0] = (y0 - l10 * kr[1] - l20 * kr[2] - l30 * kr[3]) / l00;

  // mean' = mean + innovation . K^T (demo:331-333)
  double acc = 0.0;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const double innov = z[a] - btd_shfl(m, base + a);
    acc += innov * kr[a];
  }
  const double m_new = m + acc;
  // P' = P - K S K^T (demo:334-335), T = K S in-lane, K rows of the other lanes by shuffle
  double tr[4];
#pragma unroll
  for (int b = 0; b < 4; ++b) {
    double s = 0.0;
#pragma unroll
    for (int a = 0; a < 4; ++a) s += kr[a] * S[a][b];
    tr[b] = s;
  }
#pragma unroll
  for (int cc = 0; cc < 8; ++cc) {
    double s = 0.0;
#pragma unroll
    for (int b = 0; b < 4; ++b) s += tr[b] * btd_shfl(kr[b], base + cc);
    c[cc] = c[cc] - s;
  }
  if (active) {
    mean[t * 8 + r] = m_new;
    double2* row = reinterpret_cast<double2*>(cov + t * 64 + r * 8);
#pragma unroll
    for (int j = 0; j < 4; ++j) row[j] = make_double2(c[2 * j], c[2 * j + 1]);
  }
  btd_store_tlbr(m_new, lane, r, active, t, tlbr, tlbr_f32);
  if (res_tlbr) btd_store_tlbr(m_new, lane, r, active, res_t, res_tlbr, nullptr);
}

// KalmanFilter.initiate (demo:166-197) with NumPy>=2 float32 rounding of the float32 measurement path:
// z = xywh32[s] -> track t.
__device__ __forceinline__ void btd_initiate(const float* __restrict__ xywh, size_t s, double* __restrict__ mean,
                                             double* __restrict__ cov, double* __restrict__ tlbr,
                                             float* __restrict__ tlbr_f32, size_t t, bool active,
                                             uint8_t* __restrict__ slot_f32, int lane) {
  const int r = lane & 7;
  float z[4] = {0.f, 0.f, 1.f, 1.f};
  if (active) {
    const float4 v = *reinterpret_cast<const float4*>(xywh + s * 4);
    z[0] = v.x; z[1] = v.y; z[2] = v.z; z[3] = v.w;
  }
  const float wh = (r & 1) ? z[3] : z[2];
  // 2*std_pos = 0.1, 10*std_vel = 0.0625 as Python floats, weakly promoted to float32 (NEP 50)
  const float wt = (r < 4) ? (float)(2 * BT_STD_POS) : (float)(10 * BT_STD_VEL);
  const float sd = __fmul_rn(wt, wh);
  const double var = (double)__fmul_rn(sd, sd);
  const double m = (r < 4) ? (double)z[r] : 0.0;
  if (active) {
    mean[t * 8 + r] = m;
    double2* row = reinterpret_cast<double2*>(cov + t * 64 + r * 8);
#pragma unroll
    for (int j = 0; j < 4; ++j)
      row[j] = make_double2((2 * j == r) ? var : 0.0, (2 * j + 1 == r) ? var : 0.0);
    if (slot_f32 && r == 0) slot_f32[t] = 1;   // float32 state until the first predict / update (NumPy >= 2)
  }
  btd_store_tlbr(m, lane, r, active, t, tlbr, tlbr_f32);
}
