// Batched 8-state Kalman filter kernels (fp64 state), replacing KalmanFilter of the reference
// (demo:118-336; demo = /root/reference/demo_bottrack_onnx_tflite.py).
//
// Layout: AoS per track, mean[t][8] and cov[t][8][8] float64, both contiguous per track, so
// the 8 lanes that own one track (lane r <-> state row r) read one 64 B mean line and eight
// 64 B covariance rows = one contiguous 512 B span: every 32 B sector moved is used.
// 4 tracks per warp; cross-row terms move by warp shuffle; nothing is staged in shared memory
// (there is no reuse across tracks).  HBM-bound by construction: algorithmic bytes
// 2*72*8 = 1152 B/track for predict, (72+4+72)*8 = 1184 B/match for update.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;  // 32 tracks per CTA

__device__ __forceinline__ double shfl_d(double v, int src_lane) {
  return __shfl_sync(0xffffffffu, v, src_lane);
}

// Writes the cached tlbr of a track from its (new) mean held one component per lane.
// STrack.tlwh / .tlbr, demo:624-648: x1 = cx - w/2, x2 = w + x1 (same op order, fp64).
__device__ __forceinline__ void store_tlbr(double m, int lane, int r, bool active, int t,
                                           double* __restrict__ tlbr, float* __restrict__ tlbr_f32) {
  const int base = lane & ~7;
  const double c = shfl_d(m, base + (r & 1));
  const double wh = shfl_d(m, base + 2 + (r & 1));
  const double lo = c - wh / 2;
  const double val = (r < 2) ? lo : (wh + lo);
  if (active && r < 4) {
    if (tlbr) tlbr[(size_t)t * 4 + r] = val;
    // conservative fp32 interval for the fast no-overlap test of the association epilogue
    if (tlbr_f32) tlbr_f32[(size_t)t * 4 + r] = (r < 2) ? __double2float_rd(val) : __double2float_ru(val);
  }
}

// KalmanFilter.multi_predict (demo:265-302) + STrack.multi_predict's velocity reset (demo:529-532).
__global__ void __launch_bounds__(kThreads)
kalman_predict_kernel(double* __restrict__ mean, double* __restrict__ cov, double* __restrict__ tlbr,
                      float* __restrict__ tlbr_f32, const int32_t* __restrict__ state,
                      const int32_t* __restrict__ idx, int n, int noise_f32, uint8_t* __restrict__ slot_f32) {
  bt_grid_launch_dependents();   // frame step: the association kernel behind this one sets itself up meanwhile
  const int gid = blockIdx.x * kThreads + threadIdx.x;
  const int g = gid >> 3;
  const int r = threadIdx.x & 7;
  const int lane = threadIdx.x & 31;
  const int base = lane & ~7;
  const bool active = g < n;
  const int t = active ? (idx ? idx[g] : g) : 0;

  double m = 0.0;
  double c[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) c[j] = 0.0;
  if (active) {
    m = mean[(size_t)t * 8 + r];
    const double2* row = reinterpret_cast<const double2*>(cov + (size_t)t * 64 + r * 8);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      double2 v = row[j];
      c[2 * j] = v.x;
      c[2 * j + 1] = v.y;
    }
    if (state && state[g] != BT_STATE_TRACKED && r >= 6) m = 0.0;  // demo:529-532
  }
  // process noise from the PRE-predict w,h (demo:281-291)
  const double w = shfl_d(m, base + 2);
  const double h = shfl_d(m, base + 3);
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
  const double m_hi = shfl_d(m, base + ((r + 4) & 7));
  if (r < 4) m = m + m_hi;
  // P <- F P F^T: rows 0..3 += rows 4..7, then cols 0..3 += cols 4..7 (same association order
  // as the two np.dot calls of demo:299-300)
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const double o = shfl_d(c[j], base + ((r + 4) & 7));
    if (r < 4) c[j] = c[j] + o;
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) c[j] = c[j] + c[j + 4];
#pragma unroll
  for (int j = 0; j < 8; ++j)
    if (j == r) c[j] += q;

  if (active) {
    mean[(size_t)t * 8 + r] = m;
    double2* row = reinterpret_cast<double2*>(cov + (size_t)t * 64 + r * 8);
#pragma unroll
    for (int j = 0; j < 4; ++j) row[j] = make_double2(c[2 * j], c[2 * j + 1]);
    if (slot_f32 && r == 0) slot_f32[t] = 0;   // the state is float64 from now on
  }
  store_tlbr(m, lane, r, active, t, tlbr, tlbr_f32);
}

// KalmanFilter.update (demo:304-336) for k (track, measurement) pairs; project = demo:236-263.
__global__ void __launch_bounds__(kThreads)
kalman_update_kernel(double* __restrict__ mean, double* __restrict__ cov, double* __restrict__ tlbr,
                     float* __restrict__ tlbr_f32, const double* __restrict__ meas,
                     const int32_t* __restrict__ track_idx, const int32_t* __restrict__ meas_idx,
                     const uint8_t* __restrict__ noise_f32, int k, const int32_t* __restrict__ x1,
                     const int32_t* __restrict__ x2, const int32_t* __restrict__ x3,
                     uint8_t* __restrict__ slot_f32, double* __restrict__ res_tlbr) {
  bt_grid_launch_dependents();   // frame step: the duplicate test is queued behind this kernel and waits for its boxes
  const int gid = blockIdx.x * kThreads + threadIdx.x;
  const int g = gid >> 3;
  const int r = threadIdx.x & 7;
  const int lane = threadIdx.x & 31;
  const int base = lane & ~7;
  bool active = g < k;
  int t = active ? (track_idx ? track_idx[g] : g) : 0;
  int zi = active ? (meas_idx ? meas_idx[g] : g) : 0;
  if (x1) {
    // tracker mode: group g is track slot g, its measurement is whatever detection one of the three
    // association stages assigned to it (a slot is matched in at most one stage)
    t = g;
    zi = -1;
    if (active) {
      zi = x1[g];
      if (zi < 0) zi = x2[g];
      if (zi < 0) zi = x3[g];
    }
    if (res_tlbr && g < k && zi < 0 && r < 4) res_tlbr[(size_t)g * 4 + r] = tlbr[(size_t)g * 4 + r];  // unchanged box
    active = active && zi >= 0;
    if (!active) zi = 0;
  }

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
    m = mean[(size_t)t * 8 + r];
    const double* P = cov + (size_t)t * 64;
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
    const double2* zp = reinterpret_cast<const double2*>(meas + (size_t)zi * 4);
    double2 z0 = zp[0], z1 = zp[1];
    z[0] = z0.x; z[1] = z0.y; z[2] = z1.x; z[3] = z1.y;
  }
  const double w = shfl_d(m, base + 2);
  const double h = shfl_d(m, base + 3);
  // innovation covariance noise, demo:253-258 (w,h of the predicted mean)
  double nw, nh;
  const bool f32 = active && (slot_f32 ? slot_f32[t] != 0 : (noise_f32 && noise_f32[g]));
  if (active && slot_f32 && r == 0) slot_f32[t] = 0;
  if (f32) {
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
    const double innov = z[a] - shfl_d(m, base + a);
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
    for (int b = 0; b < 4; ++b) s += tr[b] * shfl_d(kr[b], base + cc);
    c[cc] = c[cc] - s;
  }
  if (active) {
    mean[(size_t)t * 8 + r] = m_new;
    double2* row = reinterpret_cast<double2*>(cov + (size_t)t * 64 + r * 8);
#pragma unroll
    for (int j = 0; j < 4; ++j) row[j] = make_double2(c[2 * j], c[2 * j + 1]);
  }
  store_tlbr(m_new, lane, r, active, t, tlbr, tlbr_f32);
  if (res_tlbr) store_tlbr(m_new, lane, r, active, t, res_tlbr, nullptr);
}

// KalmanFilter.initiate (demo:166-197) with NumPy>=2 float32 rounding of the float32 measurement path.
__global__ void __launch_bounds__(kThreads)
kalman_initiate_kernel(const float* __restrict__ xywh, const int32_t* __restrict__ src_idx,
                       double* __restrict__ mean, double* __restrict__ cov, double* __restrict__ tlbr,
                       float* __restrict__ tlbr_f32, const int32_t* __restrict__ dst_idx, int k,
                       uint8_t* __restrict__ slot_f32) {
  const int gid = blockIdx.x * kThreads + threadIdx.x;
  const int g = gid >> 3;
  const int r = threadIdx.x & 7;
  const int lane = threadIdx.x & 31;
  const bool active = g < k;
  const int s = active ? (src_idx ? src_idx[g] : g) : 0;
  const int t = active ? (dst_idx ? dst_idx[g] : g) : 0;
  float z[4] = {0.f, 0.f, 1.f, 1.f};
  if (active) {
    const float4 v = *reinterpret_cast<const float4*>(xywh + (size_t)s * 4);
    z[0] = v.x; z[1] = v.y; z[2] = v.z; z[3] = v.w;
  }
  const float wh = (r & 1) ? z[3] : z[2];
  // 2*std_pos = 0.1, 10*std_vel = 0.0625 as Python floats, weakly promoted to float32 (NEP 50)
  const float wt = (r < 4) ? (float)(2 * BT_STD_POS) : (float)(10 * BT_STD_VEL);
  const float sd = __fmul_rn(wt, wh);
  const double var = (double)__fmul_rn(sd, sd);
  const double m = (r < 4) ? (double)z[r] : 0.0;
  if (active) {
    mean[(size_t)t * 8 + r] = m;
    double2* row = reinterpret_cast<double2*>(cov + (size_t)t * 64 + r * 8);
#pragma unroll
    for (int j = 0; j < 4; ++j)
      row[j] = make_double2((2 * j == r) ? var : 0.0, (2 * j + 1 == r) ? var : 0.0);
    if (slot_f32 && r == 0) slot_f32[t] = 1;   // float32 state until the first predict / update (NumPy >= 2)
  }
  store_tlbr(m, lane, r, active, t, tlbr, tlbr_f32);
}

__global__ void kalman_project_kernel(const double* __restrict__ mean, const double* __restrict__ cov,
                                      double* __restrict__ pmean, double* __restrict__ pcov, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * 16) return;
  const int t = i >> 4, a = (i >> 2) & 3, b = i & 3;
  double v = cov[(size_t)t * 64 + a * 8 + b];
  if (a == b) {
    const double s = BT_STD_POS * mean[(size_t)t * 8 + 2 + (a & 1)];
    v += s * s;
    pmean[(size_t)t * 4 + a] = mean[(size_t)t * 8 + a];
  }
  pcov[(size_t)t * 16 + a * 4 + b] = v;
}

inline int blocks_for_tracks(int n) { return (n * 8 + kThreads - 1) / kThreads; }

}  // namespace

int32_t btk_kalman_initiate(bt_ctx* ctx, const float* xywh, const int32_t* src_idx, double* mean,
                            double* cov, double* tlbr, float* tlbr_f32, const int32_t* dst_idx, int32_t k,
                            uint8_t* slot_f32) {
  if (k <= 0) return BT_OK;
  kalman_initiate_kernel<<<blocks_for_tracks(k), kThreads, 0, ctx->stream>>>(xywh, src_idx, mean, cov, tlbr,
                                                                              tlbr_f32, dst_idx, k, slot_f32);
  BT_LAUNCHED(ctx);
  return BT_OK;
}

int32_t btk_kalman_predict(bt_ctx* ctx, double* mean, double* cov, double* tlbr, float* tlbr_f32,
                           const int32_t* state, const int32_t* idx, int32_t n, int32_t noise_f32,
                           uint8_t* slot_f32) {
  if (n <= 0) return BT_OK;
  kalman_predict_kernel<<<blocks_for_tracks(n), kThreads, 0, ctx->stream>>>(mean, cov, tlbr, tlbr_f32, state,
                                                                             idx, n, noise_f32, slot_f32);
  BT_LAUNCHED(ctx);
  return BT_OK;
}

int32_t btk_kalman_update(bt_ctx* ctx, double* mean, double* cov, double* tlbr, float* tlbr_f32,
                          const double* meas, const int32_t* track_idx, const int32_t* meas_idx,
                          const uint8_t* noise_f32, int32_t k) {
  if (k <= 0) return BT_OK;
  kalman_update_kernel<<<blocks_for_tracks(k), kThreads, 0, ctx->stream>>>(
      mean, cov, tlbr, tlbr_f32, meas, track_idx, meas_idx, noise_f32, k, nullptr, nullptr, nullptr, nullptr, nullptr);
  BT_LAUNCHED(ctx);
  return BT_OK;
}

int32_t btk_kalman_update_x(bt_ctx* ctx, double* mean, double* cov, double* tlbr, float* tlbr_f32,
                            const double* meas, const int32_t* x1, const int32_t* x2, const int32_t* x3,
                            uint8_t* slot_f32, int32_t n_slots, double* res_tlbr) {
  if (n_slots <= 0) return BT_OK;
  kalman_update_kernel<<<blocks_for_tracks(n_slots), kThreads, 0, ctx->stream>>>(
      mean, cov, tlbr, tlbr_f32, meas, nullptr, nullptr, nullptr, n_slots, x1, x2, x3, slot_f32, res_tlbr);
  BT_LAUNCHED(ctx);
  return BT_OK;
}

int32_t btk_kalman_project(bt_ctx* ctx, const double* mean, const double* cov, double* pmean,
                           double* pcov, int32_t n) {
  if (n <= 0) return BT_OK;
  kalman_project_kernel<<<(n * 16 + 255) / 256, 256, 0, ctx->stream>>>(mean, cov, pmean, pcov, n);
  BT_LAUNCHED(ctx);
  return BT_OK;
}
