// Batched 8-state Kalman filter kernels (fp64 state), replacing KalmanFilter of the reference
// (demo:118-336; demo = /root/reference/demo_bottrack_onnx_tflite.py).
//
// Layout: AoS per track, mean[t][8] and cov[t][8][8] float64, both contiguous per track, so
// the 8 lanes that own one track (lane r <-> state row r) read one 64 B mean line and eight
// 64 B covariance rows = one contiguous 512 B span: every 32 B sector moved is used.
// 4 tracks per warp; cross-row terms move by warp shuffle; nothing is staged in shared memory
// (there is no reuse across tracks).  HBM-bound by construction: algorithmic bytes
// 2*72*8 = 1152 B/track for predict, (72+4+72)*8 = 1184 B/match for update.
#include "kalman_dev.cuh"

namespace {

constexpr int kThreads = 256;  // 32 tracks per CTA

// KalmanFilter.multi_predict (demo:265-302) + STrack.multi_predict's velocity reset (demo:529-532).
__global__ void __launch_bounds__(kThreads)
kalman_predict_kernel(double* __restrict__ mean, double* __restrict__ cov, double* __restrict__ tlbr,
                      float* __restrict__ tlbr_f32, const int32_t* __restrict__ state,
                      const int32_t* __restrict__ idx, int n, int noise_f32, uint8_t* __restrict__ slot_f32) {
  const int g = (blockIdx.x * kThreads + threadIdx.x) >> 3;
  const bool active = g < n;
  const size_t t = active ? (size_t)(idx ? idx[g] : g) : 0;
  const bool reset = active && state && state[g] != BT_STATE_TRACKED;
  btd_predict(mean, cov, tlbr, tlbr_f32, t, active, reset, noise_f32, slot_f32, threadIdx.x & 31);
}

// KalmanFilter.update (demo:304-336) for k (track, measurement) pairs; project = demo:236-263.
__global__ void __launch_bounds__(kThreads)
kalman_update_kernel(double* __restrict__ mean, double* __restrict__ cov, double* __restrict__ tlbr,
                     float* __restrict__ tlbr_f32, const double* __restrict__ meas,
                     const int32_t* __restrict__ track_idx, const int32_t* __restrict__ meas_idx,
                     const uint8_t* __restrict__ noise_f32, int k) {
  const int g = (blockIdx.x * kThreads + threadIdx.x) >> 3;
  const bool active = g < k;
  const size_t t = active ? (size_t)(track_idx ? track_idx[g] : g) : 0;
  const size_t zi = active ? (size_t)(meas_idx ? meas_idx[g] : g) : 0;
  const bool f32 = active && noise_f32 && noise_f32[g];
  btd_update(mean, cov, tlbr, tlbr_f32, meas, t, zi, active, f32, nullptr, 0, threadIdx.x & 31);
}

// KalmanFilter.initiate (demo:166-197)
__global__ void __launch_bounds__(kThreads)
kalman_initiate_kernel(const float* __restrict__ xywh, const int32_t* __restrict__ src_idx,
                       double* __restrict__ mean, double* __restrict__ cov, double* __restrict__ tlbr,
                       float* __restrict__ tlbr_f32, const int32_t* __restrict__ dst_idx, int k,
                       uint8_t* __restrict__ slot_f32) {
  const int g = (blockIdx.x * kThreads + threadIdx.x) >> 3;
  const bool active = g < k;
  const size_t s = active ? (size_t)(src_idx ? src_idx[g] : g) : 0;
  const size_t t = active ? (size_t)(dst_idx ? dst_idx[g] : g) : 0;
  btd_initiate(xywh, s, mean, cov, tlbr, tlbr_f32, t, active, slot_f32, threadIdx.x & 31);
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
      mean, cov, tlbr, tlbr_f32, meas, track_idx, meas_idx, noise_f32, k);
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
