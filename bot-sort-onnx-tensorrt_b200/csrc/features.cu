// ReID feature kernels: per-detection normalisation + fp16 staging for the tensor-core
// similarity GEMM, and the batched feature EMA of STrack.update_body_features
// (demo:492-502; demo = /root/reference/demo_bottrack_onnx_tflite.py).
// One CTA per feature row (2048 floats = 8 KB: 256 threads x 2 float4), block reduction for
// the L2 norm; HBM-bound: prep 4d in + (4d + 2d) out per detection, EMA 3*4*d B per match.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ float block_sum(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kThreads / 32; ++i) s += red[i];
  return s;
}

// Detection STrack construction normalises its feature row in place (demo:497-502 on the first
// call): out = feat / ||feat||_2 (float32).  The fp16 copy is the B operand of the similarity GEMM.
__global__ void __launch_bounds__(kThreads)
feature_prep_kernel(const float* __restrict__ feat, int d, float* __restrict__ out_f32,
                    __half* __restrict__ out_f16, int normalise) {
  __shared__ float red[kThreads / 32];
  const size_t row = blockIdx.x;
  const float* src = feat + row * d;
  constexpr int kHold = 4;               // float4 per thread kept in registers: rows up to 4096 floats are read once
  const bool vec = (d & 3) == 0;
  const bool held = vec && d <= kThreads * 4 * kHold;
  float4 v[kHold];
  float ss = 0.f;
  if (held) {
#pragma unroll
    for (int t = 0; t < kHold; ++t) {
      const int i = (threadIdx.x + t * kThreads) * 4;
      v[t] = (i < d) ? __ldg(reinterpret_cast<const float4*>(src + i)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int t = 0; t < kHold; ++t) ss += v[t].x * v[t].x + v[t].y * v[t].y + v[t].z * v[t].z + v[t].w * v[t].w;
  } else if (vec) {
    for (int i = threadIdx.x * 4; i < d; i += kThreads * 4) {
      const float4 w = *reinterpret_cast<const float4*>(src + i);
      ss += w.x * w.x + w.y * w.y + w.z * w.z + w.w * w.w;
    }
  } else {
    for (int i = threadIdx.x; i < d; i += kThreads) ss += src[i] * src[i];
  }
  float norm = 1.f;
  if (normalise) norm = sqrtf(block_sum(ss, red));
  auto put = [&](int i, float4 w) {
    if (normalise) { w.x /= norm; w.y /= norm; w.z /= norm; w.w /= norm; }
    if (out_f32) *reinterpret_cast<float4*>(out_f32 + row * d + i) = w;
    if (out_f16) {
      __half2 h0 = __floats2half2_rn(w.x, w.y), h1 = __floats2half2_rn(w.z, w.w);
      uint2 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&h0);
      pk.y = *reinterpret_cast<uint32_t*>(&h1);
      *reinterpret_cast<uint2*>(out_f16 + row * d + i) = pk;
    }
  };
  if (held) {
#pragma unroll
    for (int t = 0; t < kHold; ++t) {
      const int i = (threadIdx.x + t * kThreads) * 4;
      if (i < d) put(i, v[t]);
    }
  } else if (vec) {
    for (int i = threadIdx.x * 4; i < d; i += kThreads * 4) put(i, *reinterpret_cast<const float4*>(src + i));
  } else {
    for (int i = threadIdx.x; i < d; i += kThreads) {
      float w = src[i];
      if (normalise) w /= norm;
      if (out_f32) out_f32[row * d + i] = w;
      if (out_f16) out_f16[row * d + i] = __float2half_rn(w);
    }
  }
  // Launched as a programmatic dependent of det_prep (frame step): nothing above reads what det_prep
  // writes, so the two run side by side; completion stays ordered behind it for the association kernel.
  bt_grid_dependency_wait();
}

// mode (first[i]): 0 = EMA (demo:499-502), 1 = first call on a raw feature: smooth = feat/||feat||,
// 2 = adopt an already-normalised detection feature (birth: smooth = curr = feat).
// Optionally also refreshes the fp16 bank row used as the GEMM A operand.
__global__ void __launch_bounds__(kThreads)
feature_ema_kernel(float* __restrict__ smooth, float* __restrict__ curr, const float* __restrict__ feat,
                   __half* __restrict__ bank16, const __half* __restrict__ det16,
                   const int32_t* __restrict__ track_idx, const int32_t* __restrict__ feat_idx,
                   const uint8_t* __restrict__ first, int d, float alpha, const int32_t* __restrict__ x1,
                   const int32_t* __restrict__ x2, const int32_t* __restrict__ x3) {
  __shared__ float red[kThreads / 32];
  const int i = blockIdx.x;
  size_t t = track_idx ? track_idx[i] : i;
  size_t f = feat_idx ? feat_idx[i] : i;
  if (x1) {   // tracker mode: slot i, detection assigned by one of the three stages (block-uniform exit)
    int z = x1[i];
    if (z < 0) z = x2[i];
    if (z < 0) z = x3[i];
    if (z < 0) return;
    t = i;
    f = z;
  }
  const int mode = first ? first[i] : 0;
  if (bank16 && det16) {
    if ((d & 3) == 0) {
      const uint2* s = reinterpret_cast<const uint2*>(det16 + f * d);
      uint2* o = reinterpret_cast<uint2*>(bank16 + t * d);
      for (int j = threadIdx.x; j < d / 4; j += kThreads) o[j] = s[j];
    } else {
      for (int j = threadIdx.x; j < d; j += kThreads) bank16[t * d + j] = det16[f * d + j];
    }
  }
  if (!smooth && !curr) return;
  const float one_minus = (float)(1.0 - (double)alpha);
  constexpr int kHold = 4;               // float4 per thread kept in registers: rows up to 4096 floats are read once
  if ((d & 3) == 0 && d <= kThreads * 4 * kHold) {
    float4 xv[kHold], sv[kHold];
#pragma unroll
    for (int k = 0; k < kHold; ++k) {
      const int j = (threadIdx.x + k * kThreads) * 4;
      xv[k] = (j < d) ? __ldg(reinterpret_cast<const float4*>(feat + f * d + j)) : make_float4(0.f, 0.f, 0.f, 0.f);
      sv[k] = xv[k];
      if (mode == 0 && j < d) {
        const float4 o = *reinterpret_cast<const float4*>(smooth + t * d + j);
        sv[k].x = __fadd_rn(__fmul_rn(alpha, o.x), __fmul_rn(one_minus, xv[k].x));
        sv[k].y = __fadd_rn(__fmul_rn(alpha, o.y), __fmul_rn(one_minus, xv[k].y));
        sv[k].z = __fadd_rn(__fmul_rn(alpha, o.z), __fmul_rn(one_minus, xv[k].z));
        sv[k].w = __fadd_rn(__fmul_rn(alpha, o.w), __fmul_rn(one_minus, xv[k].w));
      }
    }
    float ss = 0.f;
#pragma unroll
    for (int k = 0; k < kHold; ++k) ss += sv[k].x * sv[k].x + sv[k].y * sv[k].y + sv[k].z * sv[k].z + sv[k].w * sv[k].w;
    float norm = 1.f;
    if (mode != 2) norm = sqrtf(block_sum(ss, red));
#pragma unroll
    for (int k = 0; k < kHold; ++k) {
      const int j = (threadIdx.x + k * kThreads) * 4;
      if (j >= d) continue;
      float4 o = sv[k];
      if (mode != 2) { o.x /= norm; o.y /= norm; o.z /= norm; o.w /= norm; }
      if (smooth) *reinterpret_cast<float4*>(smooth + t * d + j) = o;
      if (curr) *reinterpret_cast<float4*>(curr + t * d + j) = (mode == 1) ? o : xv[k];
    }
    return;
  }
  // generic sizes: two passes, the new smooth is recomputed in the second
  float ss = 0.f;
  for (int j = threadIdx.x; j < d; j += kThreads) {
    const float x = feat[f * d + j];
    float s;
    if (mode == 0) s = __fadd_rn(__fmul_rn(alpha, smooth[t * d + j]), __fmul_rn(one_minus, x));
    else s = x;
    ss += s * s;
  }
  float norm = 1.f;
  if (mode != 2) norm = sqrtf(block_sum(ss, red));
  for (int j = threadIdx.x; j < d; j += kThreads) {
    const float x = feat[f * d + j];
    float s;
    if (mode == 0) s = __fadd_rn(__fmul_rn(alpha, smooth[t * d + j]), __fmul_rn(one_minus, x));
    else s = x;
    if (mode != 2) s = s / norm;
    if (smooth) smooth[t * d + j] = s;
    if (curr) curr[t * d + j] = (mode == 1) ? s : x;
  }
}

}  // namespace

int32_t btk_feature_prep(bt_ctx* ctx, const float* feat, int32_t m, int32_t d, float* out_f32,
                         __half* out_f16, int32_t normalise, int32_t dependent) {
  if (m <= 0) return BT_OK;
  BT_CUDA(bt_launch(ctx, dependent != 0, feature_prep_kernel, dim3(m), dim3(kThreads), 0, feat, d, out_f32, out_f16, normalise));
  BT_LAUNCHED(ctx);
  return BT_OK;
}

int32_t btk_feature_ema16(bt_ctx* ctx, float* smooth, float* curr, const float* feat, __half* bank16,
                          const __half* det16, const int32_t* track_idx, const int32_t* feat_idx,
                          const uint8_t* first, int32_t k, int32_t d, float alpha) {
  if (k <= 0) return BT_OK;
  feature_ema_kernel<<<k, kThreads, 0, ctx->stream>>>(smooth, curr, feat, bank16, det16, track_idx,
                                                      feat_idx, first, d, alpha, nullptr, nullptr, nullptr);
  BT_LAUNCHED(ctx);
  return BT_OK;
}

int32_t btk_feature_ema_x(bt_ctx* ctx, float* smooth, float* curr, const float* feat, __half* bank16,
                          const __half* det16, const int32_t* x1, const int32_t* x2, const int32_t* x3,
                          int32_t n_slots, int32_t d, float alpha) {
  if (n_slots <= 0) return BT_OK;
  feature_ema_kernel<<<n_slots, kThreads, 0, ctx->stream>>>(smooth, curr, feat, bank16, det16, nullptr, nullptr,
                                                            nullptr, d, alpha, x1, x2, x3);
  BT_LAUNCHED(ctx);
  return BT_OK;
}

int32_t btk_feature_ema(bt_ctx* ctx, float* smooth, float* curr, const float* feat,
                        const int32_t* track_idx, const int32_t* feat_idx, const uint8_t* first,
                        int32_t k, int32_t d, float alpha) {
  return btk_feature_ema16(ctx, smooth, curr, feat, nullptr, nullptr, track_idx, feat_idx, first, k, d, alpha);
}
