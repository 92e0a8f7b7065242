// Stand-alone ReID feature kernels behind bt_embedding_distance / bt_fused_cost (fp32 -> fp16 operand
// staging, optional row normalisation) and bt_feature_ema (STrack.update_body_features, demo:492-502;
// demo = /root/reference/demo_bottrack_onnx_tflite.py).  One CTA per feature row.  The tracker's own
// per-frame feature work is fused into frame_kernels.cu.
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
// 2 = adopt an already-normalised feature (smooth = curr = feat).  Stand-alone entry point
// (bt_feature_ema); the tracker's own EMA runs inside frame_post_kernel (frame_kernels.cu).
__global__ void __launch_bounds__(kThreads)
feature_ema_kernel(float* __restrict__ smooth, float* __restrict__ curr, const float* __restrict__ feat,
                   const int32_t* __restrict__ track_idx, const int32_t* __restrict__ feat_idx,
                   const uint8_t* __restrict__ first, int d, float alpha, float one_minus) {
  __shared__ float red[kThreads / 32];
  const int i = blockIdx.x;
  const size_t t = track_idx ? track_idx[i] : i;
  const size_t f = feat_idx ? feat_idx[i] : i;
  const int mode = first ? first[i] : 0;
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

int32_t btk_feature_ema(bt_ctx* ctx, float* smooth, float* curr, const float* feat,
                        const int32_t* track_idx, const int32_t* feat_idx, const uint8_t* first,
                        int32_t k, int32_t d, float alpha, float one_minus_alpha) {
  if (k <= 0) return BT_OK;
  feature_ema_kernel<<<k, kThreads, 0, ctx->stream>>>(smooth, curr, feat, track_idx, feat_idx, first, d, alpha,
                                                      one_minus_alpha);
  BT_LAUNCHED(ctx);
  return BT_OK;
}
