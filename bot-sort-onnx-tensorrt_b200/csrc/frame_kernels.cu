// The per-frame launches of the tracker around the association and LAP kernels, each serving every
// video stream of the batch at once (blockIdx.y = batch entry; SURVEY 8(e): "streams are a leading batch
// dimension of every kernel"):
//   frame_cast  (fp32 ingest only) float32 feature rows -> raw fp16 rows (GEMM B operand) + L2 norms
//   frame_prep  detection prep (demo:1493-1532 boxes / score classes) + batched Kalman predict of every
//               pool (demo:524-536, demo:265-302) + optional detection feature norms: ONE launch
//   frame_post  Kalman update (demo:304-336) of every slot matched by one of the three association stages
//   frame_ema   feature EMA (demo:492-502) of the same slots, on the ctx's side stream beside it
//   frame_dup   remove_duplicate_stracks' IoU test (demo:1665-1668) over all live slots, sparse output
//   births      Kalman initiate (demo:166-197) + feature adoption of the frame's new tracks
// demo = /root/reference/demo_bottrack_onnx_tflite.py
#include "kalman_dev.cuh"

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ float block_sum256(float v, float* red) {
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

// ---- control block upload ----------------------------------------------------------------------------
// The frame's packed control block (launch description + pool lists, ~22 KB at 2000 tracks) is read by the GPU
// straight from the pinned host buffer: a cudaMemcpyAsync of this size is staged through the command stream by
// the driver and costs the enqueueing thread ~17 us at the very front of the frame's critical path, a launch ~4.
__global__ void __launch_bounds__(256)
ctrl_upload_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int n16) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i < n16) dst[i] = src[i];
}

// ---- fp32 ingest: cast + norm ------------------------------------------------------------------
// One CTA per detection row.  det16 = round-to-nearest fp16 of the RAW row (the reference feeds the raw
// encoder output to the first association's similarity, demo:1453-1460; rows are normalised only
// when a track adopts them, demo:497-502), det_norm = ||row||_2 in float32.
__global__ void __launch_bounds__(kThreads)
frame_cast_kernel(bt_store st, const bt_batch* __restrict__ bp) {
  const bt_batch& b = *bp;
  __shared__ float red[kThreads / 32];
  const int k = blockIdx.y, j = blockIdx.x;
  if (j >= b.m[k]) return;
  const int D = st.D;
  const size_t gd = (size_t)b.sid[k] * st.md + j;
  const size_t in_row = ((size_t)b.parity[k] * st.S * st.md + gd) * D;
  const float* src = st.det32 + in_row;
  __half* dst = st.det16 + in_row;
  float ss = 0.f;
  if ((D & 3) == 0) {
    for (int i = threadIdx.x * 4; i < D; i += kThreads * 4) {
      const float4 w = __ldg(reinterpret_cast<const float4*>(src + i));
      ss += w.x * w.x + w.y * w.y + w.z * w.z + w.w * w.w;
      __half2 h0 = __floats2half2_rn(w.x, w.y), h1 = __floats2half2_rn(w.z, w.w);
      uint2 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&h0);
      pk.y = *reinterpret_cast<uint32_t*>(&h1);
      *reinterpret_cast<uint2*>(dst + i) = pk;
    }
  } else {
    for (int i = threadIdx.x; i < D; i += kThreads) {
      const float w = src[i];
      ss += w * w;
      dst[i] = __float2half_rn(w);
    }
  }
  const float tot = block_sum256(ss, red);
  if (threadIdx.x == 0) st.det_norm[gd] = sqrtf(tot);
}

// ---- prep: detection prep | Kalman predict | detection norms ---------------------------------------
// boxes int32 tlbr -> tlbr float64, xywh float64 (Kalman measurement, demo:599 + demo:664-670),
// xywh float32 (initiate input, demo:561), the score class of the detection
// (demo:1501: high = score > 0.40; demo:1531: low = 0.1 <= score <= 0.40; float(score) against Python
// doubles), packed integer corners for the association kernel's overlap screen.
__global__ void __launch_bounds__(kThreads)
frame_prep_kernel(bt_store st, const bt_batch* __restrict__ bp, bt_frame_cfg fc, int pred_blocks,
                  int det_blocks, int norm_blocks, int pf_blocks, bt_res_layout L) {
  const bt_batch& b = *bp;
  bt_grid_launch_dependents();   // the association kernel behind this one loads its operands meanwhile
  __shared__ float red[kThreads / 32];
  const int k = blockIdx.y;
  const int sid = b.sid[k];
  int bx = blockIdx.x;
  if (bx < pred_blocks) {
    // ---- Kalman predict over the stream's pool (8 lanes per track) ----
    const int g = (bx * kThreads + threadIdx.x) >> 3;
    const int n_pool = b.n_pool[k];
    if (bx * (kThreads / 8) >= n_pool) return;     // block-uniform
    const bool active = g < n_pool;
    const int32_t* pool_idx = reinterpret_cast<const int32_t*>(st.ctrl + b.ctrl_off[k]);
    const int32_t* pool_state = pool_idx + n_pool;
    const size_t t = (size_t)sid * st.cap + (active ? pool_idx[g] : 0);
    const bool reset = active && pool_state[g] != BT_STATE_TRACKED;
    btd_predict(st.mean, st.cov, st.tlbr, st.tlbr_f32, t, active, reset, b.noise_f32[k], st.slot_f32,
                threadIdx.x & 31);
    return;
  }
  bx -= pred_blocks;
  const int m = b.m[k];
  const size_t gd0 = (size_t)sid * st.md;
  const size_t in0 = (size_t)b.parity[k] * st.S * st.md + gd0;
  if (bx < det_blocks) {
    const int j = bx * kThreads + threadIdx.x;
    if (j >= m) return;
    const int4 bb = *reinterpret_cast<const int4*>(st.det_boxes + (in0 + j) * 4);
    // STrack(tlbr_to_tlwh(int box)) -> float32 tlwh (demo:465); xywh = tl + wh/2 in float32 (exact)
    const float x1 = (float)bb.x, y1 = (float)bb.y, w = (float)(bb.z - bb.x), h = (float)(bb.w - bb.y);
    const float cx = x1 + w / 2, cy = y1 + h / 2;
    double* t = st.det_tlbr + (gd0 + j) * 4;
    // detection tlbr = tlwh (float32) with wh += tl (demo:643-648)
    t[0] = x1; t[1] = y1; t[2] = (double)(w + x1); t[3] = (double)(h + y1);
    double* z = st.det_xywh + (gd0 + j) * 4;
    z[0] = cx; z[1] = cy; z[2] = w; z[3] = h;
    *reinterpret_cast<float4*>(st.det_xywh32 + (gd0 + j) * 4) = make_float4(cx, cy, w, h);
    const float s = st.det_scores[in0 + j];
    if (fc.device_inputs) {   // device-resident inputs: the host reads them back with the frame's result block
      int32_t* res = reinterpret_cast<int32_t*>(st.res + (size_t)k * L.stride);
      reinterpret_cast<float*>(res + L.o_sc)[j] = s;
      *reinterpret_cast<int4*>(res + L.o_bx + (size_t)j * 4) = bb;
      if (st.res_host) {      // direct results: straight into the pinned host copy (visible before the LAP kernel's flag)
        int32_t* hres = reinterpret_cast<int32_t*>(st.res_host + (size_t)k * L.stride);
        reinterpret_cast<float*>(hres + L.o_sc)[j] = s;
        *reinterpret_cast<int4*>(hres + L.o_bx + (size_t)j * 4) = bb;
      }
    }
    const double sd = (double)s;     // float(score), demo:1022
    st.col_kind[gd0 + j] = (sd > fc.high) ? BT_COL_HIGH : ((sd >= fc.low) ? BT_COL_LOW : BT_COL_NONE);
    st.col_pk[gd0 + j] = bt_pack16_f32(x1, y1, w + x1, h + y1, true);
    return;
  }
  bx -= det_blocks;
  const int D = st.D;
  if (bx >= norm_blocks) {
    // ---- L2 prefetch of the association kernel's operands (the bank rows in use, this frame's detection rows):
    //      the association kernel behind this launch streams them through a 4-stage TMA ring, which cannot cover
    //      DRAM latency on first touch; these blocks pull the 16 MB at full memory-level parallelism meanwhile ----
    bx -= norm_blocks;
    if (bx >= pf_blocks || ((size_t)D * sizeof(__half)) % 128 != 0) return;
    const size_t lines_a = (size_t)b.n_rows[k] * D * sizeof(__half) / 128, lines_b = (size_t)m * D * sizeof(__half) / 128;
    const char* a0 = reinterpret_cast<const char*>(st.feat16 + (size_t)sid * st.cap * D);
    const char* b0 = reinterpret_cast<const char*>(st.det16 + in0 * D);
    for (size_t i = (size_t)bx * kThreads + threadIdx.x; i < lines_a + lines_b; i += (size_t)pf_blocks * kThreads) {
      const char* ptr = i < lines_a ? a0 + i * 128 : b0 + (i - lines_a) * 128;
      asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr));
    }
    return;
  }
  // ---- detection feature norms (fp16 ingest; only for streams with unconfirmed rows) ----
  if (!b.want_norm[k] || bx >= m) return;
  const __half* src = st.det16 + (in0 + bx) * D;
  float ss = 0.f;
  if ((D & 7) == 0) {
    for (int i = threadIdx.x * 8; i < D; i += kThreads * 8) {
      const uint4 q = __ldg(reinterpret_cast<const uint4*>(src + i));
      const __half2* h = reinterpret_cast<const __half2*>(&q);
#pragma unroll
      for (int e = 0; e < 4; ++e) { const float2 f = __half22float2(h[e]); ss += f.x * f.x + f.y * f.y; }
    }
  } else {
    for (int i = threadIdx.x; i < D; i += kThreads) { const float f = __half2float(src[i]); ss += f * f; }
  }
  const float tot = block_sum256(ss, red);
  if (threadIdx.x == 0) st.det_norm[gd0 + bx] = sqrtf(tot);
}

// ---- post: Kalman update | feature EMA ----------------------------------------------------------
// Slot g of batch entry k takes the detection x1[g] / x2[g] / x3[g] (first non-negative; a slot is
// matched in at most one stage) as its measurement: STrack.update / re_activate arithmetic,
// demo:570-610.  Features (STrack.update_body_features, demo:492-502): xn = x / ||x||;
// smooth = normalise(alpha * smooth + (1 - alpha) * xn); curr = xn.  The track keeps the RAW fp16 row
// (A operand of the next frame's similarity GEMM) and its norm.
template <bool kF16>
__device__ __forceinline__ void ema_row(const bt_store& st, const bt_frame_cfg& fc, size_t gs, size_t in_row,
                                        int mode, float* red) {
  const int D = st.D;
  constexpr int kHold = 4;     // 4 floats x kHold per thread in registers: rows up to 4096 floats are read once
  const bool held = (D & 3) == 0 && D <= kThreads * 4 * kHold;
  __half* bank = st.feat16 + gs * D;
  const __half* raw16 = st.det16 + in_row * D;
  const float* raw32 = kF16 ? nullptr : st.det32 + in_row * D;
  float* smooth = (fc.keep_smooth && st.smooth32) ? st.smooth32 + gs * D : nullptr;
  float* curr = (!kF16 && st.curr32) ? st.curr32 + gs * D : nullptr;
  if (held) {
    // all loads of the row first (the detection's raw row AND the track's old smooth row: two independent DRAM
    // round trips in flight together), then the two norms
    float4 xv[kHold], ov[kHold];
    float ss = 0.f;
    const bool blend = mode == 0 && smooth != nullptr;
#pragma unroll
    for (int t = 0; t < kHold; ++t) {
      const int i = (threadIdx.x + t * kThreads) * 4;
      xv[t] = make_float4(0.f, 0.f, 0.f, 0.f);
      ov[t] = xv[t];
      if (i < D) {
        const uint2 q = __ldg(reinterpret_cast<const uint2*>(raw16 + i));
        if (blend) ov[t] = *reinterpret_cast<const float4*>(smooth + i);
        *reinterpret_cast<uint2*>(bank + i) = q;     // the track adopts the raw fp16 row
        if (kF16) {
          const __half2* h = reinterpret_cast<const __half2*>(&q);
          const float2 f0 = __half22float2(h[0]), f1 = __half22float2(h[1]);
          xv[t] = make_float4(f0.x, f0.y, f1.x, f1.y);
        } else {
          xv[t] = __ldg(reinterpret_cast<const float4*>(raw32 + i));
        }
      }
    }
#pragma unroll
    for (int t = 0; t < kHold; ++t) ss += xv[t].x * xv[t].x + xv[t].y * xv[t].y + xv[t].z * xv[t].z + xv[t].w * xv[t].w;
    const float norm = sqrtf(block_sum256(ss, red));
    if (threadIdx.x == 0) st.norm[gs] = norm;
    if (!smooth && !curr) return;
    float4 sv[kHold];
    float ss2 = 0.f;
#pragma unroll
    for (int t = 0; t < kHold; ++t) {
      if (norm > 0.f) { xv[t].x /= norm; xv[t].y /= norm; xv[t].z /= norm; xv[t].w /= norm; }
      sv[t] = xv[t];
      if (blend) {
        sv[t].x = __fadd_rn(__fmul_rn(fc.alpha, ov[t].x), __fmul_rn(fc.one_minus_alpha, xv[t].x));
        sv[t].y = __fadd_rn(__fmul_rn(fc.alpha, ov[t].y), __fmul_rn(fc.one_minus_alpha, xv[t].y));
        sv[t].z = __fadd_rn(__fmul_rn(fc.alpha, ov[t].z), __fmul_rn(fc.one_minus_alpha, xv[t].z));
        sv[t].w = __fadd_rn(__fmul_rn(fc.alpha, ov[t].w), __fmul_rn(fc.one_minus_alpha, xv[t].w));
      }
      ss2 += sv[t].x * sv[t].x + sv[t].y * sv[t].y + sv[t].z * sv[t].z + sv[t].w * sv[t].w;
    }
    float n2 = 1.f;
    if (blend) n2 = sqrtf(block_sum256(ss2, red));   // block-uniform condition
#pragma unroll
    for (int t = 0; t < kHold; ++t) {
      const int i = (threadIdx.x + t * kThreads) * 4;
      if (i >= D) continue;
      if (curr) *reinterpret_cast<float4*>(curr + i) = xv[t];
      if (smooth) {
        float4 o = sv[t];
        if (blend && n2 > 0.f) { o.x /= n2; o.y /= n2; o.z /= n2; o.w /= n2; }
        *reinterpret_cast<float4*>(smooth + i) = o;
      }
    }
    return;
  }
  // generic feature sizes: element-wise passes
  float ss = 0.f;
  for (int i = threadIdx.x; i < D; i += kThreads) {
    bank[i] = raw16[i];
    const float x = kF16 ? __half2float(raw16[i]) : raw32[i];
    ss += x * x;
  }
  const float norm = sqrtf(block_sum256(ss, red));
  if (threadIdx.x == 0) st.norm[gs] = norm;
  if (!smooth && !curr) return;
  float ss2 = 0.f;
  for (int i = threadIdx.x; i < D; i += kThreads) {
    float x = kF16 ? __half2float(raw16[i]) : raw32[i];
    if (norm > 0.f) x /= norm;
    float s = x;
    if (mode == 0 && smooth) s = __fadd_rn(__fmul_rn(fc.alpha, smooth[i]), __fmul_rn(fc.one_minus_alpha, x));
    ss2 += s * s;
  }
  float n2 = 1.f;
  if (mode == 0 && smooth) n2 = sqrtf(block_sum256(ss2, red));
  for (int i = threadIdx.x; i < D; i += kThreads) {
    float x = kF16 ? __half2float(raw16[i]) : raw32[i];
    if (norm > 0.f) x /= norm;
    if (curr) curr[i] = x;
    if (smooth) {
      float s = x;
      if (mode == 0) {
        s = __fadd_rn(__fmul_rn(fc.alpha, smooth[i]), __fmul_rn(fc.one_minus_alpha, x));
        if (n2 > 0.f) s /= n2;
      }
      smooth[i] = s;
    }
  }
}

// The same row update by HALF a CTA (128 threads, 16 floats of the row per thread): two rows per CTA keep twice
// as many rows in flight per SM at the same register budget (the update is a chain of two DRAM round trips and two
// reductions per row -- rows in flight, not bandwidth, bound it).  `half` selects named barrier 1 / 2.
__device__ __forceinline__ float half_sum128(float v, float* red, int half) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = (threadIdx.x & 127) >> 5, lane = threadIdx.x & 31;
  asm volatile("bar.sync %0, 128;" ::"r"(1 + half) : "memory");
  if (lane == 0) red[half * 4 + warp] = v;
  asm volatile("bar.sync %0, 128;" ::"r"(1 + half) : "memory");
  return red[half * 4] + red[half * 4 + 1] + red[half * 4 + 2] + red[half * 4 + 3];
}
template <bool kF16>
__device__ __forceinline__ void ema_row_half(const bt_store& st, const bt_frame_cfg& fc, size_t gs, size_t in_row,
                                             float* red, int half) {
  const int D = st.D;
  constexpr int kHold = 4;
  const int tid = threadIdx.x & 127;
  __half* bank = st.feat16 + gs * D;
  const __half* raw16 = st.det16 + in_row * D;
  const float* raw32 = kF16 ? nullptr : st.det32 + in_row * D;
  float* smooth = (fc.keep_smooth && st.smooth32) ? st.smooth32 + gs * D : nullptr;
  float* curr = (!kF16 && st.curr32) ? st.curr32 + gs * D : nullptr;
  float4 xv[kHold], ov[kHold];
  float ss = 0.f;
  const bool blend = smooth != nullptr;
#pragma unroll
  for (int t = 0; t < kHold; ++t) {
    const int i = (tid + t * 128) * 4;
    xv[t] = make_float4(0.f, 0.f, 0.f, 0.f);
    ov[t] = xv[t];
    if (i < D) {
      const uint2 q = __ldg(reinterpret_cast<const uint2*>(raw16 + i));
      if (blend) ov[t] = *reinterpret_cast<const float4*>(smooth + i);
      *reinterpret_cast<uint2*>(bank + i) = q;     // the track adopts the raw fp16 row
      if (kF16) {
        const __half2* h = reinterpret_cast<const __half2*>(&q);
        const float2 f0 = __half22float2(h[0]), f1 = __half22float2(h[1]);
        xv[t] = make_float4(f0.x, f0.y, f1.x, f1.y);
      } else {
        xv[t] = __ldg(reinterpret_cast<const float4*>(raw32 + i));
      }
    }
  }
#pragma unroll
  for (int t = 0; t < kHold; ++t) ss += xv[t].x * xv[t].x + xv[t].y * xv[t].y + xv[t].z * xv[t].z + xv[t].w * xv[t].w;
  const float norm = sqrtf(half_sum128(ss, red, half));
  if (tid == 0) st.norm[gs] = norm;
  if (!smooth && !curr) return;
  float ss2 = 0.f;
#pragma unroll
  for (int t = 0; t < kHold; ++t) {
    if (norm > 0.f) { xv[t].x /= norm; xv[t].y /= norm; xv[t].z /= norm; xv[t].w /= norm; }
    if (blend) {      // the blended row replaces the old one in place (registers)
      ov[t].x = __fadd_rn(__fmul_rn(fc.alpha, ov[t].x), __fmul_rn(fc.one_minus_alpha, xv[t].x));
      ov[t].y = __fadd_rn(__fmul_rn(fc.alpha, ov[t].y), __fmul_rn(fc.one_minus_alpha, xv[t].y));
      ov[t].z = __fadd_rn(__fmul_rn(fc.alpha, ov[t].z), __fmul_rn(fc.one_minus_alpha, xv[t].z));
      ov[t].w = __fadd_rn(__fmul_rn(fc.alpha, ov[t].w), __fmul_rn(fc.one_minus_alpha, xv[t].w));
      ss2 += ov[t].x * ov[t].x + ov[t].y * ov[t].y + ov[t].z * ov[t].z + ov[t].w * ov[t].w;
    }
  }
  float n2 = 1.f;
  if (blend) n2 = sqrtf(half_sum128(ss2, red, half));   // half-uniform condition
#pragma unroll
  for (int t = 0; t < kHold; ++t) {
    const int i = (tid + t * 128) * 4;
    if (i >= D) continue;
    if (curr) *reinterpret_cast<float4*>(curr + i) = xv[t];
    if (smooth) {
      float4 o = ov[t];
      if (n2 > 0.f) { o.x /= n2; o.y /= n2; o.z /= n2; o.w /= n2; }
      *reinterpret_cast<float4*>(smooth + i) = o;
    }
  }
}

__global__ void __launch_bounds__(kThreads)
frame_post_kernel(bt_store st, const bt_batch* __restrict__ bp, bt_res_layout L) {
  const bt_batch& b = *bp;
  // (possibly) a programmatic dependent of the LAP kernel: the assignment vectors are final once the wait returns.
  // Only then may the launch behind this one start: its feature-update CTAs read the assignments without waiting.
  bt_grid_dependency_wait();
  bt_grid_launch_dependents();   // the duplicate test is queued behind this kernel and waits for its boxes
  const int k = blockIdx.y;
  const int sid = b.sid[k];
  const int n_rows = b.n_rows[k];
  const int32_t* res = reinterpret_cast<const int32_t*>(st.res + (size_t)k * L.stride);
  const int32_t* x1 = res + L.o_x;
  const int32_t* x2 = x1 + st.cap;
  const int32_t* x3 = x2 + st.cap;
  const int bx = blockIdx.x;
  if (bx * (kThreads / 8) >= n_rows) return;
  const int g = (bx * kThreads + threadIdx.x) >> 3;
  const int lane = threadIdx.x & 31, r = lane & 7, base = lane & ~7;
  bool active = g < n_rows;
  int zi = -1;
  if (active) {
    zi = x1[g];
    if (zi < 0) zi = x2[g];
    if (zi < 0) zi = x3[g];
  }
  const size_t t = (size_t)sid * st.cap + (active ? g : 0);
  double* res_tlbr = reinterpret_cast<double*>(st.resB + (size_t)k * L.strideB + L.o_tlbr_bytes);
  if (active && zi < 0 && r < 4) res_tlbr[(size_t)g * 4 + r] = st.tlbr[t * 4 + r];   // unchanged box
  active = active && zi >= 0;
  // the slot still holds initiate()'s float32 state: one lane reads the flag, the group shares it
  int f32 = (active && r == 0) ? (int)st.slot_f32[t] : 0;
  f32 = __shfl_sync(0xffffffffu, f32, base);
  if (active && r == 0 && f32) st.slot_f32[t] = 0;
  btd_update(st.mean, st.cov, st.tlbr, st.tlbr_f32, st.det_xywh, t, (size_t)sid * st.md + (active ? zi : 0), active,
             f32 != 0, res_tlbr, (size_t)g, lane);
}

// one CTA per slot: the matched ones adopt their detection's feature (independent of the Kalman update:
// runs beside it on the ctx's side stream)
__global__ void __launch_bounds__(kThreads)
frame_ema_kernel(bt_store st, const bt_batch* __restrict__ bp, bt_frame_cfg fc, bt_res_layout L) {
  const bt_batch& b = *bp;
  __shared__ float red[kThreads / 32];
  const int k = blockIdx.y;
  const int sid = b.sid[k];
  const int bx = blockIdx.x;
  if (bx >= b.n_rows[k]) return;
  const int32_t* x1 = reinterpret_cast<const int32_t*>(st.res + (size_t)k * L.stride) + L.o_x;
  int z = x1[bx];
  if (z < 0) z = x1[st.cap + bx];
  if (z < 0) z = x1[2 * (size_t)st.cap + bx];
  if (z < 0) return;       // block-uniform
  const size_t gs = (size_t)sid * st.cap + bx;
  const size_t in_row = (size_t)b.parity[k] * st.S * st.md + (size_t)sid * st.md + z;
  if (fc.f16_inputs) ema_row<true>(st, fc, gs, in_row, 0, red);
  else ema_row<false>(st, fc, gs, in_row, 0, red);
}

// births: blocks [0, init_blocks) Kalman initiate (8 lanes per birth), then one block per birth adopts
// the detection's feature (smooth = curr = xn: the new track IS the detection object, demo:556-568)
__global__ void __launch_bounds__(kThreads)
frame_births_kernel(bt_store st, bt_frame_cfg fc, int sid, int parity, const int32_t* __restrict__ d_slot,
                    const int32_t* __restrict__ d_det, int n_births, int init_blocks, int with_feat) {
  __shared__ float red[kThreads / 32];
  int bx = blockIdx.x;
  if (bx < init_blocks) {
    const int g = (bx * kThreads + threadIdx.x) >> 3;
    const bool active = g < n_births;
    const size_t s = (size_t)sid * st.md + (active ? d_det[g] : 0);
    const size_t t = (size_t)sid * st.cap + (active ? d_slot[g] : 0);
    btd_initiate(st.det_xywh32, s, st.mean, st.cov, st.tlbr, st.tlbr_f32, t, active, st.slot_f32, threadIdx.x & 31);
    return;
  }
  bx -= init_blocks;
  if (!with_feat || bx >= n_births) return;
  const size_t gs = (size_t)sid * st.cap + d_slot[bx];
  const size_t in_row = (size_t)parity * st.S * st.md + (size_t)sid * st.md + d_det[bx];
  if (fc.f16_inputs) ema_row<true>(st, fc, gs, in_row, 2, red);
  else ema_row<false>(st, fc, gs, in_row, 2, red);
}

// ---- duplicate candidates -------------------------------------------------------------------------
struct Box { double x1, y1, x2, y2; };
__device__ __forceinline__ double iou_of(const Box& a, const Box& b) {
  const double ixmin = fmax(a.x1, b.x1), iymin = fmax(a.y1, b.y1);
  const double ixmax = fmin(a.x2, b.x2), iymax = fmin(a.y2, b.y2);
  if (ixmax <= ixmin || iymax <= iymin) return 0.0;
  const double inter = (ixmax - ixmin) * (iymax - iymin);
  const double area1 = (a.x2 - a.x1) * (a.y2 - a.y1);
  const double area2 = (b.x2 - b.x1) * (b.y2 - b.y1);
  return inter / (area1 + area2 - inter);
}

// All pairs i < j of live slots closer than `limit` (IoU distance).  remove_duplicate_stracks
// (demo:1665-1680) only looks at tracked x lost pairs; both lists are subsets of the live slots, so the
// host filters this (tiny) superset by list membership without a second round trip.
// grid (j tile of 64, i tile of 256, batch entry), only tiles that can hold a pair j > i.
//
// The same launch also carries the matched tracks' feature update (EMA role, blocks past the duplicate tiles, two
// rows per CTA): it only needs the assignment vectors, which are final before the Kalman update kernel starts, so
// these CTAs do not wait for it -- they run beside the update and beside the duplicate test without a second
// stream, a fork / join event pair or a launch of their own.
__global__ void __launch_bounds__(256)
frame_dup_ema_kernel(bt_store st, const bt_batch* __restrict__ bp, bt_frame_cfg fc, bt_res_layout L, int dup_jt,
                     int dup_tiles) {
  const bt_batch& b = *bp;
  constexpr int JT = 64;
  const int k = blockIdx.y;
  if ((int)blockIdx.x >= dup_tiles) {
    __shared__ float red[8];
    const int half = threadIdx.x >> 7;
    const int row = 2 * ((int)blockIdx.x - dup_tiles) + half;
    if (row >= b.n_rows[k]) return;   // half-uniform (named barriers below)
    const int32_t* x = reinterpret_cast<const int32_t*>(st.res + (size_t)k * L.stride) + L.o_x;
    const int z1 = x[row], z2 = x[st.cap + row], z3 = x[2 * (size_t)st.cap + row];   // one round trip
    const int z = z1 >= 0 ? z1 : (z2 >= 0 ? z2 : z3);
    if (z < 0) return;
    const size_t gs = (size_t)b.sid[k] * st.cap + row;
    const size_t in_row = (size_t)b.parity[k] * st.S * st.md + (size_t)b.sid[k] * st.md + z;
    const bool held = (st.D & 3) == 0 && st.D <= 128 * 4 * 4;
    if (held) {
      if (fc.f16_inputs) ema_row_half<true>(st, fc, gs, in_row, red, half);
      else ema_row_half<false>(st, fc, gs, in_row, red, half);
    }
    return;
  }
  bt_grid_dependency_wait();   // programmatic dependent of the Kalman update that writes the boxes
  const int bx_j = (int)blockIdx.x % dup_jt, bx_i = (int)blockIdx.x / dup_jt;
  const int n = b.n_rows[k];
  if (bx_i * 256 >= n || bx_j * JT >= n) return;
  if ((bx_j + 1) * JT <= bx_i * 256) return;
  const size_t gs0 = (size_t)b.sid[k] * st.cap;
  const double* tlbr = st.tlbr + gs0 * 4;
  const float* tlbr_f32 = st.tlbr_f32 + gs0 * 4;
  const uint8_t* kind = reinterpret_cast<const uint8_t*>(st.ctrl + b.ctrl_off[k] + 8 * (size_t)b.n_pool[k]);
  int32_t* resB = reinterpret_cast<int32_t*>(st.resB + (size_t)k * L.strideB);
  int32_t* pair_count = resB + L.o_hdr;
  int32_t* pairs_small = resB + L.o_pairs;
  int32_t* pairs = st.pairs + (size_t)k * 2 * st.pair_cap;
  __shared__ float4 sb[JT];
  __shared__ uint8_t sk[JT];
  const int i = bx_i * 256 + threadIdx.x;
  const int j0 = bx_j * JT;
  const bool live_i = i < n && kind[i] != 0;
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  if (live_i) a = *reinterpret_cast<const float4*>(tlbr_f32 + (size_t)i * 4);
  const float area_a = (a.z - a.x) * (a.w - a.y);
  const int jj = j0 + threadIdx.x;
  if (threadIdx.x < JT) {
    sk[threadIdx.x] = (jj < n) ? kind[jj] : 0;
    sb[threadIdx.x] = (jj < n) ? *reinterpret_cast<const float4*>(tlbr_f32 + (size_t)jj * 4) : a;
  }
  __syncthreads();
  if (!live_i) return;
  const int lim = min(JT, n - j0);
#pragma unroll 4
  for (int q = 0; q < lim; ++q) {
    const int j = j0 + q;
    const float4 bb = sb[q];
    // fp32 screen on the outward-rounded boxes: a pair this far from IoU 0.5 cannot reach 1 - limit
    const float iw = fminf(a.z, bb.z) - fmaxf(a.x, bb.x), ih = fminf(a.w, bb.w) - fmaxf(a.y, bb.y);
    const float inter = iw * ih;
    const float area_b = (bb.z - bb.x) * (bb.w - bb.y);
    if (iw <= 0.f || ih <= 0.f || j <= i || sk[q] == 0 || inter < 0.5f * fmaxf(area_a, area_b)) continue;
    const double* pa = tlbr + (size_t)i * 4;
    const double* pb = tlbr + (size_t)j * 4;
    const Box A{pa[0], pa[1], pa[2], pa[3]}, B{pb[0], pb[1], pb[2], pb[3]};
    if (1.0 - iou_of(A, B) < fc.dup_limit) {
      const int slot = atomicAdd(pair_count, 1);
      if (slot < st.pair_cap) { pairs[2 * slot] = i; pairs[2 * slot + 1] = j; }
      if (slot < fc.prefetch_pairs) { pairs_small[2 * slot] = i; pairs_small[2 * slot + 1] = j; }
    }
  }
}

// ---- gathers -------------------------------------------------------------------------------------------
__global__ void gather_rows_f64_kernel(const double* __restrict__ src, const int32_t* __restrict__ idx, int n,
                                       int width, double* __restrict__ dst) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)n * width) return;
  const int r = (int)(i / width), c = (int)(i % width);
  dst[i] = src[(size_t)idx[r] * width + c];
}
__global__ void gather_rows_f32_kernel(const float* __restrict__ src, const int32_t* __restrict__ idx, int n,
                                       int width, float* __restrict__ dst) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)n * width) return;
  const int r = (int)(i / width), c = (int)(i % width);
  dst[i] = src[(size_t)idx[r] * width + c];
}
__global__ void gather_curr_f16_kernel(const __half* __restrict__ feat16, const float* __restrict__ norm,
                                       const int32_t* __restrict__ idx, int n, int d, float* __restrict__ dst) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)n * d) return;
  const int r = (int)(i / d), c = (int)(i % d);
  const float nr = norm[idx[r]];
  const float x = __half2float(feat16[(size_t)idx[r] * d + c]);
  dst[i] = nr > 0.f ? x / nr : x;
}

}  // namespace

bt_res_layout bt_res_layout_for(int cap, int md, int prefetch_pairs) {
  bt_res_layout L;
  L.o_x = 0;
  L.o_sc = 3 * (size_t)cap;
  L.o_bx = (L.o_sc + md + 3) & ~size_t(3);             // int4 stores: 16 B aligned
  L.o_endA = L.o_bx + 4 * (size_t)md;
  L.stride = (sizeof(int32_t) * L.o_endA + 255) & ~size_t(255);
  L.o_hdr = 0;
  L.o_pairs = 2;
  const size_t o_endB_i = (L.o_pairs + 2 * (size_t)prefetch_pairs + 1) & ~size_t(1);
  L.o_tlbr_bytes = sizeof(int32_t) * o_endB_i;
  L.strideB = (L.o_tlbr_bytes + sizeof(double) * 4 * cap + 255) & ~size_t(255);
  return L;
}

int32_t btk_ctrl_upload(bt_ctx* ctx, const void* h_src, void* d_dst, size_t bytes) {
  const int n16 = (int)((bytes + 15) / 16);
  if (n16 <= 0) return BT_OK;
  // a plain launch: the kernels behind it are full dependents (no programmatic overlap with the upload)
  BT_CUDA(bt_launch(ctx, false, ctrl_upload_kernel, dim3((n16 + 255) / 256), dim3(256), 0,
                    reinterpret_cast<const uint4*>(h_src), reinterpret_cast<uint4*>(d_dst), n16));
  BT_LAUNCHED(ctx);
  return BT_OK;
}

int32_t btk_frame_cast(bt_ctx* ctx, const bt_store& st, const bt_batch& b, const bt_batch* db, int fixed) {
  const int mx = fixed ? st.md : bt_batch_max(b.m, b.count);
  if (mx <= 0) return BT_OK;
  BT_CUDA(bt_launch(ctx, false, frame_cast_kernel, dim3(mx, b.count), dim3(kThreads), 0, st, db));
  BT_LAUNCHED(ctx);
  return BT_OK;
}

int32_t btk_frame_prep(bt_ctx* ctx, const bt_store& st, const bt_batch& b, const bt_batch* db, const bt_frame_cfg& fc, int fixed) {
  const int mx_m = fixed ? st.md : bt_batch_max(b.m, b.count), mx_pool = fixed ? st.cap : bt_batch_max(b.n_pool, b.count);
  const int pred_blocks = (mx_pool * 8 + kThreads - 1) / kThreads;
  const int det_blocks = (mx_m + kThreads - 1) / kThreads;
  int norm_blocks = 0;
  if (fixed) norm_blocks = (fc.f16_inputs && fc.with_reid) ? st.md : 0;
  else
    for (int k = 0; k < b.count; ++k)
      if (b.want_norm[k] && b.m[k] > norm_blocks) norm_blocks = b.m[k];
  // operand prefetch blocks: only when the tensor-core association kernel follows (fp16 operands, rows x dets worth it)
  const int pf_blocks = (fc.with_reid && fc.l2_prefetch && mx_m > 0 && mx_pool > 0) ? 256 : 0;
  const int gx = pred_blocks + det_blocks + norm_blocks + pf_blocks;
  if (gx <= 0) return BT_OK;
  const bt_res_layout L = bt_res_layout_for(st.cap, st.md, fc.prefetch_pairs);
  BT_CUDA(bt_launch(ctx, false, frame_prep_kernel, dim3(gx, b.count), dim3(kThreads), 0, st, db, fc, pred_blocks, det_blocks,
                    norm_blocks, pf_blocks, L));
  BT_LAUNCHED(ctx);
  return BT_OK;
}

int32_t btk_frame_post(bt_ctx* ctx, const bt_store& st, const bt_batch& b, const bt_batch* db, const bt_frame_cfg& fc, int fixed,
                       int dependent) {
  const int mx_rows = fixed ? st.cap : bt_batch_max(b.n_rows, b.count);
  if (mx_rows <= 0) return BT_OK;
  const int upd_blocks = (mx_rows * 8 + kThreads - 1) / kThreads;
  const bt_res_layout L = bt_res_layout_for(st.cap, st.md, fc.prefetch_pairs);
  BT_CUDA(bt_launch(ctx, dependent != 0, frame_post_kernel, dim3(upd_blocks, b.count), dim3(kThreads), 0, st, db, L));
  BT_LAUNCHED(ctx);
  return BT_OK;
}

int32_t btk_frame_ema(bt_ctx* ctx, const bt_store& st, const bt_batch& b, const bt_batch* db, const bt_frame_cfg& fc,
                      cudaStream_t stream, int fixed) {
  const int mx_rows = fixed ? st.cap : bt_batch_max(b.n_rows, b.count);
  if (mx_rows <= 0) return BT_OK;
  const bt_res_layout L = bt_res_layout_for(st.cap, st.md, fc.prefetch_pairs);
  BT_CUDA(bt_launch_on(ctx, stream, false, frame_ema_kernel, dim3(mx_rows, b.count), dim3(kThreads), 0, st, db, fc, L));
  BT_LAUNCHED(ctx);
  return BT_OK;
}

int32_t btk_frame_dup(bt_ctx* ctx, const bt_store& st, const bt_batch& b, const bt_batch* db, const bt_frame_cfg& fc, int fixed,
                      int with_ema) {
  const int n = fixed ? st.cap : bt_batch_max(b.n_rows, b.count);
  const bool held = (st.D & 3) == 0 && st.D <= 128 * 4 * 4;
  if (with_ema && !held) {      // generic feature sizes: the one-row-per-CTA kernel, in stream order
    BT_TRY(btk_frame_ema(ctx, st, b, db, fc, ctx->stream, fixed));
    with_ema = 0;
  }
  const int dup_jt = n > 1 ? (n + 63) / 64 : 0, dup_it = n > 1 ? (n + 255) / 256 : 0;
  const int dup_tiles = dup_jt * dup_it;
  const int ema_ctas = (with_ema && n > 0) ? (n + 1) / 2 : 0;
  if (dup_tiles + ema_ctas <= 0) return BT_OK;
  const bt_res_layout L = bt_res_layout_for(st.cap, st.md, fc.prefetch_pairs);
  BT_CUDA(bt_launch(ctx, true, frame_dup_ema_kernel, dim3(dup_tiles + ema_ctas, b.count), dim3(256), 0, st, db, fc, L,
                    dup_jt > 0 ? dup_jt : 1, dup_tiles));
  BT_LAUNCHED(ctx);
  return BT_OK;
}

int32_t btk_frame_births(bt_ctx* ctx, const bt_store& st, int32_t sid, int32_t parity, const int32_t* d_slot,
                         const int32_t* d_det, int32_t n_births, const bt_frame_cfg& fc, int32_t with_feat) {
  if (n_births <= 0) return BT_OK;
  const int init_blocks = (n_births * 8 + kThreads - 1) / kThreads;
  frame_births_kernel<<<init_blocks + (with_feat ? n_births : 0), kThreads, 0, ctx->stream>>>(
      st, fc, sid, parity, d_slot, d_det, n_births, init_blocks, with_feat);
  BT_LAUNCHED(ctx);
  return BT_OK;
}

int32_t btk_gather_rows_f64(bt_ctx* ctx, const double* src, const int32_t* idx, int32_t n, int32_t width, double* dst) {
  if (n <= 0) return BT_OK;
  gather_rows_f64_kernel<<<(unsigned)(((size_t)n * width + 255) / 256), 256, 0, ctx->stream>>>(src, idx, n, width, dst);
  BT_LAUNCHED(ctx);
  return BT_OK;
}
int32_t btk_gather_rows_f32(bt_ctx* ctx, const float* src, const int32_t* idx, int32_t n, int32_t width, float* dst) {
  if (n <= 0) return BT_OK;
  gather_rows_f32_kernel<<<(unsigned)(((size_t)n * width + 255) / 256), 256, 0, ctx->stream>>>(src, idx, n, width, dst);
  BT_LAUNCHED(ctx);
  return BT_OK;
}
int32_t btk_gather_curr_f16(bt_ctx* ctx, const __half* feat16, const float* norm, const int32_t* idx, int32_t n,
                            int32_t d, float* dst) {
  if (n <= 0) return BT_OK;
  gather_curr_f16_kernel<<<(unsigned)(((size_t)n * d + 255) / 256), 256, 0, ctx->stream>>>(feat16, norm, idx, n, d, dst);
  BT_LAUNCHED(ctx);
  return BT_OK;
}
