// Detector-side kernels of the hot path (north_star (d)):
//   * yolox_postprocess : YOLOX head decode + per-class NMS (the part the reference keeps inside
//     its ONNX graph: README.md:179-183, README.md:197-244, model name demo:34
//     "..._post_..._score015_iou080_box050") fused with YOLOX._postprocess (demo:968-1030).
//     The in-graph arithmetic is not in the reference tree: this follows the standard YOLOX
//     decode and ONNX NonMaxSuppression-11 (parity unpinned by the reference; pinned against
//     oracle/detector_np.py).
//   * reid_crop_gather  : crop + FastReID._preprocess (demo:1434-1436, demo:1101-1142):
//     cv2.resize INTER_LINEAR 8U fixed-point bilinear (bit-exact), BGR->RGB,
//     HWC->CHW, (x/255 - mean)/std in float64 then float32.
// demo = /root/reference/demo_bottrack_onnx_tflite.py
#include "common.cuh"

namespace {

// ------------------------------------------------------------------------------------------------
// YOLOX post-process: one CTA, one group of 256 threads per class for the NMS part.
// ------------------------------------------------------------------------------------------------
constexpr int kYoloThreads = 1024;
constexpr int kMaxClasses = 4;
constexpr int kKeepCap = 64;   // >= max_per_class
constexpr int kNmsFast = 1024; // sorted candidates per class whose boxes are staged in shared memory

struct YoloScratch {
  unsigned long long* keys;  // [classes][anchors] (score bits << 32) | (0xffffffff - anchor)
  unsigned long long* sorted;
  float4* boxes;             // [anchors] decoded x1,y1,x2,y2 (model-input pixels)
  int32_t* counts;           // [classes]
};

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ float iou_f32(const float4& a, const float4& b) {
  // ONNX NonMaxSuppression (ORT SuppressByIOU), float32
  const float area1 = (a.z - a.x) * (a.w - a.y), area2 = (b.z - b.x) * (b.w - b.y);
  if (area1 <= 0.f || area2 <= 0.f) return 0.f;
  const float ix1 = fmaxf(a.x, b.x), iy1 = fmaxf(a.y, b.y), ix2 = fminf(a.z, b.z), iy2 = fminf(a.w, b.w);
  const float iw = fmaxf(ix2 - ix1, 0.f), ih = fmaxf(iy2 - iy1, 0.f);
  const float inter = iw * ih;
  return inter / (area1 + area2 - inter);
}

__device__ __forceinline__ int key_anchor(unsigned long long key) {
  return (int)(0xffffffffu - (unsigned)(key & 0xffffffffull));
}

// Dynamic shared memory: [classes][anchors] suppression flags | [classes][kNmsFast] sorted boxes (first
// used as the sort's key staging area).
__global__ void __launch_bounds__(kYoloThreads)
yolox_post_kernel(const float* __restrict__ raw, bt_yolox_config cfg, int anchors, YoloScratch sc,
                  double* __restrict__ out, int max_out, int32_t* __restrict__ out_count, int supp_bytes) {
  __shared__ int s_cnt[kMaxClasses];
  __shared__ int s_keep_idx[kMaxClasses][kKeepCap];
  __shared__ int s_nkeep[kMaxClasses];
  __shared__ int s_cursor[kMaxClasses];
  __shared__ int s_pass[kMaxClasses];
  extern __shared__ __align__(16) unsigned char s_dyn[];
  unsigned char* s_supp_raw = s_dyn;
  float4* s_boxes = reinterpret_cast<float4*>(s_dyn + supp_bytes);
  const int tid = threadIdx.x;
  const int C = cfg.num_classes, ch = 5 + C;
  if (tid < kMaxClasses) { s_cnt[tid] = 0; s_nkeep[tid] = 0; s_cursor[tid] = 0; s_pass[tid] = 0; }
  __syncthreads();

  // ---- decode + candidate selection ----
  const int w8 = cfg.in_w / 8, h8 = cfg.in_h / 8, w16 = cfg.in_w / 16, h16 = cfg.in_h / 16, w32 = cfg.in_w / 32;
  const int n8 = w8 * h8, n16 = w16 * h16;
  for (int a = tid; a < anchors; a += kYoloThreads) {
    int stride, gx, gy;
    if (a < n8) { stride = 8; gx = a % w8; gy = a / w8; }
    else if (a < n8 + n16) { const int b = a - n8; stride = 16; gx = b % w16; gy = b / w16; }
    else { const int b = a - n8 - n16; stride = 32; gx = b % w32; gy = b / w32; }
    const float* r = raw + (size_t)a * ch;
    const float s = (float)stride;
    const float cx = (r[0] + (float)gx) * s, cy = (r[1] + (float)gy) * s;
    const float bw = expf(r[2]) * s, bh = expf(r[3]) * s;
    const float x1 = cx - bw * 0.5f, y1 = cy - bh * 0.5f, x2 = cx + bw * 0.5f, y2 = cy + bh * 0.5f;
    sc.boxes[a] = make_float4(x1, y1, x2, y2);
    const float obj = sigmoidf_(r[4]);
    for (int c = 0; c < C; ++c) {
      const float score = __fmul_rn(obj, sigmoidf_(r[5 + c]));
      if (score > cfg.nms_score_thresh) {
        const int k = atomicAdd(&s_cnt[c], 1);
        sc.keys[(size_t)c * anchors + k] =
            ((unsigned long long)__float_as_uint(score) << 32) | (unsigned long long)(0xffffffffu - (unsigned)a);
      }
    }
  }
  __syncthreads();

  // ---- per class (one group of 256 threads each): rank sort (descending score, ties: lower anchor
  //      first), then greedy NMS: one sweep of the group per kept box, boxes and suppression flags in
  //      shared memory, two barriers per sweep (pick + broadcast | suppress). ----
  const int group = tid >> 8, gt = tid & 255;  // 4 groups of 256 threads
#define GROUP_SYNC() asm volatile("bar.sync %0, 256;" ::"r"(group + 1) : "memory")
  for (int c = group; c < C; c += kYoloThreads / 256) {
    const int n = s_cnt[c];
    const unsigned long long* keys = sc.keys + (size_t)c * anchors;
    unsigned long long* sorted = sc.sorted + (size_t)c * anchors;
    unsigned char* supp = s_supp_raw + (size_t)c * anchors;
    float4* sbox = s_boxes + (size_t)c * kNmsFast;
    constexpr int kStage = kNmsFast * 2;                     // u64 keys that fit into the (not yet used) box area
    constexpr int kPer = (kStage + 255) / 256;
    if (n <= kStage) {
      unsigned long long* skey = reinterpret_cast<unsigned long long*>(sbox);
      for (int i = gt; i < n; i += 256) skey[i] = keys[i];
      GROUP_SYNC();
      unsigned long long mine[kPer];
      int myrank[kPer];
#pragma unroll
      for (int t = 0; t < kPer; ++t) {
        const int i = gt + t * 256;
        if (i < n) {
          const unsigned long long k = skey[i];
          int rank = 0;
          for (int j = 0; j < n; ++j) rank += (skey[j] > k);
          mine[t] = k; myrank[t] = rank;
          supp[i] = 0;
        }
      }
      GROUP_SYNC();                                          // everybody is done reading the staged keys
#pragma unroll
      for (int t = 0; t < kPer; ++t) {
        const int i = gt + t * 256;
        if (i < n) {
          sorted[myrank[t]] = mine[t];
          if (myrank[t] < kNmsFast) sbox[myrank[t]] = sc.boxes[key_anchor(mine[t])];
        }
      }
    } else {
      for (int i = gt; i < n; i += 256) {
        const unsigned long long k = keys[i];
        int rank = 0;
        for (int j = 0; j < n; ++j) rank += (keys[j] > k);
        sorted[rank] = k;
        supp[i] = 0;
      }
      GROUP_SYNC();
      for (int i = gt; i < n && i < kNmsFast; i += 256) sbox[i] = sc.boxes[key_anchor(sorted[i])];
    }
    GROUP_SYNC();
    auto box_of = [&](int j) -> float4 { return j < kNmsFast ? sbox[j] : sc.boxes[key_anchor(sorted[j])]; };
    int next = 0;                                            // thread 0: where the search for the next survivor starts
    while (true) {
      if (gt == 0) {
        int i = next;
        while (i < n && supp[i]) ++i;
        if (i < n && s_nkeep[c] < cfg.max_per_class) {
          s_keep_idx[c][s_nkeep[c]] = i;
          s_nkeep[c] += 1;
          s_cursor[c] = i;
          next = i + 1;
        } else {
          s_cursor[c] = n;  // done
        }
      }
      GROUP_SYNC();
      const int i = s_cursor[c];
      if (i >= n) break;
      if (s_nkeep[c] < cfg.max_per_class) {                  // the last kept box suppresses nobody that matters
        const float4 bi = box_of(i);
        for (int j = i + 1 + gt; j < n; j += 256)
          if (!supp[j] && iou_f32(bi, box_of(j)) > cfg.nms_iou_thresh) supp[j] = 1;
      }
      GROUP_SYNC();
    }
  }
#undef GROUP_SYNC
  __syncthreads();

  // ---- YOLOX._postprocess (demo:1001-1027): score filter, rescale, truncate; class-major order.
  //      One thread per kept box; the boxes of a class are sorted by score, so the ones that pass the
  //      score filter are a prefix of its list and the output position is a prefix sum over classes. ----
  const int pc = tid / kKeepCap, pk = tid % kKeepCap;
  bool pass = false;
  float score = 0.f;
  float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
  if (pc < C && pk < s_nkeep[pc]) {
    const unsigned long long key = sc.sorted[(size_t)pc * anchors + s_keep_idx[pc][pk]];
    score = __uint_as_float((unsigned)(key >> 32));
    pass = score > cfg.post_score_thresh;
    b = sc.boxes[key_anchor(key)];
    if (pass) atomicAdd(&s_pass[pc], 1);
  }
  __syncthreads();
  if (pass) {
    int pos = pk;
    for (int c = 0; c < pc; ++c) pos += s_pass[c];
    if (pos < max_out) {
      const float inw = (float)cfg.in_w, inh = (float)cfg.in_h;
      const float imw = (float)cfg.img_w, imh = (float)cfg.img_h;
      // float32 multiply then divide, truncation toward zero (demo:1009-1012)
      const int x_min = (int)__fdiv_rn(__fmul_rn(fmaxf(0.f, b.x), imw), inw);
      const int y_min = (int)__fdiv_rn(__fmul_rn(fmaxf(0.f, b.y), imh), inh);
      const int x_max = (int)__fdiv_rn(__fmul_rn(fminf(b.z, inw), imw), inw);
      const int y_max = (int)__fdiv_rn(__fmul_rn(fminf(b.w, inh), imh), inh);
      double* o = out + (size_t)pos * 6;
      o[0] = (double)pc; o[1] = (double)score;
      o[2] = (double)x_min; o[3] = (double)y_min; o[4] = (double)x_max; o[5] = (double)y_max;
    }
  }
  if (tid == 0) {
    int n_out = 0;
    for (int c = 0; c < C; ++c) n_out += s_pass[c];
    *out_count = n_out < max_out ? n_out : max_out;
  }
}

// ------------------------------------------------------------------------------------------------
// ReID crop gather
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void resize_coord(int d, int src, int dst, bool is_x, int* s0, int* s1, int* a0,
                                             int* a1) {
  // OpenCV resize (INTER_LINEAR, 8U): scale = 1 / (dst / src) in double, f in float32
  const double inv = (double)dst / (double)src;
  const double scale = 1.0 / inv;
  float f = (float)(((double)d + 0.5) * scale - 0.5);
  int s = (int)floorf(f);
  f -= (float)s;
  if (is_x) {
    if (s < 0) { f = 0.f; s = 0; }
    if (s >= src - 1) { f = 0.f; s = src - 1; }
    *s0 = s;
    *s1 = min(s + 1, src - 1);
  } else {
    *s0 = min(max(s, 0), src - 1);
    *s1 = min(max(s + 1, 0), src - 1);
  }
  // saturate_cast<short>(w * 2048): round half to even (cvRound)
  *a0 = __float2int_rn(__fmul_rn(1.f - f, 2048.f));
  *a1 = __float2int_rn(__fmul_rn(f, 2048.f));
}

__global__ void __launch_bounds__(256)
reid_crop_kernel(const uint8_t* __restrict__ frame, int h, int w, const int32_t* __restrict__ boxes, int out_h,
                 int out_w, float* __restrict__ out, const int32_t* __restrict__ n_dev) {
  const int det = blockIdx.y;
  if (n_dev && det >= *n_dev) return;       // device-chained use: the number of bodies never visits the host
  const int pix = blockIdx.x * 256 + threadIdx.x;
  if (pix >= out_h * out_w) return;
  const int dy = pix / out_w, dx = pix % out_w;
  const int4 b = *reinterpret_cast<const int4*>(boxes + (size_t)det * 4);
  // numpy slicing image[y1:y2, x1:x2] clamps to the image
  const int x1 = min(max(b.x, 0), w), y1 = min(max(b.y, 0), h);
  const int x2 = min(max(b.z, 0), w), y2 = min(max(b.w, 0), h);
  const int sw = x2 - x1, sh = y2 - y1;
  const float mean[3] = {0.485f, 0.456f, 0.406f};
  const float stdv[3] = {0.229f, 0.224f, 0.225f};
  float* o = out + (size_t)det * 3 * out_h * out_w;
  if (sw <= 0 || sh <= 0) {
    for (int c = 0; c < 3; ++c) o[(size_t)c * out_h * out_w + pix] = 0.f;
    return;
  }
  int sx0, sx1, ax0, ax1, sy0, sy1, by0, by1;
  resize_coord(dx, sw, out_w, true, &sx0, &sx1, &ax0, &ax1);
  resize_coord(dy, sh, out_h, false, &sy0, &sy1, &by0, &by1);
  const uint8_t* r0 = frame + ((size_t)(y1 + sy0) * w + x1) * 3;
  const uint8_t* r1 = frame + ((size_t)(y1 + sy1) * w + x1) * 3;
#pragma unroll
  for (int c = 0; c < 3; ++c) {  // c = BGR channel of the source
    const int h0 = (int)r0[sx0 * 3 + c] * ax0 + (int)r0[sx1 * 3 + c] * ax1;
    const int h1 = (int)r1[sx0 * 3 + c] * ax0 + (int)r1[sx1 * 3 + c] * ax1;
    const int v = (((by0 * (h0 >> 4)) >> 16) + ((by1 * (h1 >> 4)) >> 16) + 2) >> 2;
    const int pv = min(max(v, 0), 255);
    const int rgb = 2 - c;  // [..., ::-1]
    const double val = ((double)pv / 255.0 - (double)mean[rgb]) / (double)stdv[rgb];
    o[(size_t)rgb * out_h * out_w + pix] = (float)val;
  }
}

}  // namespace

int32_t btk_yolox_postprocess(bt_ctx* ctx, const float* raw, const bt_yolox_config& cfg, double* out_boxes,
                              int32_t max_out, int32_t* out_count) {
  BT_CHECK(cfg.num_classes >= 1 && cfg.num_classes <= kMaxClasses, BT_ERR_INVALID, "num_classes must be 1..%d",
           kMaxClasses);
  BT_CHECK(cfg.max_per_class >= 1 && cfg.max_per_class <= kKeepCap, BT_ERR_INVALID, "max_per_class must be 1..%d",
           kKeepCap);
  BT_CHECK(cfg.in_h % 32 == 0 && cfg.in_w % 32 == 0 && cfg.in_h > 0 && cfg.in_w > 0, BT_ERR_INVALID,
           "input size must be a positive multiple of 32");
  const int anchors = (cfg.in_h / 8) * (cfg.in_w / 8) + (cfg.in_h / 16) * (cfg.in_w / 16) +
                      (cfg.in_h / 32) * (cfg.in_w / 32);
  const size_t supp_bytes = ((size_t)cfg.num_classes * anchors + 15) & ~size_t(15);
  const size_t smem = supp_bytes + (size_t)kMaxClasses * kNmsFast * sizeof(float4);
  BT_CHECK(smem <= 220 * 1024, BT_ERR_CAPACITY, "too many anchors (%d) for the single-CTA NMS", anchors);
  YoloScratch sc;
  BT_TRY(bt_arena(ctx, (size_t)cfg.num_classes * anchors, &sc.keys));
  BT_TRY(bt_arena(ctx, (size_t)cfg.num_classes * anchors, &sc.sorted));
  BT_TRY(bt_arena(ctx, (size_t)anchors, &sc.boxes));
  sc.counts = nullptr;
  BT_CUDA(cudaFuncSetAttribute(yolox_post_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  yolox_post_kernel<<<1, kYoloThreads, smem, ctx->stream>>>(raw, cfg, anchors, sc, out_boxes, max_out, out_count,
                                                            (int)supp_bytes);
  BT_LAUNCHED(ctx);
  return BT_OK;
}

int32_t btk_reid_crop_gather(bt_ctx* ctx, const uint8_t* frame, int32_t h, int32_t w, const int32_t* boxes,
                             int32_t n, int32_t out_h, int32_t out_w, float* out) {
  if (n <= 0) return BT_OK;
  dim3 grid((out_h * out_w + 255) / 256, n);
  reid_crop_kernel<<<grid, 256, 0, ctx->stream>>>(frame, h, w, boxes, out_h, out_w, out, nullptr);
  BT_LAUNCHED(ctx);
  return BT_OK;
}

namespace {
// The decoded list is ordered by class, then by descending score: the class-0 (body) rows are its head.  They
// become the tracker's detections of the frame, written where the next bt_submit_streams of the stream expects
// them; rows past the last body get score 0 (below track_low_thresh: ignored by every association stage and never
// born), so the tracker can be stepped with m = max_bodies without the count ever reaching the host.
__global__ void stage_bodies_kernel(const double* __restrict__ det, const int32_t* __restrict__ det_count, int max_bodies,
                                    int32_t* __restrict__ boxes, float* __restrict__ scores, int32_t* __restrict__ n_bodies) {
  const int j = threadIdx.x;
  const int cnt = *det_count;
  const bool body = j < max_bodies && j < cnt && det[(size_t)j * 6] == 0.0;
  if (j < max_bodies) {
    int4 b = make_int4(0, 0, 0, 0);
    float sc = 0.f;
    if (body) {
      const double* r = det + (size_t)j * 6;
      b = make_int4((int)r[2], (int)r[3], (int)r[4], (int)r[5]);
      sc = (float)r[1];
    }
    *reinterpret_cast<int4*>(boxes + (size_t)j * 4) = b;
    scores[j] = sc;
  }
  const int nb = __syncthreads_count(body ? 1 : 0);
  if (j == 0) *n_bodies = nb;
}
}  // namespace


extern "C" {

int32_t bt_yolox_postprocess(bt_ctx* ctx, const float* raw_head, const bt_yolox_config* cfg, double* out_boxes,
                             int32_t max_out, int32_t* out_count, int32_t loc) {
  if (!ctx) return BT_ERR_INVALID;
  BT_CUDA(cudaSetDevice(ctx->device));
  BT_CHECK(loc == BT_HOST || loc == BT_DEVICE, BT_ERR_INVALID, "bad loc");
  BT_CHECK(raw_head && cfg && out_boxes && out_count && max_out > 0, BT_ERR_INVALID, "NULL buffer / bad max_out");
  BT_CHECK(cfg->in_h % 32 == 0 && cfg->in_w % 32 == 0 && cfg->in_h > 0 && cfg->in_w > 0, BT_ERR_INVALID,
           "input size must be a positive multiple of 32");
  const int anchors = (cfg->in_h / 8) * (cfg->in_w / 8) + (cfg->in_h / 16) * (cfg->in_w / 16) +
                      (cfg->in_h / 32) * (cfg->in_w / 32);
  const int ch = 5 + cfg->num_classes;
  size_t need = (size_t)anchors * (16 + 16 * kMaxClasses) + 4096;
  if (loc == BT_HOST) need += (size_t)anchors * ch * 4 + (size_t)max_out * 48 + 256;
  BT_TRY(bt_arena_reserve(ctx, need));
  const float* d_raw; double* d_out; int32_t* d_cnt;
  BT_TRY(bt_in(ctx, raw_head, (size_t)anchors * ch, loc, &d_raw));
  BT_TRY(bt_out(ctx, out_boxes, (size_t)max_out * 6, loc, &d_out));
  BT_TRY(bt_out(ctx, out_count, (size_t)1, loc, &d_cnt));
  BT_TRY(btk_yolox_postprocess(ctx, d_raw, *cfg, d_out, max_out, d_cnt));
  BT_TRY(bt_unstage_out(ctx, out_boxes, d_out, sizeof(double) * 6 * max_out, loc));
  BT_TRY(bt_unstage_out(ctx, out_count, d_cnt, sizeof(int32_t), loc));
  return bt_finish(ctx, loc);
}

int32_t bt_detect_stage(bt_ctx* ctx, int32_t stream_id, const float* raw_head, const bt_yolox_config* cfg,
                        const uint8_t* frame, int32_t h, int32_t w, int32_t out_h, int32_t out_w, float* crops,
                        double* det_out, int32_t max_out, int32_t* det_count) {
  if (!ctx) return BT_ERR_INVALID;
  BT_CUDA(cudaSetDevice(ctx->device));
  BT_CHECK(raw_head && cfg && frame && crops, BT_ERR_INVALID, "NULL buffer");
  BT_CHECK(h > 0 && w > 0 && out_h > 0 && out_w > 0, BT_ERR_INVALID, "bad size");
  BT_CHECK(cfg->in_h % 32 == 0 && cfg->in_w % 32 == 0 && cfg->in_h > 0 && cfg->in_w > 0, BT_ERR_INVALID,
           "input size must be a positive multiple of 32");
  BT_CHECK(cfg->max_per_class >= 1 && cfg->max_per_class <= ctx->max_dets && cfg->max_per_class <= 1024, BT_ERR_CAPACITY,
           "max_per_class %d exceeds the ctx's max_dets %d", cfg->max_per_class, ctx->max_dets);
  const int max_bodies = cfg->max_per_class;
  const int list_cap = det_out ? max_out : cfg->num_classes * cfg->max_per_class;
  BT_CHECK(list_cap >= max_bodies, BT_ERR_INVALID, "max_out %d is smaller than max_per_class %d", list_cap, max_bodies);
  int32_t* in_boxes = nullptr; float* in_scores = nullptr;
  BT_TRY(bt_input_buffers(ctx, stream_id, &in_boxes, &in_scores, nullptr));
  const int anchors = (cfg->in_h / 8) * (cfg->in_w / 8) + (cfg->in_h / 16) * (cfg->in_w / 16) +
                      (cfg->in_h / 32) * (cfg->in_w / 32);
  // scratch from the ctx arena (grown on the first call only: no synchronisation in steady state)
  BT_TRY(bt_arena_reserve(ctx, (size_t)anchors * (16 + 16 * kMaxClasses) + 4096 + (size_t)list_cap * 48 + 1024));
  double* d_det = det_out;
  int32_t* d_cnt = det_count;
  if (!d_det) BT_TRY(bt_arena(ctx, (size_t)list_cap * 6, &d_det));
  if (!d_cnt) BT_TRY(bt_arena(ctx, (size_t)4, &d_cnt));
  int32_t* d_nb = nullptr;
  BT_TRY(bt_arena(ctx, (size_t)4, &d_nb));
  BT_TRY(btk_yolox_postprocess(ctx, raw_head, *cfg, d_det, list_cap, d_cnt));
  stage_bodies_kernel<<<1, 1024, 0, ctx->stream>>>(d_det, d_cnt, max_bodies, in_boxes, in_scores, d_nb);
  BT_LAUNCHED(ctx);
  dim3 grid((out_h * out_w + 255) / 256, max_bodies);
  reid_crop_kernel<<<grid, 256, 0, ctx->stream>>>(frame, h, w, in_boxes, out_h, out_w, crops, d_nb);
  BT_LAUNCHED(ctx);
  return BT_OK;
}

int32_t bt_reid_crop_gather(bt_ctx* ctx, const uint8_t* frame, int32_t h, int32_t w, const int32_t* boxes,
                            int32_t n, int32_t out_h, int32_t out_w, float* out, int32_t loc) {
  if (!ctx) return BT_ERR_INVALID;
  BT_CUDA(cudaSetDevice(ctx->device));
  BT_CHECK(loc == BT_HOST || loc == BT_DEVICE, BT_ERR_INVALID, "bad loc");
  BT_CHECK(n >= 0 && h > 0 && w > 0 && out_h > 0 && out_w > 0, BT_ERR_INVALID, "bad size");
  if (n == 0) return BT_OK;
  BT_CHECK(frame && boxes && out, BT_ERR_INVALID, "NULL buffer");
  const size_t out_elems = (size_t)n * 3 * out_h * out_w;
  BT_TRY(bt_arena_reserve(ctx, loc == BT_HOST ? (size_t)h * w * 3 + (size_t)n * 16 + out_elems * 4 : 0));
  const uint8_t* d_frame; const int32_t* d_boxes; float* d_out;
  BT_TRY(bt_in(ctx, frame, (size_t)h * w * 3, loc, &d_frame));
  BT_TRY(bt_in(ctx, boxes, (size_t)n * 4, loc, &d_boxes));
  BT_TRY(bt_out(ctx, out, out_elems, loc, &d_out));
  BT_TRY(btk_reid_crop_gather(ctx, d_frame, h, w, d_boxes, n, out_h, out_w, d_out));
  BT_TRY(bt_unstage_out(ctx, out, d_out, sizeof(float) * out_elems, loc));
  return bt_finish(ctx, loc);
}

}  // extern "C"
