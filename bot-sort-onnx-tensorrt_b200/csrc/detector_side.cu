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

#include <mutex>

namespace {

// ------------------------------------------------------------------------------------------------
// YOLOX post-process: one thread-block CLUSTER, one CTA of 1024 threads per class.
//   phase 1  every CTA scans the raw head for ITS class (obj and class logit only; 2 of the 5+C words
//            of a row) and appends the candidates (score > nms_score_thresh) to a shared-memory list;
//   phase 2  rank sort of the list (descending score, ties: lower anchor first) with all 1024 threads
//            (several threads per element when the list is short), boxes decoded from the head only for
//            the candidates and stored in sorted order;
//   phase 3  greedy NMS on an "alive" bit mask: every warp finds the next survivor by itself (a word
//            scan + ffs), tests the candidates it owns against it and clears their bits -- ONE barrier
//            per kept box and no thread-0 serial section;
//   phase 4  YOLOX._postprocess (demo:1001-1027) for the kept boxes; the classes exchange their row counts
//            through distributed shared memory and write the class-major output list; the body CTA also
//            stages its rows into the tracker's input buffers.
// ------------------------------------------------------------------------------------------------
constexpr int kYoloThreads = 1024;
constexpr int kMaxClasses = 4;
constexpr int kKeepCap = 64;   // >= max_per_class

struct YoloArgs {
  const float* raw;                // [anchors][5 + C]
  bt_yolox_config cfg;
  int anchors;
  unsigned long long* sorted;      // global scratch [C][anchors]: (score bits << 32) | (0xffffffff - anchor), sorted
  double* out;                     // [max_out][6] or nullptr
  int max_out;
  int32_t* out_count;              // or nullptr
  // staging of the class-0 rows (optional)
  int32_t* st_boxes;               // [max_bodies][4]
  float* st_scores;                // [max_bodies]
  int32_t* st_nbodies;
  int max_bodies;
  int debug;                       // BT_YOLOX_DEBUG=1: phase timestamps (ns) by device printf
};

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ bool iou_above(const float4& a, const float4& b, float thr) {
  // ONNX NonMaxSuppression (ORT SuppressByIOU), float32: the division only happens for overlapping pairs
  const float ix1 = fmaxf(a.x, b.x), iy1 = fmaxf(a.y, b.y), ix2 = fminf(a.z, b.z), iy2 = fminf(a.w, b.w);
  const float iw = fmaxf(ix2 - ix1, 0.f), ih = fmaxf(iy2 - iy1, 0.f);
  const float inter = iw * ih;
  if (!(inter > 0.f)) return false;                    // iou == 0 (thr >= 0) -- or an empty box below
  const float area1 = (a.z - a.x) * (a.w - a.y), area2 = (b.z - b.x) * (b.w - b.y);
  if (area1 <= 0.f || area2 <= 0.f) return false;
  return inter / (area1 + area2 - inter) > thr;
}

// The same decision without the division on the critical path: inter / uni > thr is decided from inter against
// thr * uni whenever the two are further apart than any rounding of the division (or of the product) can bridge;
// only a near tie (relative distance below 2^-20) takes the exact path.  Branch-free otherwise, so that several
// tests of one thread overlap (the NMS loop is a latency chain).
__device__ __forceinline__ float box_area(const float4& a) { return (a.z - a.x) * (a.w - a.y); }
__device__ __forceinline__ bool iou_above_fast(const float4& a, float area_a, const float4& b, float area_b, float thr,
                                               bool& unsure) {
  const float iw = fmaxf(fminf(a.z, b.z) - fmaxf(a.x, b.x), 0.f);
  const float ih = fmaxf(fminf(a.w, b.w) - fmaxf(a.y, b.y), 0.f);
  const float inter = iw * ih;
  const float t = __fmul_rn(thr, area_a + area_b - inter);
  const bool valid = inter > 0.f && fminf(area_a, area_b) > 0.f;
  const bool hi = inter > __fmul_rn(t, 1.000001f);
  unsure = unsure || (valid && !hi && inter > __fmul_rn(t, 0.999999f));   // near tie: the caller takes the exact path
  return valid && hi;
}
// the exact test (the original arithmetic), kept out of line: the NMS loop is bound by the instructions one warp
// walks through per trip, and near ties are rare
__device__ __noinline__ bool iou_above_exact(float4 a, float4 b, float thr) { return iou_above(a, b, thr); }

__device__ __forceinline__ int key_anchor(unsigned long long key) {
  return (int)(0xffffffffu - (unsigned)(key & 0xffffffffull));
}

__device__ __forceinline__ float4 decode_box(const float* __restrict__ raw, int a, int ch, int in_h, int in_w) {
  const int w8 = in_w / 8, h8 = in_h / 8, w16 = in_w / 16, h16 = in_h / 16, w32 = in_w / 32;
  const int n8 = w8 * h8, n16 = w16 * h16;
  int stride, gx, gy;
  if (a < n8) { stride = 8; gx = a % w8; gy = a / w8; }
  else if (a < n8 + n16) { const int b = a - n8; stride = 16; gx = b % w16; gy = b / w16; }
  else { const int b = a - n8 - n16; stride = 32; gx = b % w32; gy = b / w32; }
  const float4 r = make_float4(raw[(size_t)a * ch], raw[(size_t)a * ch + 1], raw[(size_t)a * ch + 2], raw[(size_t)a * ch + 3]);
  const float s = (float)stride;
  const float cx = (r.x + (float)gx) * s, cy = (r.y + (float)gy) * s;
  const float bw = expf(r.z) * s, bh = expf(r.w) * s;
  return make_float4(cx - bw * 0.5f, cy - bh * 0.5f, cx + bw * 0.5f, cy + bh * 0.5f);
}

__device__ __forceinline__ unsigned cluster_ctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ int dsmem_read_i32(const int* local_ptr, unsigned cta) {
  unsigned src = (unsigned)__cvta_generic_to_shared(local_ptr), dst;
  int v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(dst) : "r"(src), "r"(cta));
  asm volatile("ld.shared::cluster.s32 %0, [%1];" : "=r"(v) : "r"(dst) : "memory");
  return v;
}

// Dynamic shared memory: u64 keys[anchors (+ pad)] | float4 boxes[anchors] (sorted order) | u32 alive[2][(anchors + 31) / 32]
__global__ void __launch_bounds__(kYoloThreads, 1) yolox_post_kernel(YoloArgs A) {
  __shared__ int s_n;
  __shared__ int s_keep[kKeepCap];
  __shared__ int s_pass;                     // read by the other CTAs of the cluster
  extern __shared__ __align__(16) unsigned char s_dyn[];
  const int anchors = A.anchors;
  unsigned long long* s_keys = reinterpret_cast<unsigned long long*>(s_dyn);
  const size_t keys_cap = ((size_t)anchors + 2) & ~size_t(1);      // room for the pair padding of an odd list
  float4* s_box = reinterpret_cast<float4*>(s_dyn + keys_cap * 8);
  unsigned* s_alive = reinterpret_cast<unsigned*>(s_dyn + keys_cap * 8 + (size_t)anchors * 16);
  const int tid = threadIdx.x, lane = tid & 31;
  const int c = (int)cluster_ctarank();      // the class of this CTA
  const int C = A.cfg.num_classes, ch = 5 + C;
  const float* __restrict__ raw = A.raw;
  long long t0 = 0, t1 = 0, t2 = 0, t3 = 0, t4 = 0;
  auto now = []() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; };
  if (A.debug) t0 = now();
  if (tid == 0) { s_n = 0; s_pass = 0; }
  __syncthreads();

  const float thr_s = fminf(fmaxf(A.cfg.nms_score_thresh, 1e-6f), 0.999999f);
  const float logit_lo = A.cfg.nms_score_thresh > 0.f ? logf(thr_s / (1.f - thr_s)) - 0.05f : -INFINITY;
  // ---- phase 1: candidates of class c (the loads of four anchors are in flight together: the head is cold) ----
  for (int a0 = tid; a0 < anchors; a0 += 4 * kYoloThreads) {
    float obj[4], cls[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int a = a0 + u * kYoloThreads;
      const float* r = raw + (size_t)(a < anchors ? a : a0) * ch;
      obj[u] = r[4]; cls[u] = r[5 + c];
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int a = a0 + u * kYoloThreads;
      // obj * cls > thr needs obj > thr and cls > thr (both factors are below 1): the sigmoids are only evaluated
      // for logits above logit(thr) - 0.05 (a conservative screen; the decision itself is made on the product)
      if (a < anchors && fminf(obj[u], cls[u]) > logit_lo) {
        const float score = __fmul_rn(sigmoidf_(obj[u]), sigmoidf_(cls[u]));
        if (score > A.cfg.nms_score_thresh) {
          const int k = atomicAdd(&s_n, 1);
          s_keys[k] = ((unsigned long long)__float_as_uint(score) << 32) | (unsigned long long)(0xffffffffu - (unsigned)a);
        }
      }
    }
  }
  __syncthreads();
  if (A.debug) t1 = now();
  const int n = s_n;
  const int nwords = (n + 31) >> 5;
  unsigned long long* sorted = A.sorted + (size_t)c * anchors;

  // ---- phase 2: rank sort; P = 2^p adjacent lanes share one element when the list is short; a lane counts
  //      over interleaved PAIRS of keys (one 16-byte load each).  The box of the element is decoded first so that
  //      its (cold) head row arrives while the ranks are counted. ----
  // A thread ranks TWO elements (i and i + half) against every pair it loads.
  const int half = (n + 1) >> 1;
  int P = 1;
  while (P < 32 && half * (P * 2) <= kYoloThreads) P *= 2;
  if (tid == 0 && (n & 1)) s_keys[n] = 0ull;           // pad to a pair (a zero key is smaller than every candidate)
  __syncthreads();
  const int npairs = (n + 1) >> 1;
  const ulonglong2* s_pairs = reinterpret_cast<const ulonglong2*>(s_keys);
  for (int base = 0; base < half; base += kYoloThreads / P) {
    const int i = base + tid / P, part = tid % P;
    const int i2 = i + half;
    int rank = 0, rank2 = 0;
    unsigned long long k = 0, k2 = 0;
    float4 box = make_float4(0.f, 0.f, 0.f, 0.f), box2 = box;
    if (i < half) {
      k = s_keys[i];
      k2 = i2 < n ? s_keys[i2] : ~0ull;
      if (part == 0) box = decode_box(raw, key_anchor(k), ch, A.cfg.in_h, A.cfg.in_w);
      if (part == (P > 1 ? 1 : 0) && i2 < n) box2 = decode_box(raw, key_anchor(k2), ch, A.cfg.in_h, A.cfg.in_w);
#pragma unroll 4
      for (int j = part; j < npairs; j += P) {
        const ulonglong2 kk = s_pairs[j];
        rank += (kk.x > k) + (kk.y > k);
        rank2 += (kk.x > k2) + (kk.y > k2);
      }
    }
    for (int o = 1; o < P; o <<= 1) {
      rank += __shfl_xor_sync(0xffffffffu, rank, o);
      rank2 += __shfl_xor_sync(0xffffffffu, rank2, o);
    }
    if (i < half && part == 0) {
      s_box[rank] = box;
      sorted[rank] = k;
    }
    if (i < half && i2 < n && part == (P > 1 ? 1 : 0)) {
      s_box[rank2] = box2;
      sorted[rank2] = k2;
    }
  }
  for (int w = tid; w < nwords; w += kYoloThreads)
    s_alive[w] = (w == nwords - 1 && (n & 31)) ? ((1u << (n & 31)) - 1u) : 0xffffffffu;
  __syncthreads();
  if (A.debug) t2 = now();

  // ---- phase 3: greedy NMS, four survivors per barrier.  The loop is a latency chain (search -> box -> IoU ->
  //      vote -> barrier, about 500 cycles), so each trip does as much as the data allows: every warp finds the next
  //      FOUR alive candidates by itself, lanes 0..5 test the six pairs among them while every thread already tests
  //      its own candidate against all four, the greedy choice inside the batch is replayed from the six pair bits
  //      (identical in every warp), and a thread drops its candidate if a KEPT member suppresses it.  Only the warps
  //      that own candidates take part (named barrier); a thread keeps its candidate's alive bit in a register, the
  //      shared words serve the search only. ----
  const int max_keep = A.cfg.max_per_class;
  const float thr = A.cfg.nms_iou_thresh;
  int kept = 0;
  const int act_warps = min(kYoloThreads / 32, (n + 31) >> 5);
  if ((tid >> 5) < act_warps) {
    const int act_threads = act_warps * 32;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 my_box = tid < n ? s_box[tid] : zero4;
    bool my_alive = tid < n;
    int pos = 0;
    const int wcap = (anchors + 31) >> 5;
    unsigned* cur = s_alive;            // two copies of the alive words: a warp publishes its word for the NEXT trip while
    unsigned* nxt = s_alive + wcap;     // slower warps may still be searching the current one
    const float my_area = box_area(my_box);
    while (kept < max_keep) {
      // the next (up to) four alive candidates c0 < c1 < c2 < c3: the lowest set bits of the first non-empty word
      int w = pos >> 5;
      unsigned m = w < nwords ? (cur[w] & (0xffffffffu << (pos & 31))) : 0u;
      while (m == 0u && ++w < nwords) m = cur[w];
      if (m == 0u) break;
      const int cnt = min(__popc(m), 4);
      const unsigned b0 = m & (0u - m); m ^= b0;
      const unsigned b1 = m & (0u - m); m ^= b1;
      const unsigned b2 = m & (0u - m); m ^= b2;
      const unsigned b3 = m & (0u - m);
      const int wb = w * 32 - 1;
      const int c0 = wb + __ffs(b0);
      const int c1 = cnt > 1 ? wb + __ffs(b1) : c0;    // absent members alias c0 (their loads stay in range) ...
      const int c2 = cnt > 2 ? wb + __ffs(b2) : c0;
      const int c3 = cnt > 3 ? wb + __ffs(b3) : c0;
      const float4 B0 = s_box[c0];
      float4 B1 = s_box[c1], B2 = s_box[c2], B3 = s_box[c3];
      if (cnt < 2) B1 = zero4;                         // ... and become empty boxes: they never suppress
      if (cnt < 3) B2 = zero4;
      if (cnt < 4) B3 = zero4;
      const float A0 = box_area(B0), A1 = box_area(B1), A2 = box_area(B2), A3 = box_area(B3);
      // pair p of lanes 0..5: (0,1) (0,2) (0,3) (1,2) (1,3) (2,3)
      const float4 pa = lane < 3 ? B0 : (lane < 5 ? B1 : B2);
      const float pa_area = lane < 3 ? A0 : (lane < 5 ? A1 : A2);
      const float4 pb = lane == 0 ? B1 : ((lane == 1 || lane == 3) ? B2 : B3);
      const float pb_area = lane == 0 ? A1 : ((lane == 1 || lane == 3) ? A2 : A3);
      bool unsure = false;
      bool pair_sup = iou_above_fast(pa, pa_area, pb, pb_area, thr, unsure);
      // this thread's candidate against the four (used only for the members that end up kept)
      bool s0 = iou_above_fast(B0, A0, my_box, my_area, thr, unsure);
      bool s1 = iou_above_fast(B1, A1, my_box, my_area, thr, unsure);
      bool s2 = iou_above_fast(B2, A2, my_box, my_area, thr, unsure);
      bool s3 = iou_above_fast(B3, A3, my_box, my_area, thr, unsure);
      if (unsure) {
        pair_sup = iou_above_exact(pa, pb, thr);
        s0 = iou_above_exact(B0, my_box, thr); s1 = iou_above_exact(B1, my_box, thr);
        s2 = iou_above_exact(B2, my_box, thr); s3 = iou_above_exact(B3, my_box, thr);
      }
      pair_sup = pair_sup && lane < 6;
      s0 = s0 && tid > c0; s1 = s1 && tid > c1; s2 = s2 && tid > c2; s3 = s3 && tid > c3;
      const unsigned M = __ballot_sync(0xffffffffu, pair_sup);
      const bool m01 = M & 1u, m02 = M & 2u, m03 = M & 4u, m12 = M & 8u, m13 = M & 16u, m23 = M & 32u;
      const int kept0 = kept;
      kept += 1;                                       // c0 is kept (kept < max_keep holds here)
      const bool k1 = cnt > 1 && kept < max_keep && !m01;
      kept += k1 ? 1 : 0;
      const bool k2 = cnt > 2 && kept < max_keep && !(m02 || (k1 && m12));
      kept += k2 ? 1 : 0;
      const bool k3 = cnt > 3 && kept < max_keep && !(m03 || (k1 && m13) || (k2 && m23));
      kept += k3 ? 1 : 0;
      if (tid == 0) {
        int q = kept0;
        s_keep[q++] = c0;
        if (k1) s_keep[q++] = c1;
        if (k2) s_keep[q++] = c2;
        if (k3) s_keep[q++] = c3;
      }
      if (kept >= max_keep) break;                     // the boxes kept last suppress nobody that matters
      pos = (cnt > 3 ? c3 : (cnt > 2 ? c2 : (cnt > 1 ? c1 : c0))) + 1;
      if (s0 || (k1 && s1) || (k2 && s2) || (k3 && s3)) my_alive = false;
      {
        const unsigned alive_word = __ballot_sync(0xffffffffu, my_alive);
        if (lane == 0) nxt[tid >> 5] = alive_word;       // this warp owns the word (members before pos are never searched again)
      }
      for (int base = max(act_threads, pos & ~(kYoloThreads - 1)); base < n; base += kYoloThreads) {   // n > 1024 only
        const int j = base + tid;
        bool sj = false;
        const unsigned wj = (j >> 5) < nwords ? cur[j >> 5] : 0u;
        if (j >= pos && j < n && ((wj >> (j & 31)) & 1u)) {
          const float4 bj = s_box[j];
          const float aj = box_area(bj);
          (void)aj;
          sj = iou_above(B0, bj, thr) || (k1 && iou_above(B1, bj, thr)) || (k2 && iou_above(B2, bj, thr)) ||
               (k3 && iou_above(B3, bj, thr));
        }
        const unsigned bsup = __ballot_sync(0xffffffffu, sj);
        if (lane == 0 && (j >> 5) < nwords) nxt[j >> 5] = wj & ~bsup;
      }
      asm volatile("bar.sync 1, %0;" ::"r"(act_threads) : "memory");
      { unsigned* t = cur; cur = nxt; nxt = t; }
    }
  }
  if (tid == 0) s_n = kept;
  __syncthreads();                                     // s_keep, and the global `sorted` rows of this CTA
  if (A.debug) t3 = now();
  kept = s_n;

  // ---- phase 4: YOLOX._postprocess; the kept boxes of a class are sorted by score, so the ones that pass the
  //      score filter are a prefix and the output position is a prefix sum over the classes ----
  bool pass = false;
  float score = 0.f;
  float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
  if (tid < kept) {
    const int i = s_keep[tid];
    score = __uint_as_float((unsigned)(sorted[i] >> 32));
    pass = score > A.cfg.post_score_thresh;
    b = s_box[i];
  }
  const int npass = __syncthreads_count(pass ? 1 : 0);
  if (tid == 0) s_pass = npass;
  cluster_sync_all();
  int offset = 0, total = 0;
  if (tid < 32) {
    const int v = lane < C ? dsmem_read_i32(&s_pass, (unsigned)lane) : 0;
    for (int cc = 0; cc < C; ++cc) {
      const int vv = __shfl_sync(0xffffffffu, v, cc);
      if (cc < c) offset += vv;
      total += vv;
    }
  }
  __shared__ int s_offset;
  if (tid == 0) s_offset = offset;
  __syncthreads();
  offset = s_offset;
  const float inw = (float)A.cfg.in_w, inh = (float)A.cfg.in_h;
  const float imw = (float)A.cfg.img_w, imh = (float)A.cfg.img_h;
  // float32 multiply then divide, truncation toward zero (demo:1009-1012)
  const int x_min = (int)__fdiv_rn(__fmul_rn(fmaxf(0.f, b.x), imw), inw);
  const int y_min = (int)__fdiv_rn(__fmul_rn(fmaxf(0.f, b.y), imh), inh);
  const int x_max = (int)__fdiv_rn(__fmul_rn(fminf(b.z, inw), imw), inw);
  const int y_max = (int)__fdiv_rn(__fmul_rn(fminf(b.w, inh), imh), inh);
  if (pass && A.out && offset + tid < A.max_out) {
    double* o = A.out + (size_t)(offset + tid) * 6;
    o[0] = (double)c; o[1] = (double)score;
    o[2] = (double)x_min; o[3] = (double)y_min; o[4] = (double)x_max; o[5] = (double)y_max;
  }
  if (c == 0) {
    if (tid == 0 && A.out_count) *A.out_count = A.out ? min(total, A.max_out) : total;
    if (A.st_boxes) {
      // the body rows become the tracker's detections of the frame; rows past the last body get score 0 (below
      // track_low_thresh: ignored by every association stage and never born), so the tracker can be stepped
      // with m = max_bodies without the count ever reaching the host
      const int lim = A.out ? min(A.max_bodies, A.max_out) : A.max_bodies;
      const bool body = pass && tid < lim;
      if (tid < A.max_bodies) {
        *reinterpret_cast<int4*>(A.st_boxes + (size_t)tid * 4) = body ? make_int4(x_min, y_min, x_max, y_max) : make_int4(0, 0, 0, 0);
        A.st_scores[tid] = body ? score : 0.f;
      }
      if (tid == 0) *A.st_nbodies = min(npass, lim);
    }
  }
  cluster_sync_all();                                  // nobody leaves while its s_pass may still be read
  if (A.debug && tid == 0) {
    t4 = now();
    printf("yolox class %d: n=%d kept=%d | scan %lld sort %lld nms %lld post %lld ns\n", c, n, kept, t1 - t0, t2 - t1, t3 - t2, t4 - t3);
  }
}

// ------------------------------------------------------------------------------------------------
// ReID crop gather: one CTA = kCropRows output rows of one detection.  Everything that only depends on
// the output row / column (source taps, fixed-point weights) is tabulated in shared memory once per CTA;
// the float64 normalisation is a 3 x 256 table (one value per plane and 8-bit pixel value, built at
// bt_create); the source rows the CTA needs are staged in shared memory as 4-byte BGRx pixels, so a
// bilinear tap is ONE 32-bit shared load for the three channels.  A thread produces four neighbouring
// pixels of a row and writes one float4 per channel plane.
// ------------------------------------------------------------------------------------------------
constexpr int kCropRows = 16;
constexpr int kCropStagePixels = 4096;      // 16 KB of staged source pixels per CTA; larger crops read the frame directly

// OpenCV resize (INTER_LINEAR, 8U): scale = 1 / (dst / src) in double
__device__ __forceinline__ double resize_scale(int src, int dst) {
  const double inv = (double)dst / (double)src;
  return 1.0 / inv;
}
__device__ __forceinline__ void resize_coord(int d, double scale, int src, bool is_x, int* s0, int* s1, int* a0, int* a1) {
  // f in float32
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

__device__ __forceinline__ uint32_t load_bgr(const uint8_t* __restrict__ p) {
  return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16);
}

// dynamic shared memory: int4 xtab[out_w] | int4 ytab[kCropRows] | u32 stage[kCropStagePixels]
__global__ void __launch_bounds__(256, 6)
reid_crop_kernel(const uint8_t* __restrict__ frame, int h, int w, const int32_t* __restrict__ boxes, int out_h,
                 int out_w, float* __restrict__ out, const int32_t* __restrict__ n_dev, const float* __restrict__ lut) {
  __shared__ float s_lut[3][256];           // [rgb plane][pixel value] -> (v / 255 - mean) / std
  extern __shared__ __align__(16) unsigned char s_dyn[];
  int4* s_x = reinterpret_cast<int4*>(s_dyn);
  int4* s_y = s_x + out_w;
  uint32_t* s_src = reinterpret_cast<uint32_t*>(s_y + kCropRows);
  const int det = blockIdx.y;
  if (n_dev && det >= *n_dev) return;       // device-chained use: the number of bodies never visits the host
  const int tid = threadIdx.x;
  const int row0 = blockIdx.x * kCropRows;
  const int rows = min(kCropRows, out_h - row0);
  const int4 b = *reinterpret_cast<const int4*>(boxes + (size_t)det * 4);
  // numpy slicing image[y1:y2, x1:x2] clamps to the image
  const int x1 = min(max(b.x, 0), w), y1 = min(max(b.y, 0), h);
  const int x2 = min(max(b.z, 0), w), y2 = min(max(b.w, 0), h);
  const int sw = x2 - x1, sh = y2 - y1;
  const size_t plane = (size_t)out_h * out_w;
  float* o = out + (size_t)det * 3 * plane;
  const bool vec = (out_w & 3) == 0;
  const int groups = (out_w + 3) >> 2;
  if (sw <= 0 || sh <= 0) {                 // empty crop: the reference feeds zeros
    for (int it = tid; it < rows * groups * 3; it += 256) {
      const int cpl = it / (rows * groups), rem = it % (rows * groups);
      const int r = rem / groups, g = rem % groups;
      float* dst = o + (size_t)cpl * plane + (size_t)(row0 + r) * out_w + g * 4;
      if (vec) *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
      else for (int e = 0; e < 4 && g * 4 + e < out_w; ++e) dst[e] = 0.f;
    }
    return;
  }
#pragma unroll
  for (int rgb = 0; rgb < 3; ++rgb) s_lut[rgb][tid] = lut[rgb * 256 + tid];
  __shared__ double s_scale[2];
  if (tid < 2) s_scale[tid] = tid ? resize_scale(sh, out_h) : resize_scale(sw, out_w);   // the float64 divisions: two threads
  __syncthreads();
  for (int d = tid; d < out_w + rows; d += 256) {
    const bool is_x = d < out_w;
    int s0, s1, a0, a1;
    resize_coord(is_x ? d : row0 + d - out_w, s_scale[is_x ? 0 : 1], is_x ? sw : sh, is_x, &s0, &s1, &a0, &a1);
    s_x[d] = make_int4(s0, s1, a0, a1);       // s_y follows s_x
  }
  __syncthreads();
  // source rows of this CTA: the taps are monotone in the output row
  const int ylo = s_y[0].x, yhi = s_y[rows - 1].y;
  const int nsrc = yhi - ylo + 1;
  const bool staged = nsrc * sw <= kCropStagePixels;
  if (staged) {
    for (int idx = tid; idx < nsrc * sw; idx += 256) {
      const int yy = idx / sw, xx = idx - yy * sw;
      s_src[idx] = load_bgr(frame + ((size_t)(y1 + ylo + yy) * w + x1 + xx) * 3);
    }
    __syncthreads();
  }
  // One output pixel: the four taps as BGRx words; per channel the two horizontal sums are one byte-permute + one
  // 2-way dot product each (weights packed as two 16-bit halves), the vertical step two high-multiplies (row weight
  // pre-shifted by 16): ((by * (h >> 4)) >> 16) == umulhi(h >> 4, by << 16).
  auto pixel = [&](uint32_t p00, uint32_t p01, uint32_t p10, uint32_t p11, uint32_t wx, uint32_t by0s, uint32_t by1s,
                   float* vb, float* vg, float* vr) {
#pragma unroll
    for (int cc = 0; cc < 3; ++cc) {        // cc = BGR channel of the source
      const uint32_t sel = 0x0040u + 0x0011u * cc;                       // bytes: [p0x.cc, p1x.cc, 0, 0]
      const uint32_t h0 = __dp2a_lo(wx, __byte_perm(p00, p01, sel), 0u);
      const uint32_t h1 = __dp2a_lo(wx, __byte_perm(p10, p11, sel), 0u);
      const uint32_t pv = (__umulhi(h0 >> 4, by0s) + __umulhi(h1 >> 4, by1s) + 2u) >> 2;
      const float val = s_lut[2 - cc][pv];  // [..., ::-1]: BGR -> RGB
      if (cc == 0) *vb = val; else if (cc == 1) *vg = val; else *vr = val;
    }
  };
  for (int it = tid; it < rows * groups; it += 256) {
    const int r = it / groups, g = it - r * groups;
    const int4 ty = s_y[r];
    const uint32_t by0s = (uint32_t)ty.z << 16, by1s = (uint32_t)ty.w << 16;
    float v[3][4];
    if (staged) {
      const uint32_t* q0 = s_src + (ty.x - ylo) * sw;
      const uint32_t* q1 = s_src + (ty.y - ylo) * sw;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int4 tx = s_x[min(g * 4 + e, out_w - 1)];
        pixel(q0[tx.x], q0[tx.y], q1[tx.x], q1[tx.y], (uint32_t)tx.z | ((uint32_t)tx.w << 16), by0s, by1s,
              &v[2][e], &v[1][e], &v[0][e]);
      }
    } else {
      const uint8_t* r0 = frame + ((size_t)(y1 + ty.x) * w + x1) * 3;
      const uint8_t* r1 = frame + ((size_t)(y1 + ty.y) * w + x1) * 3;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int4 tx = s_x[min(g * 4 + e, out_w - 1)];
        pixel(load_bgr(r0 + tx.x * 3), load_bgr(r0 + tx.y * 3), load_bgr(r1 + tx.x * 3), load_bgr(r1 + tx.y * 3),
              (uint32_t)tx.z | ((uint32_t)tx.w << 16), by0s, by1s, &v[2][e], &v[1][e], &v[0][e]);
      }
    }
#pragma unroll
    for (int rgb = 0; rgb < 3; ++rgb) {
      float* dst = o + (size_t)rgb * plane + (size_t)(row0 + r) * out_w + g * 4;
      if (vec) *reinterpret_cast<float4*>(dst) = make_float4(v[rgb][0], v[rgb][1], v[rgb][2], v[rgb][3]);
      else for (int e = 0; e < 4 && g * 4 + e < out_w; ++e) dst[e] = v[rgb][e];
    }
  }
}

int32_t launch_yolox(bt_ctx* ctx, YoloArgs A) {
  const bt_yolox_config& cfg = A.cfg;
  BT_CHECK(cfg.num_classes >= 1 && cfg.num_classes <= kMaxClasses, BT_ERR_INVALID, "num_classes must be 1..%d",
           kMaxClasses);
  BT_CHECK(cfg.max_per_class >= 1 && cfg.max_per_class <= kKeepCap, BT_ERR_INVALID, "max_per_class must be 1..%d",
           kKeepCap);
  BT_CHECK(cfg.in_h % 32 == 0 && cfg.in_w % 32 == 0 && cfg.in_h > 0 && cfg.in_w > 0, BT_ERR_INVALID,
           "input size must be a positive multiple of 32");
  const int anchors = (cfg.in_h / 8) * (cfg.in_w / 8) + (cfg.in_h / 16) * (cfg.in_w / 16) +
                      (cfg.in_h / 32) * (cfg.in_w / 32);
  const size_t smem = (((size_t)anchors + 2) & ~size_t(1)) * 8 + (size_t)anchors * 16 + (size_t)((anchors + 31) / 32) * 8;
  BT_CHECK(smem <= 220 * 1024, BT_ERR_CAPACITY, "too many anchors (%d) for the per-class CTA's shared memory", anchors);
  A.anchors = anchors;
  static const bool dbg = getenv("BT_YOLOX_DEBUG") != nullptr;
  A.debug = dbg;
  BT_TRY(bt_arena(ctx, (size_t)cfg.num_classes * anchors, &A.sorted));
  BT_CUDA(cudaFuncSetAttribute(yolox_post_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t lc = {};
  lc.gridDim = dim3(cfg.num_classes); lc.blockDim = dim3(kYoloThreads); lc.dynamicSmemBytes = smem; lc.stream = ctx->stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cfg.num_classes; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  lc.attrs = at; lc.numAttrs = 1;
  BT_CUDA(cudaLaunchKernelEx(&lc, yolox_post_kernel, A));
  BT_LAUNCHED(ctx);
  return BT_OK;
}

}  // namespace

int32_t btk_yolox_postprocess(bt_ctx* ctx, const float* raw, const bt_yolox_config& cfg, double* out_boxes,
                              int32_t max_out, int32_t* out_count) {
  YoloArgs A = {};
  A.raw = raw; A.cfg = cfg; A.out = out_boxes; A.max_out = max_out; A.out_count = out_count;
  return launch_yolox(ctx, A);
}

static int32_t launch_crop(bt_ctx* ctx, const uint8_t* frame, int32_t h, int32_t w, const int32_t* boxes, int32_t n,
                           int32_t out_h, int32_t out_w, float* out, const int32_t* n_dev) {
  BT_CHECK(out_w <= 2048, BT_ERR_CAPACITY, "crop width %d exceeds 2048", out_w);
  dim3 grid((out_h + kCropRows - 1) / kCropRows, n);
  const size_t smem = (size_t)(out_w + kCropRows) * sizeof(int4) + (size_t)kCropStagePixels * 4;
  static std::once_flag once[64];          // function attributes are per device
  std::call_once(once[ctx->device & 63], [] { cudaFuncSetAttribute(reid_crop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024); });
  reid_crop_kernel<<<grid, 256, smem, ctx->stream>>>(frame, h, w, boxes, out_h, out_w, out, n_dev,
                                                     reinterpret_cast<const float*>(ctx->d_desc + kBtCropLutOffset));
  BT_LAUNCHED(ctx);
  return BT_OK;
}

int32_t btk_reid_crop_gather(bt_ctx* ctx, const uint8_t* frame, int32_t h, int32_t w, const int32_t* boxes,
                             int32_t n, int32_t out_h, int32_t out_w, float* out) {
  if (n <= 0) return BT_OK;
  return launch_crop(ctx, frame, h, w, boxes, n, out_h, out_w, out, nullptr);
}



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
  size_t need = (size_t)anchors * 8 * kMaxClasses + 4096;
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
  BT_TRY(bt_arena_reserve(ctx, (size_t)anchors * 8 * kMaxClasses + 4096));
  int32_t* d_nb = nullptr;
  BT_TRY(bt_arena(ctx, (size_t)4, &d_nb));
  YoloArgs A = {};
  A.raw = raw_head; A.cfg = *cfg; A.out = det_out; A.max_out = list_cap; A.out_count = det_count;
  A.st_boxes = in_boxes; A.st_scores = in_scores; A.st_nbodies = d_nb; A.max_bodies = max_bodies;
  BT_TRY(launch_yolox(ctx, A));
  BT_TRY(launch_crop(ctx, frame, h, w, in_boxes, max_bodies, out_h, out_w, crops, d_nb));
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
