// The stateful tracker: one BoTSORT.update (demo:1291-1639; demo =
// /root/reference/demo_bottrack_onnx_tflite.py) per bt_update_arrays call, on a device-resident
// track store.
//
// Device side (all arithmetic): per-slot Kalman state (fp64 AoS), cached tlbr, fp16/fp32 feature
// banks; per frame: detection prep -> batched Kalman predict -> ONE fused association kernel over
// (all live slots) x (all detections) that emits the candidate edges of the three association
// stages at once -> three exact LAP solves chained on the device (stage 2 masked by stage 1's
// unmatched rows, stage 3 by stage 1's unmatched columns) -> batched Kalman update / initiate /
// feature EMA -> sparse duplicate test (tracked x lost, IoU distance < 0.15).
//
// Host side (this file, C++): only the list bookkeeping of demo:1414-1423 and demo:1558-1639
// (who is tracked / lost / removed, ids, list order) on small index arrays, with two stream
// synchronisations per frame.  SURVEY.md section 8(f) F1 lists moving this bookkeeping to the
// device as the next step.
#include "common.cuh"

#include <algorithm>
#include <chrono>
#include <stdlib.h>
#include <string.h>

namespace {

constexpr int kPairPrefetch = 512;   // duplicate pairs fetched with the first read-back (more: second copy)

struct SlotMeta {
  int32_t state = BT_STATE_NEW;
  int32_t activated = 0;
  int32_t track_id = 0;
  int32_t frame_id = 0;
  int32_t start_frame = 0;
  int32_t tracklet_len = 0;
  float score = 0.f;
  int32_t det_index = -1;
  uint8_t in_removed = 0;  // id is in self.removed_stracks (demo:1636)
  uint8_t f32_state = 0;   // mean/cov are still initiate()'s float32 values (NumPy >= 2 quirk)
  uint8_t used = 0;
  uint8_t mark = 0;        // scratch flag
};

// boxes int32 tlbr -> tlbr float64, xywh float64 (Kalman measurement, demo:599 + demo:664-670),
// xywh float32 (initiate input, demo:561) and the score class of the detection
// (demo:1501: high = score > 0.40; demo:1531: low = 0.1 <= score <= 0.40).
__global__ void det_prep_kernel(const int32_t* __restrict__ boxes, const float* __restrict__ scores, int m,
                                float high, float low, double* __restrict__ tlbr, double* __restrict__ xywh,
                                float* __restrict__ xywh32, uint8_t* __restrict__ kind, uint2* __restrict__ pk,
                                float* __restrict__ res_scores, int32_t* __restrict__ res_boxes) {
  bt_grid_launch_dependents();   // feature_prep does not read anything written here: let it start right away
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m) return;
  const int4 b = *reinterpret_cast<const int4*>(boxes + (size_t)j * 4);
  // STrack(tlbr_to_tlwh(int box)) -> float32 tlwh (demo:465); xywh = tl + wh/2 in float32 (exact)
  const float x1 = (float)b.x, y1 = (float)b.y, w = (float)(b.z - b.x), h = (float)(b.w - b.y);
  const float cx = x1 + w / 2, cy = y1 + h / 2;
  double* t = tlbr + (size_t)j * 4;
  // detection tlbr = tlwh (float32) with wh += tl (demo:643-648)
  t[0] = x1; t[1] = y1; t[2] = (double)(w + x1); t[3] = (double)(h + y1);
  double* z = xywh + (size_t)j * 4;
  z[0] = cx; z[1] = cy; z[2] = w; z[3] = h;
  *reinterpret_cast<float4*>(xywh32 + (size_t)j * 4) = make_float4(cx, cy, w, h);
  const float s = scores[j];
  if (res_scores) {   // device-resident inputs: the host reads them back with the frame's result block
    res_scores[j] = s;
    *reinterpret_cast<int4*>(res_boxes + (size_t)j * 4) = b;
  }
  kind[j] = (s > high) ? BT_COL_HIGH : ((s >= low) ? BT_COL_LOW : BT_COL_NONE);
  // packed integer corners for the association kernel's overlap screen, from the same float32 values
  pk[j] = bt_pack16_f32(x1, y1, w + x1, h + y1, true);
}

// debugging aid (BT_DEBUG_DELAY_SIDE=<us>): holds the side stream back so that a missing
// dependency between the two streams shows up as a wrong result instead of a rare flake
__global__ void debug_spin_kernel(long long ns) {
  unsigned long long t0, t1;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  do { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1)); } while ((long long)(t1 - t0) < ns);
}

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

}  // namespace

struct bt_tracker {
  bt_config cfg;
  int cap = 0, max_dets = 0, D = 0;
  int max_time_lost = 0;
  // ---- device track store (indexed by slot) ----
  double *mean = nullptr, *cov = nullptr, *tlbr = nullptr;
  float* tlbr_f32 = nullptr;
  __half* feat16 = nullptr;
  float *curr32 = nullptr, *smooth32 = nullptr;
  uint8_t* row_kind = nullptr;
  uint8_t* row_kind_cur = nullptr;   // this frame's row kinds inside the packed control block
  uint8_t* slot_f32 = nullptr;       // [cap] slot still holds initiate()'s float32 state (NumPy >= 2 quirk)
  // ---- device per-frame buffers ----
  int32_t* det_boxes = nullptr;
  float* det_scores = nullptr;
  float* det_feat_in = nullptr;   // staging of host features
  float* det_feat32 = nullptr;    // normalised
  __half* det_feat16 = nullptr;
  double *det_tlbr = nullptr, *det_xywh = nullptr;
  float* det_xywh32 = nullptr;
  uint8_t* col_kind = nullptr;
  uint2* col_pk = nullptr;     // packed integer corners of the detections (association overlap screen)
  int32_t *x[3] = {nullptr, nullptr, nullptr}, *y[3] = {nullptr, nullptr, nullptr};
  int32_t *d_pool_idx = nullptr, *d_pool_state = nullptr;
  int32_t *d_upd_track = nullptr, *d_upd_det = nullptr;
  uint8_t *d_upd_f32 = nullptr, *d_ema_mode = nullptr;
  int32_t *d_birth_slot = nullptr, *d_birth_det = nullptr;
  int32_t *d_lista = nullptr, *d_listb = nullptr;
  int32_t *d_pairs = nullptr, *d_pair_count = nullptr;
  char* d_ctrl = nullptr;   // packed per-frame control lists (one H2D per phase)
  char* h_ctrl = nullptr;
  cudaStream_t st2 = nullptr;   // side stream: work that is independent of the main chain runs beside it
  cudaEvent_t ev_fork1 = nullptr, ev_join1 = nullptr, ev_fork2 = nullptr, ev_join2 = nullptr, ev_x = nullptr;
  bool overlap = true;
  bool predict_side = false;  // BT_PREDICT_SIDE=1: control block + Kalman predict on the side stream (measured: the four
                              // event calls cost the host more than the 5 us kernel costs the main stream)
  int debug_delay_us = 0;
  bool host_debug = false;
  char* d_res = nullptr;    // packed per-frame result block (one D2H per frame), layout in bt_update_arrays
  char* h_res = nullptr;
  size_t res_cap = 0;
  double* d_gather = nullptr;
  int pair_cap = 0;
  // ---- pinned host mirrors ----
  char* pinned = nullptr;
  uint8_t* h_row_kind = nullptr;
  int32_t *h_pool_idx = nullptr, *h_pool_state = nullptr;
  int32_t *h_x[3] = {nullptr, nullptr, nullptr}, *h_y[3] = {nullptr, nullptr, nullptr};
  float* h_scores = nullptr;      // this frame's host view of the scores
  float* h_scores_own = nullptr;  // pinned copy used with host inputs
  int32_t *h_upd_track = nullptr, *h_upd_det = nullptr;
  uint8_t *h_upd_f32 = nullptr, *h_ema_mode = nullptr;
  int32_t *h_birth_slot = nullptr, *h_birth_det = nullptr;
  int32_t *h_lista = nullptr, *h_listb = nullptr;
  int32_t *h_pairs = nullptr, *h_pair_count = nullptr;
  double* h_tlbr = nullptr;
  // ---- host bookkeeping ----
  std::vector<SlotMeta> meta;
  std::vector<int> tracked, lost;
  std::vector<int> free_slots;  // kept sorted descending: back() is the lowest free slot
  int high_water = 0;
  int id_count = 0;
  int frame_id = 0;
  int n_removed_total = 0;
  std::vector<int32_t> matches[3];  // flattened (a, b) pairs in the reference's index spaces
  bt_assoc_params last_assoc;          // for bt_profile_replay_assoc
  int last_assoc_precision = 0;
  bool last_assoc_valid = false;
  std::vector<uint8_t> scratch_a, scratch_b;
  std::vector<int> v_unconfirmed, v_pool, v_hi_pos, v_lo_pos, v_hi_list, v_lo_list, v_activated, v_refind, v_lost_now,
      v_removed_now, v_r_tracked, v_u_det_pos, v_new_tracked, v_new_lost, v_tmp;
  std::vector<uint8_t> v_det_taken, v_pool_matched;
  std::vector<int> scratch_pos_t, scratch_pos_l;
  int32_t *d_bpairs = nullptr, *h_bpairs = nullptr, *h_boxes = nullptr;
  int bpair_cap = 1 << 16;
  std::vector<double> tlbr_cache;   // tlbr of `tracked` after the frame
  bool tlbr_cache_valid = false;
  // ---- optional segment timing (bt_profile_*) ----
  bool prof = false;
  cudaEvent_t ev[BT_SEG_COUNT][2] = {};
  bool seg_open[BT_SEG_COUNT] = {};
  double prof_ms[BT_SEG_COUNT] = {};
  int64_t prof_n[BT_SEG_COUNT] = {};
};

#define SEG_BEGIN(s)                                                  \
  do {                                                                \
    if (t->prof) { BT_CUDA(cudaEventRecord(t->ev[s][0], st)); }       \
  } while (0)
#define SEG_END(s)                                                    \
  do {                                                                \
    if (t->prof) { BT_CUDA(cudaEventRecord(t->ev[s][1], st)); t->seg_open[s] = true; } \
  } while (0)

static inline double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
#define HOST_MARK(seg)                                                \
  do {                                                                \
    if (t->prof || t->host_debug) {                                   \
      const double n__ = now_ms();                                    \
      if (t->prof) { t->prof_ms[seg] += n__ - t_host; t->prof_n[seg] += 1; } \
      hphase[seg - BT_SEG_HOST_ENQUEUE1] = n__ - t_host;              \
      t_host = n__;                                                   \
    }                                                                 \
  } while (0)

static int32_t prof_collect(bt_ctx* ctx, bt_tracker* t) {
  if (!t->prof) return BT_OK;
  for (int s = 0; s < BT_SEG_HOST_ENQUEUE1; ++s) {
    if (!t->seg_open[s]) continue;
    float ms = 0.f;
    BT_CUDA(cudaEventElapsedTime(&ms, t->ev[s][0], t->ev[s][1]));
    t->prof_ms[s] += ms;
    t->prof_n[s] += 1;
    t->seg_open[s] = false;
  }
  return BT_OK;
}

namespace {

template <typename T>
int32_t dev_alloc(bt_ctx* ctx, T** p, size_t count) {
  BT_CUDA(cudaMalloc(p, sizeof(T) * (count ? count : 1)));
  return BT_OK;
}

template <typename T>
T* carve(char*& cur, size_t count) {
  T* p = reinterpret_cast<T*>(cur);
  cur += (sizeof(T) * count + 255) & ~size_t(255);
  return p;
}

int alloc_slot(bt_tracker* t) {
  if (!t->free_slots.empty()) {
    const int s = t->free_slots.back();
    t->free_slots.pop_back();
    return s;
  }
  if (t->high_water < t->cap) return t->high_water++;
  return -1;
}

}  // namespace

int32_t bt_tracker_create(bt_ctx* ctx) {
  auto* t = new bt_tracker();
  ctx->trk = t;
  bt_default_config(&t->cfg);
  t->cap = ctx->max_tracks;
  t->max_dets = ctx->max_dets;
  t->D = ctx->feat_dim;
  const size_t cap = t->cap, md = t->max_dets, D = t->D;
  const bool keep32 = !(ctx->flags & BT_FLAG_NO_F32_FEATURES);
  BT_TRY(dev_alloc(ctx, &t->mean, cap * 8));
  BT_TRY(dev_alloc(ctx, &t->cov, cap * 64));
  BT_TRY(dev_alloc(ctx, &t->tlbr, cap * 4));
  BT_TRY(dev_alloc(ctx, &t->tlbr_f32, cap * 4));
  BT_TRY(dev_alloc(ctx, &t->feat16, cap * D));
  BT_CUDA(cudaMemset(t->feat16, 0, sizeof(__half) * cap * D));
  BT_CUDA(cudaMemset(t->tlbr, 0, sizeof(double) * cap * 4));
  BT_CUDA(cudaMemset(t->tlbr_f32, 0, sizeof(float) * cap * 4));
  if (keep32) {
    BT_TRY(dev_alloc(ctx, &t->curr32, cap * D));
    BT_TRY(dev_alloc(ctx, &t->smooth32, cap * D));
    BT_CUDA(cudaMemset(t->curr32, 0, sizeof(float) * cap * D));
    BT_CUDA(cudaMemset(t->smooth32, 0, sizeof(float) * cap * D));
  }
  BT_TRY(dev_alloc(ctx, &t->row_kind, cap));
  BT_TRY(dev_alloc(ctx, &t->slot_f32, cap));
  BT_CUDA(cudaMemset(t->slot_f32, 0, cap));
  BT_TRY(dev_alloc(ctx, &t->det_boxes, md * 4));
  BT_TRY(dev_alloc(ctx, &t->det_scores, md));
  BT_TRY(dev_alloc(ctx, &t->det_feat_in, md * D));
  BT_TRY(dev_alloc(ctx, &t->det_feat32, md * D));
  BT_TRY(dev_alloc(ctx, &t->det_feat16, md * D));
  BT_TRY(dev_alloc(ctx, &t->det_tlbr, md * 4));
  BT_TRY(dev_alloc(ctx, &t->det_xywh, md * 4));
  BT_TRY(dev_alloc(ctx, &t->det_xywh32, md * 4));
  BT_TRY(dev_alloc(ctx, &t->col_kind, md));
  BT_TRY(dev_alloc(ctx, &t->col_pk, md));
  // x[0..2] point into the per-frame result block (set in bt_update_arrays)
  for (int s = 0; s < 3; ++s) BT_TRY(dev_alloc(ctx, &t->y[s], md));
  BT_TRY(dev_alloc(ctx, &t->d_ctrl, 16 * (cap + md) + 1024));
  t->res_cap = 8 + 8 * (size_t)kPairPrefetch + 12 * cap + 20 * md + 32 * cap + 512;
  BT_TRY(dev_alloc(ctx, &t->d_res, t->res_cap));
  const size_t nupd = cap + md;
  BT_TRY(dev_alloc(ctx, &t->d_pool_idx, cap));
  BT_TRY(dev_alloc(ctx, &t->d_pool_state, cap));
  BT_TRY(dev_alloc(ctx, &t->d_upd_track, nupd));
  BT_TRY(dev_alloc(ctx, &t->d_upd_det, nupd));
  BT_TRY(dev_alloc(ctx, &t->d_upd_f32, nupd));
  BT_TRY(dev_alloc(ctx, &t->d_ema_mode, nupd));
  BT_TRY(dev_alloc(ctx, &t->d_birth_slot, md));
  BT_TRY(dev_alloc(ctx, &t->d_birth_det, md));
  BT_TRY(dev_alloc(ctx, &t->d_lista, cap));
  BT_TRY(dev_alloc(ctx, &t->d_listb, cap));
  t->pair_cap = 1 << 20;
  BT_TRY(dev_alloc(ctx, &t->d_pairs, (size_t)2 * t->pair_cap));
  BT_TRY(dev_alloc(ctx, &t->d_pair_count, 1));
  BT_TRY(dev_alloc(ctx, &t->d_gather, cap * 64));
  BT_TRY(dev_alloc(ctx, &t->d_bpairs, (size_t)2 * t->bpair_cap));

  size_t pinned_bytes = 0;
  pinned_bytes += t->res_cap + 256 + (size_t)8 * t->bpair_cap + md * 16 + 512 + 16 * (cap + md) + 1024 + cap + 2 * cap * 4 + 3 * (cap + md) * 4 + md * 4 + 2 * nupd * 4 + 2 * nupd + 2 * md * 4 +
                  2 * cap * 4 + (size_t)2 * t->pair_cap * 4 + 64 + cap * 4 * 8 + 64 * 256;
  BT_CUDA(cudaMallocHost(&t->pinned, pinned_bytes));
  char* cur = t->pinned;
  t->h_row_kind = carve<uint8_t>(cur, cap);
  t->h_pool_idx = carve<int32_t>(cur, cap);
  t->h_pool_state = carve<int32_t>(cur, cap);
  t->h_x[0] = carve<int32_t>(cur, 3 * cap);
  t->h_x[1] = t->h_x[0] + cap;
  t->h_x[2] = t->h_x[0] + 2 * cap;
  for (int s = 0; s < 3; ++s) t->h_y[s] = carve<int32_t>(cur, md);
  t->h_ctrl = carve<char>(cur, 16 * (cap + md) + 1024);
  t->h_res = carve<char>(cur, t->res_cap);
  t->h_scores_own = carve<float>(cur, md);
  t->h_scores = t->h_scores_own;
  t->h_upd_track = carve<int32_t>(cur, nupd);
  t->h_upd_det = carve<int32_t>(cur, nupd);
  t->h_upd_f32 = carve<uint8_t>(cur, nupd);
  t->h_ema_mode = carve<uint8_t>(cur, nupd);
  t->h_birth_slot = carve<int32_t>(cur, md);
  t->h_birth_det = carve<int32_t>(cur, md);
  t->h_lista = carve<int32_t>(cur, cap);
  t->h_listb = carve<int32_t>(cur, cap);
  t->h_pairs = carve<int32_t>(cur, (size_t)2 * t->pair_cap);
  t->h_pair_count = carve<int32_t>(cur, 16);
  t->h_tlbr = carve<double>(cur, cap * 4);
  t->h_bpairs = carve<int32_t>(cur, (size_t)2 * t->bpair_cap);
  t->h_boxes = carve<int32_t>(cur, md * 4);
  t->meta.assign(cap, SlotMeta());
  t->max_time_lost = (int)(t->cfg.frame_rate / 30.0 * t->cfg.track_buffer);
  BT_CUDA(cudaStreamCreateWithFlags(&t->st2, cudaStreamNonBlocking));
  for (cudaEvent_t* e : {&t->ev_fork1, &t->ev_join1, &t->ev_fork2, &t->ev_join2, &t->ev_x})
    BT_CUDA(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
  t->overlap = getenv("BT_NO_OVERLAP") == nullptr;
  t->predict_side = getenv("BT_PREDICT_SIDE") != nullptr;
  t->debug_delay_us = getenv("BT_DEBUG_DELAY_SIDE") ? atoi(getenv("BT_DEBUG_DELAY_SIDE")) : 0;
  t->host_debug = getenv("BT_HOST_DEBUG") != nullptr;
  return BT_OK;
}

void bt_tracker_destroy(bt_ctx* ctx) {
  bt_tracker* t = ctx->trk;
  if (!t) return;
  void* ptrs[] = {t->mean, t->cov, t->tlbr, t->tlbr_f32, t->feat16, t->curr32, t->smooth32, t->row_kind, t->slot_f32,
                  t->det_boxes, t->det_scores, t->det_feat_in, t->det_feat32, t->det_feat16, t->det_tlbr,
                  t->det_xywh, t->det_xywh32, t->col_kind, t->col_pk, t->d_ctrl, t->y[0], t->y[1], t->y[2],
                  t->d_pool_idx, t->d_pool_state, t->d_upd_track, t->d_upd_det, t->d_upd_f32, t->d_ema_mode,
                  t->d_birth_slot, t->d_birth_det, t->d_lista, t->d_listb, t->d_pairs, t->d_pair_count,
                  t->d_gather, t->d_bpairs, t->d_res};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  if (t->pinned) cudaFreeHost(t->pinned);
  for (cudaEvent_t e : {t->ev_fork1, t->ev_join1, t->ev_fork2, t->ev_join2, t->ev_x})
    if (e) cudaEventDestroy(e);
  if (t->st2) cudaStreamDestroy(t->st2);
  for (int s = 0; s < BT_SEG_HOST_ENQUEUE1; ++s)
    for (cudaEvent_t e : t->ev[s])
      if (e) cudaEventDestroy(e);
  delete t;
  ctx->trk = nullptr;
}

extern "C" {

int32_t bt_tracker_reset(bt_ctx* ctx, const bt_config* cfg) {
  if (!ctx) return BT_ERR_INVALID;
  BT_CUDA(cudaSetDevice(ctx->device));
  bt_tracker* t = ctx->trk;
  BT_CUDA(cudaStreamSynchronize(ctx->stream));
  if (cfg) t->cfg = *cfg;
  else bt_default_config(&t->cfg);
  t->max_time_lost = (int)(t->cfg.frame_rate / 30.0 * t->cfg.track_buffer);  // demo:1276-1277
  t->meta.assign(t->cap, SlotMeta());
  t->tracked.clear();
  t->lost.clear();
  t->free_slots.clear();
  t->high_water = 0;
  t->id_count = 0;  // BaseTrack.clear_count(), demo:1264 (per tracker here)
  t->frame_id = 0;
  t->n_removed_total = 0;
  for (auto& m : t->matches) m.clear();
  t->tlbr_cache_valid = false;
  BT_CUDA(cudaMemsetAsync(t->feat16, 0, sizeof(__half) * (size_t)t->cap * t->D, ctx->stream));
  const bt_cand& cand = *bt_lap_own_cand(ctx);
  BT_CUDA(cudaMemsetAsync(cand.cnt, 0, cand.clear_bytes, ctx->stream));
  return BT_OK;
}

int32_t bt_update_arrays(bt_ctx* ctx, const int32_t* boxes, const float* scores, const float* feats, int32_t m,
                         int32_t loc, bt_frame_info* info) {
  if (!ctx) return BT_ERR_INVALID;
  BT_CUDA(cudaSetDevice(ctx->device));
  bt_tracker* t = ctx->trk;
  const bt_config& cfg = t->cfg;
  BT_CHECK(loc == BT_HOST || loc == BT_DEVICE, BT_ERR_INVALID, "bad loc");
  BT_CHECK(m >= 0 && m <= t->max_dets, BT_ERR_CAPACITY, "%d detections exceed ctx max_dets %d", m, t->max_dets);
  BT_CHECK(m == 0 || (boxes && scores), BT_ERR_INVALID, "NULL boxes/scores");
  const bool reid = cfg.with_reid != 0;
  BT_CHECK(!reid || m == 0 || feats, BT_ERR_INVALID, "with_reid is set but feats is NULL");
  const int D = t->D;
  const bool tensor_path = !(ctx->flags & BT_FLAG_SIMT_SIM) && (D % 64 == 0);
  const bool keep32 = t->curr32 != nullptr;
  cudaStream_t st = ctx->stream;
  std::vector<SlotMeta>& meta = t->meta;

  double t_host = now_ms();
  double hphase[5] = {0, 0, 0, 0, 0};
  t->frame_id += 1;  // demo:1292
  const int frame_id = t->frame_id;
  t->tlbr_cache_valid = false;

  // ---- inputs -> device ---------------------------------------------------------------------
  // ---- per-frame result block layout (device d_res mirrors pinned h_res) ----
  //   int32 [0..1]: duplicate-pair count | int32 pairs[2*kPairPrefetch] | x1,x2,x3 [n_rows each]
  //   | (device inputs only) float scores[m], int32 boxes[4m] | pad to 8 B | double tlbr[4*n_rows]
  const int n_rows_pre = t->high_water;
  const bool inputs_on_device = (loc == BT_DEVICE);
  //   part A (read back right after the LAP, the host starts its list bookkeeping on it while the GPU
  //           still runs update / EMA / duplicate test): x1,x2,x3 [n_rows each] | scores[m], boxes[4m]
  //   part B (read back at the end): pair count (2 ints) | pairs[2*kPairPrefetch] | tlbr[4*n_rows] f64
  const size_t o_x = 0, o_sc = o_x + 3 * (size_t)n_rows_pre,
               o_bx = (o_sc + (inputs_on_device ? m : 0) + 3) & ~size_t(3),   // int4 stores: 16 B aligned
               o_hdr = (o_bx + (inputs_on_device ? 4 * (size_t)m : 0) + 3) & ~size_t(3),
               o_pairs = o_hdr + 2, o_end_i = (o_pairs + 2 * (size_t)kPairPrefetch + 1) & ~size_t(1);
  const size_t bytesA_res = sizeof(int32_t) * o_hdr;
  const size_t bytesB_res = sizeof(int32_t) * (o_end_i - o_hdr) + sizeof(double) * 4 * n_rows_pre;
  int32_t* dres_i = reinterpret_cast<int32_t*>(t->d_res);
  int32_t* hres_i = reinterpret_cast<int32_t*>(t->h_res);
  for (int s3 = 0; s3 < 3; ++s3) {
    t->x[s3] = dres_i + o_x + (size_t)s3 * n_rows_pre;
    t->h_x[s3] = hres_i + o_x + (size_t)s3 * n_rows_pre;
  }
  double* dres_tlbr = reinterpret_cast<double*>(dres_i + o_end_i);
  const double* hres_tlbr = reinterpret_cast<const double*>(hres_i + o_end_i);
  if (inputs_on_device) t->h_scores = reinterpret_cast<float*>(hres_i + o_sc);
  else t->h_scores = t->h_scores_own;
  const int32_t* host_boxes = inputs_on_device ? (hres_i + o_bx) : boxes;
  const int32_t* d_boxes = boxes;
  const float* d_scores = scores;
  const float* d_feats = feats;
  SEG_BEGIN(BT_SEG_PREP);
  if (m > 0) {
    if (loc == BT_HOST) {
      BT_CUDA(cudaMemcpyAsync(t->det_boxes, boxes, sizeof(int32_t) * 4 * m, cudaMemcpyHostToDevice, st));
      BT_CUDA(cudaMemcpyAsync(t->det_scores, scores, sizeof(float) * m, cudaMemcpyHostToDevice, st));
      d_boxes = t->det_boxes;
      d_scores = t->det_scores;
      if (reid) {
        BT_CUDA(cudaMemcpyAsync(t->det_feat_in, feats, sizeof(float) * (size_t)m * D, cudaMemcpyHostToDevice, st));
        d_feats = t->det_feat_in;
      }
      memcpy(t->h_scores_own, scores, sizeof(float) * m);
    } else {
    }
    det_prep_kernel<<<(m + 255) / 256, 256, 0, st>>>(d_boxes, d_scores, m, cfg.track_high_thresh,
                                                     cfg.track_low_thresh, t->det_tlbr, t->det_xywh,
                                                     t->det_xywh32, t->col_kind, t->col_pk,
                                                     inputs_on_device ? reinterpret_cast<float*>(dres_i + o_sc) : nullptr,
                                                     inputs_on_device ? dres_i + o_bx : nullptr);
    BT_LAUNCHED(ctx);
    if (reid)
      BT_TRY(btk_feature_prep(ctx, d_feats, m, D, keep32 || !tensor_path ? t->det_feat32 : nullptr,
                              t->det_feat16, 1, /*dependent of det_prep=*/1));
  }

  SEG_END(BT_SEG_PREP);

  // ---- split lists (demo:1415-1423) -------------------------------------------------------------
  std::vector<int>& unconfirmed = t->v_unconfirmed; unconfirmed.clear();
  std::vector<int>& pool = t->v_pool; pool.clear();
  for (int s : t->tracked) {
    if (!meta[s].activated) unconfirmed.push_back(s);
    else pool.push_back(s);
  }
  for (int s : t->lost) pool.push_back(s);  // joint_stracks: ids are unique per slot, no overlap
  const int n_pool = (int)pool.size(), n_unc = (int)unconfirmed.size();
  const int n_rows = n_rows_pre;

  // ---- Kalman predict over the pool (demo:1426) ---------------------------------------------
  bool all_f32 = n_pool > 0;
  // packed control block A: [pool_idx n_pool][pool_state n_pool][row_kind n_rows] -> one H2D
  int32_t* hA_idx = reinterpret_cast<int32_t*>(t->h_ctrl);
  int32_t* hA_state = hA_idx + n_pool;
  uint8_t* hA_kind = reinterpret_cast<uint8_t*>(hA_state + n_pool);
  const size_t bytesA = sizeof(int32_t) * 2 * n_pool + n_rows;
  int32_t* dA_idx = reinterpret_cast<int32_t*>(t->d_ctrl);
  int32_t* dA_state = dA_idx + n_pool;
  t->row_kind_cur = reinterpret_cast<uint8_t*>(dA_state + n_pool);
  memset(hA_kind, BT_ROW_NONE, n_rows);
  for (int i = 0; i < n_pool; ++i) {
    const int s = pool[i];
    hA_idx[i] = s;
    hA_state[i] = meta[s].state;
    hA_kind[s] = (meta[s].state == BT_STATE_TRACKED) ? BT_ROW_POOL_TRACKED : BT_ROW_POOL_OTHER;
    all_f32 = all_f32 && meta[s].f32_state;
  }
  for (int s : unconfirmed) hA_kind[s] = BT_ROW_UNCONFIRMED;
  // The control block + Kalman predict do not depend on the detection prep; they can run on the side
  // stream next to it and join before the association kernel (BT_PREDICT_SIDE=1), but by default they
  // stay on the main stream: the GPU is waiting for the host's launches at this point of the frame.
  const bool overlap = t->overlap && !t->prof;
  const bool overlap_predict = overlap && t->predict_side;
  cudaStream_t sp = overlap_predict ? t->st2 : st;
  if (overlap_predict) {
    BT_CUDA(cudaEventRecord(t->ev_fork1, st));       // everything already queued on the main stream
    BT_CUDA(cudaStreamWaitEvent(t->st2, t->ev_fork1, 0));
  }
  if (t->debug_delay_us > 0) debug_spin_kernel<<<1, 1, 0, sp>>>(1000ll * t->debug_delay_us);
  if (bytesA > 0) BT_CUDA(cudaMemcpyAsync(t->d_ctrl, t->h_ctrl, bytesA, cudaMemcpyHostToDevice, sp));
  SEG_BEGIN(BT_SEG_PREDICT);
  if (n_pool > 0) {
    ctx->stream = sp;
    const int32_t rc = btk_kalman_predict(ctx, t->mean, t->cov, t->tlbr, t->tlbr_f32, dA_state, dA_idx, n_pool,
                                          all_f32 ? 1 : 0, t->slot_f32);
    ctx->stream = st;
    BT_TRY(rc);
    for (int s : pool) meta[s].f32_state = 0;
  }
  SEG_END(BT_SEG_PREDICT);
  if (overlap_predict) {
    BT_CUDA(cudaEventRecord(t->ev_join1, t->st2));
    BT_CUDA(cudaStreamWaitEvent(st, t->ev_join1, 0));
  }

  bool ema_pending = false, part_a_sent = false;
  // ---- fused association over slots x detections + the three chained LAP solves -------------
  bt_cand cand = *bt_lap_own_cand(ctx);
  const int assoc_bn = (reid && tensor_path) ? btk_assoc_pick_bn(ctx, n_rows, m) : 256;
  cand.seg = assoc_bn / 2;    // one epilogue thread owns one (row, segment) pair; the LAP compaction follows
  if (n_rows > 0) {
    // the candidate counters are left zeroed by the previous frame's LAP kernel (no memset here)
    SEG_BEGIN(BT_SEG_ASSOC);
    if (m > 0) {
      bt_assoc_params p;
      memset(&p, 0, sizeof(p));
      p.a16 = t->feat16; p.b16 = t->det_feat16;
      p.a32 = t->curr32; p.b32 = t->det_feat32;
      p.n = n_rows; p.m = m; p.d = reid ? D : 0;
      p.a_rows_alloc = t->cap; p.b_rows_alloc = t->max_dets;
      p.bn = assoc_bn;
      p.row_tlbr = t->tlbr; p.row_tlbr_f32 = t->tlbr_f32; p.row_kind = t->row_kind_cur;
      p.col_tlbr = t->det_tlbr; p.col_kind = t->col_kind; p.col_pk = t->col_pk; p.face_sim = nullptr;
      p.match_thresh = cfg.match_thresh; p.second_thresh = cfg.second_thresh;
      p.unconf_thresh = cfg.unconfirmed_thresh; p.proximity = cfg.proximity_thresh;
      p.appearance = cfg.appearance_thresh;
      p.cand = cand;
      t->last_assoc = p;
      t->last_assoc_precision = (reid && tensor_path) ? 0 : 1;
      t->last_assoc_valid = true;
      BT_TRY(btk_assoc(ctx, p, t->last_assoc_precision));
    }
    SEG_END(BT_SEG_ASSOC);
    SEG_BEGIN(BT_SEG_LAP);
    const double th[3] = {cfg.match_thresh, cfg.second_thresh, cfg.unconfirmed_thresh};
    BT_TRY(btk_lap_solve3(ctx, cand, n_rows, m, th, t->x, t->y, dres_i + o_hdr));   // also zeroes the pair counter
    SEG_END(BT_SEG_LAP);
    BT_CUDA(cudaMemcpyAsync(t->h_res, t->d_res, bytesA_res, cudaMemcpyDeviceToHost, st));   // part A
    BT_CUDA(cudaEventRecord(t->ev_x, st));
    part_a_sent = true;
    // The matched tracks' Kalman update and feature EMA are a pure function of the three assignment
    // vectors, so they run on the device straight away (STrack.update / re_activate arithmetic,
    // demo:570-610) while the host is still waiting for / digesting the assignments.
    SEG_BEGIN(BT_SEG_UPDATE);
    {
      BT_TRY(btk_kalman_update_x(ctx, t->mean, t->cov, t->tlbr, t->tlbr_f32, t->det_xywh, t->x[0], t->x[1],
                                 t->x[2], t->slot_f32, n_rows, dres_tlbr));
      if (reid && m > 0) {
        // the feature EMA only needs the assignment vectors: side stream, next to update + duplicate test
        if (overlap) {
          BT_CUDA(cudaEventRecord(t->ev_fork2, st));
          BT_CUDA(cudaStreamWaitEvent(t->st2, t->ev_fork2, 0));
          ctx->stream = t->st2;
        }
        const int32_t rc = btk_feature_ema_x(ctx, t->smooth32, t->curr32, keep32 ? t->det_feat32 : nullptr, t->feat16,
                                             t->det_feat16, t->x[0], t->x[1], t->x[2], n_rows, D, cfg.ema_alpha);
        ctx->stream = st;
        BT_TRY(rc);
        if (overlap) { BT_CUDA(cudaEventRecord(t->ev_join2, t->st2)); ema_pending = true; }
      }
    }
    SEG_END(BT_SEG_UPDATE);
    // duplicate candidates among all live slots (superset of tracked x lost) + every slot's box
    SEG_BEGIN(BT_SEG_DUP);
    BT_TRY(btk_iou_pairs_live(ctx, t->tlbr, t->tlbr_f32, t->row_kind_cur, n_rows, cfg.duplicate_iou_dist,
                              t->d_pairs, dres_i + o_hdr, t->pair_cap, dres_i + o_pairs, kPairPrefetch));
    SEG_END(BT_SEG_DUP);
  }
  if (!part_a_sent && inputs_on_device && m > 0) {     // no tracks yet: only scores / boxes come back
    BT_CUDA(cudaMemcpyAsync(t->h_res, t->d_res, bytesA_res, cudaMemcpyDeviceToHost, st));
    BT_CUDA(cudaEventRecord(t->ev_x, st));
    part_a_sent = true;
  }
  if (n_rows > 0)                                      // part B: pair count + first pairs, boxes of all slots
    BT_CUDA(cudaMemcpyAsync(t->h_res + bytesA_res, t->d_res + bytesA_res, bytesB_res, cudaMemcpyDeviceToHost, st));
  HOST_MARK(BT_SEG_HOST_ENQUEUE1);
  // wait only for the assignments (+ scores / boxes): the list bookkeeping below overlaps the GPU's
  // update / EMA / duplicate-test tail
  if (part_a_sent) BT_CUDA(cudaEventSynchronize(t->ev_x));
  HOST_MARK(BT_SEG_HOST_WAIT1);

  // ---- detection lists (demo:1493-1532) -----------------------------------------------------
  const float* sc = t->h_scores;
  std::vector<int>& hi_pos = t->v_hi_pos; hi_pos.assign(m, -1);
  std::vector<int>& lo_pos = t->v_lo_pos; lo_pos.assign(m, -1);
  std::vector<int>& hi_list = t->v_hi_list; hi_list.clear();
  std::vector<int>& lo_list = t->v_lo_list; lo_list.clear();
  for (int j = 0; j < m; ++j) {
    if (sc[j] > cfg.track_high_thresh) { hi_pos[j] = (int)hi_list.size(); hi_list.push_back(j); }
    else if (sc[j] >= cfg.track_low_thresh) { lo_pos[j] = (int)lo_list.size(); lo_list.push_back(j); }
  }
  std::vector<uint8_t>& det_taken = t->v_det_taken; det_taken.assign(m, 0);

  std::vector<int>& activated = t->v_activated; activated.clear();
  std::vector<int>& refind = t->v_refind; refind.clear();
  std::vector<int>& lost_now = t->v_lost_now; lost_now.clear();
  std::vector<int>& removed_now = t->v_removed_now; removed_now.clear();
  auto apply_match = [&](int slot, int det) {
    SlotMeta& tm = meta[slot];
    tm.f32_state = 0;
    if (tm.state == BT_STATE_TRACKED) {  // STrack.update, demo:586-610
      tm.tracklet_len += 1;
      activated.push_back(slot);
    } else {                             // STrack.re_activate, demo:570-584
      tm.tracklet_len = 0;
      refind.push_back(slot);
    }
    tm.frame_id = frame_id;
    tm.state = BT_STATE_TRACKED;
    tm.activated = 1;
    tm.score = sc[det];
    tm.det_index = det;
    det_taken[det] = 1;
  };

  for (auto& mm : t->matches) mm.clear();
  double dbg_t[8]; int dbg_n = 0;
  const bool dbg_host = t->host_debug;
  if (dbg_host) dbg_t[dbg_n++] = now_ms();
  // first association (demo:1556-1566): matches in ascending pool order
  std::vector<uint8_t>& pool_matched = t->v_pool_matched; pool_matched.assign(n_pool, 0);
  for (int i = 0; i < n_pool; ++i) {
    const int s = pool[i];
    const int j = (n_rows > 0) ? t->h_x[0][s] : -1;
    if (j >= 0) {
      t->matches[0].push_back(i);
      t->matches[0].push_back(hi_pos[j]);
      pool_matched[i] = 1;
    }
  }
  // the state test of stage 2 (demo:1569) reads the state BEFORE stage-1 updates touch the
  // unmatched tracks, which they never do; build r_tracked first, then apply stage 1.
  std::vector<int>& r_tracked = t->v_r_tracked; r_tracked.clear();
  for (int i = 0; i < n_pool; ++i)
    if (!pool_matched[i] && meta[pool[i]].state == BT_STATE_TRACKED) r_tracked.push_back(pool[i]);
  for (int i = 0; i < n_pool; ++i)
    if (pool_matched[i]) apply_match(pool[i], t->h_x[0][pool[i]]);
  if (dbg_host) dbg_t[dbg_n++] = now_ms();
  // second association (demo:1568-1586)
  for (int i = 0; i < (int)r_tracked.size(); ++i) {
    const int s = r_tracked[i];
    const int j = t->h_x[1][s];
    if (j >= 0) {
      t->matches[1].push_back(i);
      t->matches[1].push_back(lo_pos[j]);
      apply_match(s, j);
    }
  }
  for (int s : r_tracked) {
    if (t->h_x[1][s] < 0 && meta[s].state != BT_STATE_LOST) {
      meta[s].state = BT_STATE_LOST;  // mark_lost
      lost_now.push_back(s);
    }
  }
  // unconfirmed (demo:1588-1612): detections = unmatched high detections, in order
  std::vector<int>& u_det_pos = t->v_u_det_pos; u_det_pos.assign(m, -1);
  {
    int k = 0;
    for (int j : hi_list)
      if (!det_taken[j]) u_det_pos[j] = k++;
  }
  for (int i = 0; i < n_unc; ++i) {
    const int s = unconfirmed[i];
    const int j = t->h_x[2][s];
    if (j >= 0) {
      t->matches[2].push_back(i);
      t->matches[2].push_back(u_det_pos[j]);
    }
  }
  for (int i = 0; i < n_unc; ++i) {
    const int s = unconfirmed[i];
    const int j = t->h_x[2][s];
    if (j >= 0) apply_match(s, j);
  }
  for (int s : unconfirmed) {
    if (t->h_x[2][s] < 0) {
      meta[s].state = BT_STATE_REMOVED;  // mark_removed, demo:1609-1612
      removed_now.push_back(s);
    }
  }
  if (dbg_host) dbg_t[dbg_n++] = now_ms();
  // births (demo:1614-1621, STrack.activate demo:556-568)
  int n_births = 0;
  for (int j : hi_list) {
    if (det_taken[j]) continue;
    if (sc[j] < cfg.new_track_thresh) continue;
    const int s = alloc_slot(t);
    BT_CHECK(s >= 0, BT_ERR_CAPACITY, "track store full (%d slots)", t->cap);
    SlotMeta& tm = meta[s];
    tm = SlotMeta();
    tm.used = 1;
    tm.track_id = ++t->id_count;
    tm.state = BT_STATE_TRACKED;
    tm.activated = (frame_id == 1) ? 1 : 0;
    tm.frame_id = frame_id;
    tm.start_frame = frame_id;
    tm.tracklet_len = 0;
    tm.score = sc[j];
    tm.det_index = j;
    tm.f32_state = 1;
    t->h_birth_slot[n_births] = s;
    t->h_birth_det[n_births] = j;
    ++n_births;
    activated.push_back(s);
  }
  // expiry (demo:1623-1627)
  for (int s : t->lost) {
    if (frame_id - meta[s].frame_id > t->max_time_lost) {
      meta[s].state = BT_STATE_REMOVED;
      removed_now.push_back(s);
    }
  }

  if (dbg_host) dbg_t[dbg_n++] = now_ms();
  // ---- merge lists (demo:1629-1636) -----------------------------------------------------------
  std::vector<int>& new_tracked = t->v_new_tracked; new_tracked.clear();
  for (int s : t->tracked)
    if (meta[s].state == BT_STATE_TRACKED) { new_tracked.push_back(s); meta[s].mark = 1; }
  for (int s : activated)
    if (!meta[s].mark) { new_tracked.push_back(s); meta[s].mark = 1; }
  for (int s : refind)
    if (!meta[s].mark) { new_tracked.push_back(s); meta[s].mark = 1; }
  std::vector<int>& new_lost = t->v_new_lost; new_lost.clear();
  for (int s : t->lost)
    if (!meta[s].mark) new_lost.push_back(s);       // sub_stracks(lost, tracked)
  for (int s : lost_now) new_lost.push_back(s);     // extend(lost_stracks)
  {
    std::vector<int>& tmp = t->v_tmp; tmp.clear();
    for (int s : new_lost)
      if (!meta[s].in_removed) tmp.push_back(s);    // sub_stracks(lost, removed) BEFORE this frame's removals
    new_lost.swap(tmp);
  }
  for (int s : new_tracked) meta[s].mark = 0;
  for (int s : removed_now) meta[s].in_removed = 1; // removed_stracks.extend
  t->n_removed_total += (int)removed_now.size();

  if (dbg_host) {
    dbg_t[dbg_n++] = now_ms();
    fprintf(stderr, "host lists: detlists %.1f stage1 %.1f stage2+3 %.1f births+expiry %.1f merge %.1f us\n",
            1e3 * (dbg_t[0] - t_host), 1e3 * (dbg_t[1] - dbg_t[0]), 1e3 * (dbg_t[2] - dbg_t[1]),
            1e3 * (dbg_t[3] - dbg_t[2]), 1e3 * (dbg_t[4] - dbg_t[3]));
  }
  // ---- births on the device: one packed H2D, Kalman initiate + feature adoption --------------------
  const int nt = (int)new_tracked.size(), nl = (int)new_lost.size();
  HOST_MARK(BT_SEG_HOST_LISTS);
  BT_CUDA(cudaStreamSynchronize(st));    // the frame's device work is complete, part B is on the host
  HOST_MARK(BT_SEG_HOST_WAIT2);
  if (ema_pending) BT_CUDA(cudaEventSynchronize(t->ev_join2));   // the side stream's EMA (usually done already)
  BT_TRY(prof_collect(ctx, t));
  int n_pairs = hres_i[o_hdr];           // live-slot duplicate candidates found by the device
  if (n_rows <= 1) n_pairs = 0;
  BT_CHECK(n_pairs <= t->pair_cap, BT_ERR_CAPACITY, "%d duplicate pairs exceed capacity %d", n_pairs, t->pair_cap);
  bool need_sync = false;
  const int32_t* live_pairs = hres_i + o_pairs;
  if (n_pairs > kPairPrefetch) {
    BT_CUDA(cudaMemcpyAsync(t->h_pairs, t->d_pairs, sizeof(int32_t) * 2 * (size_t)n_pairs, cudaMemcpyDeviceToHost, st));
    live_pairs = t->h_pairs;
    need_sync = true;
  }
  int n_bpairs = 0;
  if (n_births > 0) {
    int32_t* hB = reinterpret_cast<int32_t*>(t->h_ctrl);
    int32_t* dB = reinterpret_cast<int32_t*>(t->d_ctrl);
    const bool birth_dup = nl > 0;       // a newborn can duplicate a lost track: test births x lost
    memcpy(hB, t->h_birth_slot, sizeof(int32_t) * n_births);
    memcpy(hB + n_births, t->h_birth_det, sizeof(int32_t) * n_births);
    if (birth_dup) memcpy(hB + 2 * n_births, new_lost.data(), sizeof(int32_t) * nl);
    uint8_t* hB8 = reinterpret_cast<uint8_t*>(hB + 2 * n_births + (birth_dup ? nl : 0));
    memset(hB8, 2, n_births);              // feature mode 2: adopt the detection's normalised feature
    const size_t bytesB = sizeof(int32_t) * (2 * (size_t)n_births + (birth_dup ? nl : 0)) + n_births;
    BT_CUDA(cudaMemcpyAsync(t->d_ctrl, t->h_ctrl, bytesB, cudaMemcpyHostToDevice, st));
    const int32_t *d_birth_slot = dB, *d_birth_det = dB + n_births, *d_listb = dB + 2 * n_births;
    const uint8_t* d_mode = reinterpret_cast<const uint8_t*>(dB + 2 * n_births + (birth_dup ? nl : 0));
    BT_TRY(btk_kalman_initiate(ctx, t->det_xywh32, d_birth_det, t->mean, t->cov, t->tlbr, t->tlbr_f32,
                               d_birth_slot, n_births, t->slot_f32));
    if (reid)
      BT_TRY(btk_feature_ema16(ctx, t->smooth32, t->curr32, keep32 ? t->det_feat32 : nullptr, t->feat16,
                               t->det_feat16, d_birth_slot, d_birth_det, d_mode, n_births, D, cfg.ema_alpha));
    if (birth_dup) {
      BT_CUDA(cudaMemsetAsync(t->d_pair_count, 0, sizeof(int32_t), st));
      BT_TRY(btk_iou_pairs_below(ctx, t->tlbr, d_birth_slot, n_births, d_listb, nl, cfg.duplicate_iou_dist,
                                 t->d_bpairs, t->d_pair_count, t->bpair_cap));
      BT_CUDA(cudaMemcpyAsync(t->h_pair_count, t->d_pair_count, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
      need_sync = true;
    }
  }
  if (need_sync) {
    BT_CUDA(cudaStreamSynchronize(st));  // rare second sync: births next to lost tracks / very many candidates
    if (n_births > 0 && nl > 0) {
      n_bpairs = *t->h_pair_count;
      BT_CHECK(n_bpairs <= t->bpair_cap, BT_ERR_CAPACITY, "%d duplicate pairs exceed capacity %d", n_bpairs,
               t->bpair_cap);
      if (n_bpairs > 0) {
        BT_CUDA(cudaMemcpyAsync(t->h_bpairs, t->d_bpairs, sizeof(int32_t) * 2 * n_bpairs,
                                cudaMemcpyDeviceToHost, st));
        BT_CUDA(cudaStreamSynchronize(st));
      }
    }
  }

  // ---- remove_duplicate_stracks (demo:1637, demo:1665-1680) --------------------------------------
  std::vector<uint8_t>& dupa = t->scratch_a; dupa.assign(nt, 0);
  std::vector<uint8_t>& dupb = t->scratch_b; dupb.assign(nl, 0);
  if (n_pairs > 0 || n_bpairs > 0) {
    std::vector<int>& pos_t = t->scratch_pos_t; pos_t.assign(t->high_water, -1);
    std::vector<int>& pos_l = t->scratch_pos_l; pos_l.assign(t->high_water, -1);
    for (int i = 0; i < nt; ++i) pos_t[new_tracked[i]] = i;
    for (int i = 0; i < nl; ++i) pos_l[new_lost[i]] = i;
    auto resolve = [&](int p, int q) {   // p: position in tracked, q: position in lost (demo:1669-1677)
      const int timep = meta[new_tracked[p]].frame_id - meta[new_tracked[p]].start_frame;
      const int timeq = meta[new_lost[q]].frame_id - meta[new_lost[q]].start_frame;
      if (timep > timeq) dupb[q] = 1;
      else dupa[p] = 1;
    };
    for (int k = 0; k < n_pairs; ++k) {
      const int i = live_pairs[2 * k], j = live_pairs[2 * k + 1];
      if (pos_t[i] >= 0 && pos_l[j] >= 0) resolve(pos_t[i], pos_l[j]);
      if (pos_t[j] >= 0 && pos_l[i] >= 0) resolve(pos_t[j], pos_l[i]);
    }
    for (int k = 0; k < n_bpairs; ++k)     // (birth index, lost position)
      resolve(pos_t[t->h_birth_slot[t->h_bpairs[2 * k]]], t->h_bpairs[2 * k + 1]);
  }
  t->tracked.clear();
  t->lost.clear();
  t->tlbr_cache.resize(4 * (size_t)nt);
  size_t n_out = 0;
  for (int i = 0; i < nt; ++i) {
    if (dupa[i]) continue;
    const int s = new_tracked[i];
    t->tracked.push_back(s);
    double* dst = t->tlbr_cache.data() + 4 * n_out++;
    const bool born_now = meta[s].f32_state && meta[s].start_frame == frame_id;
    if (!born_now) {
      memcpy(dst, hres_tlbr + 4 * (size_t)s, 4 * sizeof(double));
    } else {
      // a track born this frame: its box is the detection's (demo:624-648 on initiate's mean)
      const int32_t* bx = host_boxes + 4 * (size_t)meta[s].det_index;
      for (int c = 0; c < 4; ++c) dst[c] = (double)bx[c];
    }
  }
  t->tlbr_cache.resize(4 * n_out);
  for (int i = 0; i < nl; ++i)
    if (!dupb[i]) t->lost.push_back(new_lost[i]);
  t->tlbr_cache_valid = true;

  // ---- recycle slots that left both lists -----------------------------------------------------
  for (int s : t->tracked) meta[s].mark = 1;
  for (int s : t->lost) meta[s].mark = 1;
  bool freed = false;
  for (int s = 0; s < t->high_water; ++s) {
    if (meta[s].used && !meta[s].mark) {
      meta[s] = SlotMeta();
      t->free_slots.push_back(s);
      freed = true;
    }
    meta[s].mark = 0;
  }
  if (freed) std::sort(t->free_slots.begin(), t->free_slots.end(), std::greater<int>());

  HOST_MARK(BT_SEG_HOST_FINAL);
  if (t->host_debug)
    fprintf(stderr, "frame %d host phases (us): enqueue %.1f wait_x %.1f lists+births %.1f wait_end %.1f final %.1f\n", frame_id,
            1e3 * hphase[0], 1e3 * hphase[1], 1e3 * hphase[2], 1e3 * hphase[3], 1e3 * hphase[4]);
  if (info) {
    info->frame_id = frame_id;
    info->n_tracked = (int)t->tracked.size();
    info->n_lost = (int)t->lost.size();
    info->n_removed_total = t->n_removed_total;
    info->n_pool = n_pool;
    info->n_high = (int)hi_list.size();
    info->n_low = (int)lo_list.size();
    info->n_unconfirmed = n_unc;
    info->n_matches1 = (int)t->matches[0].size() / 2;
    info->n_matches2 = (int)t->matches[1].size() / 2;
    info->n_matches3 = (int)t->matches[2].size() / 2;
    info->n_births = n_births;
  }
  return BT_OK;
}

int32_t bt_profile_enable(bt_ctx* ctx, int32_t on) {
  if (!ctx) return BT_ERR_INVALID;
  BT_CUDA(cudaSetDevice(ctx->device));
  bt_tracker* t = ctx->trk;
  BT_CUDA(cudaStreamSynchronize(ctx->stream));
  if (on) {
    for (int s = 0; s < BT_SEG_HOST_ENQUEUE1; ++s)
      for (cudaEvent_t& e : t->ev[s])
        if (!e) BT_CUDA(cudaEventCreate(&e));
    for (int s = 0; s < BT_SEG_COUNT; ++s) { t->prof_ms[s] = 0.0; t->prof_n[s] = 0; t->seg_open[s] = false; }
  }
  t->prof = on != 0;
  return BT_OK;
}

int32_t bt_profile_replay_assoc(bt_ctx* ctx, int32_t iters, double* total_ms) {
  if (!ctx) return BT_ERR_INVALID;
  BT_CUDA(cudaSetDevice(ctx->device));
  bt_tracker* t = ctx->trk;
  BT_CHECK(t->last_assoc_valid, BT_ERR_STATE, "no association launch to replay yet");
  BT_CHECK(iters > 0 && iters <= 1000 && total_ms, BT_ERR_INVALID, "iters must be 1..1000");
  cudaStream_t st = ctx->stream;
  const bt_cand& cand = *bt_lap_own_cand(ctx);
  // The tensor-core kernel's emission is idempotent (every (row, segment) owner rewrites the same
  // entries and the same count), so its launches can run back to back inside ONE event pair; the
  // CUDA-core kernel appends with atomics and needs its lists cleared between launches.
  const bool idempotent = t->last_assoc_precision == 0;
  cudaEvent_t e0, e1;
  BT_CUDA(cudaEventCreate(&e0));
  BT_CUDA(cudaEventCreate(&e1));
  BT_CUDA(cudaStreamSynchronize(st));
  double sum = 0.0;
  if (idempotent) {
    BT_TRY(btk_assoc(ctx, t->last_assoc, t->last_assoc_precision));   // warm
    BT_CUDA(cudaEventRecord(e0, st));
    for (int i = 0; i < iters; ++i) BT_TRY(btk_assoc(ctx, t->last_assoc, t->last_assoc_precision));
    BT_CUDA(cudaEventRecord(e1, st));
    BT_CUDA(cudaStreamSynchronize(st));
    float ms = 0.f;
    BT_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    sum = ms;
  } else {
    for (int i = 0; i < iters; ++i) {
      BT_CUDA(cudaMemsetAsync(cand.cnt, 0, cand.clear_bytes, st));
      BT_CUDA(cudaEventRecord(e0, st));
      BT_TRY(btk_assoc(ctx, t->last_assoc, t->last_assoc_precision));
      BT_CUDA(cudaEventRecord(e1, st));
      BT_CUDA(cudaStreamSynchronize(st));
      float ms = 0.f;
      BT_CUDA(cudaEventElapsedTime(&ms, e0, e1));
      sum += ms;
    }
  }
  BT_CUDA(cudaMemsetAsync(cand.cnt, 0, cand.clear_bytes, st));
  BT_CUDA(cudaStreamSynchronize(st));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *total_ms = sum;
  return BT_OK;
}

int32_t bt_profile_read(bt_ctx* ctx, int32_t segment, double* total_ms, int64_t* samples) {
  if (!ctx) return BT_ERR_INVALID;
  BT_CHECK(segment >= 0 && segment < BT_SEG_COUNT, BT_ERR_INVALID, "bad segment %d", segment);
  if (total_ms) *total_ms = ctx->trk->prof_ms[segment];
  if (samples) *samples = ctx->trk->prof_n[segment];
  return BT_OK;
}

int32_t bt_get_tracks(bt_ctx* ctx, int32_t which, int32_t cap, int32_t* n, int32_t* ids, int32_t* state,
                      int32_t* activated, int32_t* frame_id, int32_t* start_frame, int32_t* tracklet_len,
                      int32_t* det_index, float* score, double* tlbr, double* mean, double* cov) {
  if (!ctx) return BT_ERR_INVALID;
  BT_CUDA(cudaSetDevice(ctx->device));
  bt_tracker* t = ctx->trk;
  BT_CHECK(which == 0 || which == 1, BT_ERR_INVALID, "which must be 0 (tracked) or 1 (lost)");
  const std::vector<int>& lst = which == 0 ? t->tracked : t->lost;
  const int cnt = (int)lst.size();
  if (n) *n = cnt;
  BT_CHECK(cnt <= cap || !(ids || state || activated || frame_id || start_frame || tracklet_len || det_index ||
                           score || tlbr || mean || cov),
           BT_ERR_CAPACITY, "list has %d tracks, buffers hold %d", cnt, cap);
  for (int i = 0; i < cnt; ++i) {
    const SlotMeta& tm = t->meta[lst[i]];
    if (ids) ids[i] = tm.track_id;
    if (state) state[i] = tm.state;
    if (activated) activated[i] = tm.activated;
    if (frame_id) frame_id[i] = tm.frame_id;
    if (start_frame) start_frame[i] = tm.start_frame;
    if (tracklet_len) tracklet_len[i] = tm.tracklet_len;
    if (det_index) det_index[i] = tm.det_index;
    if (score) score[i] = tm.score;
  }
  if (cnt == 0) return BT_OK;
  if (tlbr && which == 0 && t->tlbr_cache_valid) {
    memcpy(tlbr, t->tlbr_cache.data(), sizeof(double) * 4 * cnt);
    tlbr = nullptr;
  }
  if (tlbr || mean || cov) {
    cudaStream_t st = ctx->stream;
    memcpy(t->h_lista, lst.data(), sizeof(int32_t) * cnt);
    BT_CUDA(cudaMemcpyAsync(t->d_lista, t->h_lista, sizeof(int32_t) * cnt, cudaMemcpyHostToDevice, st));
    struct Job { const double* src; double* dst; int width; };
    const Job jobs[3] = {{t->tlbr, tlbr, 4}, {t->mean, mean, 8}, {t->cov, cov, 64}};
    for (const Job& j : jobs) {
      if (!j.dst) continue;
      gather_rows_f64_kernel<<<(unsigned)(((size_t)cnt * j.width + 255) / 256), 256, 0, st>>>(
          j.src, t->d_lista, cnt, j.width, t->d_gather);
      BT_LAUNCHED(ctx);
      BT_CUDA(cudaMemcpyAsync(j.dst, t->d_gather, sizeof(double) * (size_t)cnt * j.width, cudaMemcpyDeviceToHost, st));
      BT_CUDA(cudaStreamSynchronize(st));
    }
  }
  return BT_OK;
}

int32_t bt_get_track_features(bt_ctx* ctx, int32_t which, int32_t cap, float* curr, float* smooth) {
  if (!ctx) return BT_ERR_INVALID;
  BT_CUDA(cudaSetDevice(ctx->device));
  bt_tracker* t = ctx->trk;
  BT_CHECK(which == 0 || which == 1, BT_ERR_INVALID, "which must be 0 (tracked) or 1 (lost)");
  BT_CHECK(t->curr32 != nullptr, BT_ERR_STATE, "ctx was created with BT_FLAG_NO_F32_FEATURES");
  const std::vector<int>& lst = which == 0 ? t->tracked : t->lost;
  const int cnt = (int)lst.size();
  BT_CHECK(cnt <= cap, BT_ERR_CAPACITY, "list has %d tracks, buffers hold %d", cnt, cap);
  if (cnt == 0) return BT_OK;
  cudaStream_t st = ctx->stream;
  memcpy(t->h_lista, lst.data(), sizeof(int32_t) * cnt);
  BT_CUDA(cudaMemcpyAsync(t->d_lista, t->h_lista, sizeof(int32_t) * cnt, cudaMemcpyHostToDevice, st));
  const float* srcs[2] = {t->curr32, t->smooth32};
  float* dsts[2] = {curr, smooth};
  for (int k = 0; k < 2; ++k) {
    if (!dsts[k]) continue;
    // reuse the normalised-detection staging buffer as gather scratch in chunks of max_dets rows
    for (int off = 0; off < cnt; off += t->max_dets) {
      const int rows = std::min(t->max_dets, cnt - off);
      gather_rows_f32_kernel<<<(unsigned)(((size_t)rows * t->D + 255) / 256), 256, 0, st>>>(
          srcs[k], t->d_lista + off, rows, t->D, t->det_feat_in);
      BT_LAUNCHED(ctx);
      BT_CUDA(cudaMemcpyAsync(dsts[k] + (size_t)off * t->D, t->det_feat_in, sizeof(float) * (size_t)rows * t->D,
                              cudaMemcpyDeviceToHost, st));
      BT_CUDA(cudaStreamSynchronize(st));
    }
  }
  return BT_OK;
}

int32_t bt_get_matches(bt_ctx* ctx, int32_t stage, int32_t cap, int32_t* n, int32_t* pairs) {
  if (!ctx) return BT_ERR_INVALID;
  bt_tracker* t = ctx->trk;
  BT_CHECK(stage >= 1 && stage <= 3, BT_ERR_INVALID, "stage must be 1, 2 or 3");
  const std::vector<int32_t>& mm = t->matches[stage - 1];
  const int cnt = (int)mm.size() / 2;
  if (n) *n = cnt;
  if (pairs) {
    BT_CHECK(cnt <= cap, BT_ERR_CAPACITY, "%d matches, buffer holds %d", cnt, cap);
    if (cnt) memcpy(pairs, mm.data(), sizeof(int32_t) * 2 * cnt);
  }
  return BT_OK;
}

}  // extern "C"
