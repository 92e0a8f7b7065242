// The stateful trackers of a ctx: one BoTSORT.update (demo:1291-1639; demo =
// /root/reference/demo_bottrack_onnx_tflite.py) per video stream and frame step, on a device-resident
// track store, any number of the ctx's video streams per call.
//
// Device side (all arithmetic; every launch serves every stream of the batch): input cast (fp32 ingest
// only) -> detection prep + batched Kalman predict -> ONE fused association kernel over (all live slots)
// x (all detections) of every stream that emits the candidate edges of the three association stages at
// once -> the three exact LAP solves chained on the device, one CTA per stream (stage 2 masked by stage
// 1's unmatched rows, stage 3 by stage 1's unmatched columns) -> Kalman update + feature EMA -> sparse
// duplicate test (tracked x lost, IoU distance < 0.15).
//
// Host side (this file, C++): only the list bookkeeping of demo:1414-1423 and demo:1558-1639
// (who is tracked / lost / removed, ids, list order) on small index arrays, with two stream
// synchronisations per step.  Inputs are double-buffered: bt_submit_streams moves frame k+1 in on a copy
// stream while bt_step_streams works on frame k.
#include "common.cuh"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <stdlib.h>
#include <string.h>

namespace {

constexpr int kPairPrefetch = 512;   // duplicate pairs fetched with the second read-back (more: extra copy)
constexpr int kPairCap = 1 << 16;    // duplicate candidates the device keeps per stream (more: exact host fallback)

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

struct FrameIn {           // one submitted frame of one video stream
  int m = 0;
  int dtype = BT_F32;
  int loc = BT_HOST;
  const float* face_sim = nullptr;   // device pointer (staged) or null
  const int32_t* host_boxes = nullptr;
  const float* host_scores = nullptr;
  int ev = -1;             // index of the copy-stream event that marks the inputs complete (-1: none)
};

// host bookkeeping of one video stream (= one BoTSORT instance of the reference)
struct StreamState {
  bt_config cfg;
  int max_time_lost = 0;
  std::vector<SlotMeta> meta;
  std::vector<int> tracked, lost;
  std::vector<int> free_slots;  // kept sorted descending: back() is the lowest free slot
  int high_water = 0;
  int id_count = 0;
  int frame_id = 0;
  int n_removed_total = 0;
  int feat_dtype = -1;          // dtype of the features the store holds (fixed while tracks exist)
  std::vector<int32_t> matches[3];  // flattened (a, b) pairs in the reference's index spaces
  // input queue: parities of submitted-but-not-stepped frames, oldest first
  FrameIn in[2];
  int pending[2] = {0, 0};
  int n_pending = 0;
  int next_parity = 0;
  float* face_dev = nullptr;    // staged face similarities (device), grown on demand
  size_t face_cap = 0;
  // per-step temporaries
  int parity = 0, m = 0, n_pool = 0, n_unc = 0, n_rows = 0, n_births = 0, n_births_skipped = 0;
  bool device_inputs = false;
  const float* sc = nullptr;            // this frame's scores on the host
  const int32_t* host_boxes = nullptr;  // this frame's boxes on the host
  const int32_t* hx[3] = {nullptr, nullptr, nullptr};
  const int32_t* pool = nullptr;        // this step's pool (slots, pool order) inside the pinned control block
  int n_high = 0, n_low = 0, n_m1 = 0, n_m2 = 0, n_m3 = 0;
  bool matches_valid = false;
  std::vector<int> v_unconfirmed, v_pool, v_hi_pos, v_lo_pos, v_hi_list, v_lo_list, v_activated, v_refind, v_lost_now,
      v_removed_now, v_r_tracked, v_u_det_pos, v_new_tracked, v_new_lost, v_tmp, v_birth_slot, v_birth_det;
  std::vector<uint8_t> v_det_taken, v_pool_matched, scratch_a, scratch_b;
  std::vector<int> scratch_pos_t, scratch_pos_l;
  const double* tlbr_src = nullptr; // boxes of all slots after the last step (host copy of result part B)
  int tlbr_n_rows = 0;
  int tlbr_region = 0;              // result region the copy lives in, and that region's generation at the time
  uint64_t tlbr_gen = 0;
  bool boxes_valid = false;         // host_boxes (this frame's detections) has not been overwritten since the step
  bool tlbr_cache_valid = false;
  // The NEXT frame's pool lists (control segment without the face positions), built from this frame's merged lists
  // while the GPU finishes the frame (the host would otherwise idle in the final sync); dropped when the finish
  // phase changes the lists (duplicates) or a pooled track still carries its float32 birth state.
  bool pre_valid = false;
  int pre_n_pool = 0, pre_n_rows = 0;
  std::vector<int32_t> pre_idx, pre_state;
  std::vector<uint8_t> pre_kind;
  std::vector<int> pre_unconfirmed;
};

}  // namespace

// Per-frame launch description, first thing in the control block (pinned host copy -> device copy with ONE H2D
// that also carries the per-stream control segments): everything a frame's kernels need to know that changes
// from frame to frame.  Their kernel ARGUMENTS therefore never change, and a captured CUDA graph of the frame
// can be replayed with no per-launch patching.
struct FrameDesc {
  bt_batch B;
  bt_assoc_frame AF;
  bt_lap_batch LB;
};
constexpr size_t kDescBytes = (sizeof(FrameDesc) + 255) & ~size_t(255);

struct GraphKey {
  int count, f32, reid, dev, bn, precision, f16;
  bool operator==(const GraphKey& o) const {
    return count == o.count && f32 == o.f32 && reid == o.reid && dev == o.dev && bn == o.bn && precision == o.precision && f16 == o.f16;
  }
};
struct FrameGraph {
  GraphKey key;
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  int n_kernels = 0;
};

struct bt_tracker {
  int S = 1, cap = 0, md = 0, D = 0;
  // the frame step as a CUDA graph, one per launch shape (opt-in: BT_GRAPH=1)
  bool use_graph = false;
  std::vector<GraphKey> seen_keys;      // shapes that ran once the plain way (module loading, function attributes)
  std::vector<FrameGraph> graphs;
  bt_store st = {};
  bt_res_layout L = {};
  std::vector<StreamState> streams;
  size_t ctrl_stride = 0;           // worst-case bytes of one stream's control segment
  // pinned host mirrors
  char* h_ctrl = nullptr;
  char* h_res = nullptr;            // part A regions
  uint32_t* h_flags = nullptr;      // direct results: one word per batch entry, set by the LAP kernel (pinned)
  // device-side aliases of the pinned buffers kernels touch directly (cudaHostGetDevicePointer; identical under UVA)
  char* h_ctrl_dev = nullptr; char* h_res_dev = nullptr; uint32_t* h_flags_dev = nullptr;
  bool direct_results = false;      // BT_DIRECT_RESULT=1: the LAP kernel publishes the assignments into pinned host memory (A/B: slower)
  char* h_resB = nullptr;           // part B regions
  int32_t* h_in_boxes = nullptr;    // [2][S*md][4] pinned copies of the submitted boxes / scores of host inputs: the
  float* h_in_scores = nullptr;     // [2][S*md]     list bookkeeping reads them at step time (the caller's may be gone)
  int32_t* h_birth = nullptr;       // [S][2*md + cap] birth slot / det lists + lost list
  int32_t* d_birth = nullptr;
  int32_t* h_pairs = nullptr;       // [2*kPairCap] overflow fetch (one stream at a time)
  int32_t* h_bpairs = nullptr; int32_t* d_bpairs = nullptr; int32_t* d_bpair_count = nullptr; int32_t* h_bpair_count = nullptr;
  int bpair_cap = 1 << 16;
  int32_t* h_list = nullptr; int32_t* d_list = nullptr;   // list read-backs
  double* d_gather = nullptr;       // [cap*64]
  float* d_gather32 = nullptr;      // [cap*D] (allocated on first use)
  uint64_t region_gen[BT_MAX_BATCH] = {};   // bumped whenever a step overwrites result region k
  cudaEvent_t ev_x = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;   // side stream (feature EMA) fork / join
  cudaEvent_t ev_tail = nullptr;    // end of the work a step left running on the main stream (births)
  bool tail_pending = false;
  std::vector<cudaEvent_t> in_events;
  int in_seq = 0;
  bool host_debug = false;
  bool no_refine = false;           // BT_NO_REFINE=1: tests show what the exact re-costing buys
  bool no_l2_prefetch = false;      // BT_NO_L2_PREFETCH=1 (A/B)
  bool copy_inline = false;         // BT_COPY_INLINE=1: assignment read-back on the main stream (A/B)
  bool no_prebuild = false;         // BT_NO_PREBUILD=1: pool lists built at the start of the step (A/B)
  bool ctrl_by_copy = false;        // BT_CTRL_COPY=1: control block by cudaMemcpyAsync instead of the upload kernel (A/B)
  bt_assoc_params last_assoc;       // for bt_profile_replay_assoc
  bt_assoc_frame last_AF;
  int last_assoc_precision = 0;
  bool last_assoc_valid = false;
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

int alloc_slot(StreamState& s, int cap) {
  if (!s.free_slots.empty()) {
    const int v = s.free_slots.back();
    s.free_slots.pop_back();
    return v;
  }
  if (s.high_water < cap) return s.high_water++;
  return -1;
}

void reset_stream(bt_tracker* t, StreamState& s, const bt_config* cfg) {
  if (cfg) s.cfg = *cfg;
  else bt_default_config(&s.cfg);
  s.max_time_lost = (int)(s.cfg.frame_rate / 30.0 * s.cfg.track_buffer);  // demo:1276-1277
  s.meta.assign(t->cap, SlotMeta());
  s.tracked.clear();
  s.lost.clear();
  s.pre_valid = false;
  s.free_slots.clear();
  s.high_water = 0;
  s.id_count = 0;  // BaseTrack.clear_count(), demo:1264 (per tracker here)
  s.frame_id = 0;
  s.n_removed_total = 0;
  s.feat_dtype = -1;
  for (auto& mm : s.matches) mm.clear();
  s.tlbr_cache_valid = false;
  s.matches_valid = false;
  s.n_pending = 0;
  s.next_parity = 0;
}

// control segment of one stream: [pool_idx n_pool][pool_state n_pool][row_kind n_rows (pad 4)][pool_pos n_rows]
inline size_t ctrl_bytes(int n_pool, int n_rows, bool with_pos) {
  return 8 * (size_t)n_pool + (((size_t)n_rows + 3) & ~size_t(3)) + (with_pos ? 4 * (size_t)n_rows : 0);
}

}  // namespace

int32_t bt_tracker_create(bt_ctx* ctx) {
  auto* t = new bt_tracker();
  ctx->trk = t;
  const int S = ctx->n_streams;
  t->S = S;
  t->cap = ctx->max_tracks;
  t->md = ctx->max_dets;
  t->D = ctx->feat_dim;
  const size_t cap = t->cap, md = t->md, D = t->D, NS = (size_t)S * cap, ND = (size_t)S * md;
  const bool keep_smooth = !(ctx->flags & BT_FLAG_NO_F32_FEATURES);
  bt_store& st = t->st;
  st.S = S; st.cap = t->cap; st.md = t->md; st.D = t->D;
  st.pair_cap = kPairCap;
  t->L = bt_res_layout_for(t->cap, t->md, kPairPrefetch);
  BT_TRY(dev_alloc(ctx, &st.mean, NS * 8));
  BT_TRY(dev_alloc(ctx, &st.cov, NS * 64));
  BT_TRY(dev_alloc(ctx, &st.tlbr, NS * 4));
  BT_TRY(dev_alloc(ctx, &st.tlbr_f32, NS * 4));
  BT_TRY(dev_alloc(ctx, &st.feat16, NS * D));
  BT_TRY(dev_alloc(ctx, &st.norm, NS));
  BT_CUDA(cudaMemset(st.feat16, 0, sizeof(__half) * NS * D));
  BT_CUDA(cudaMemset(st.norm, 0, sizeof(float) * NS));
  BT_CUDA(cudaMemset(st.tlbr, 0, sizeof(double) * NS * 4));
  BT_CUDA(cudaMemset(st.tlbr_f32, 0, sizeof(float) * NS * 4));
  if (keep_smooth) {
    BT_TRY(dev_alloc(ctx, &st.smooth32, NS * D));
    BT_CUDA(cudaMemset(st.smooth32, 0, sizeof(float) * NS * D));
  }
  BT_TRY(dev_alloc(ctx, &st.slot_f32, NS));
  BT_CUDA(cudaMemset(st.slot_f32, 0, NS));
  BT_TRY(dev_alloc(ctx, &st.det_boxes, 2 * ND * 4));
  BT_TRY(dev_alloc(ctx, &st.det_scores, 2 * ND));
  BT_TRY(dev_alloc(ctx, &st.det16, 2 * ND * D));
  BT_CUDA(cudaMemset(st.det16, 0, sizeof(__half) * 2 * ND * D));
  BT_TRY(dev_alloc(ctx, &st.det_norm, ND));
  BT_TRY(dev_alloc(ctx, &st.det_tlbr, ND * 4));
  BT_TRY(dev_alloc(ctx, &st.det_xywh, ND * 4));
  BT_TRY(dev_alloc(ctx, &st.det_xywh32, ND * 4));
  BT_TRY(dev_alloc(ctx, &st.col_kind, ND));
  BT_TRY(dev_alloc(ctx, &st.col_pk, ND));
  t->ctrl_stride = (ctrl_bytes(t->cap, t->cap, true) + 255) & ~size_t(255);
  BT_TRY(dev_alloc(ctx, &st.ctrl, kDescBytes + t->ctrl_stride * S));
  BT_TRY(dev_alloc(ctx, &st.res, t->L.stride * S));
  BT_TRY(dev_alloc(ctx, &st.resB, t->L.strideB * S));
  BT_TRY(dev_alloc(ctx, &st.y, (size_t)S * 3 * md));
  BT_TRY(dev_alloc(ctx, &st.pairs, (size_t)S * 2 * kPairCap));
  BT_TRY(dev_alloc(ctx, &t->d_birth, (size_t)S * (2 * md + cap)));
  BT_TRY(dev_alloc(ctx, &t->d_bpairs, (size_t)2 * t->bpair_cap));
  BT_TRY(dev_alloc(ctx, &t->d_bpair_count, 1));
  BT_TRY(dev_alloc(ctx, &t->d_list, cap > md ? cap : md));
  BT_TRY(dev_alloc(ctx, &t->d_gather, cap * 64));

  BT_CUDA(cudaMallocHost(&t->h_ctrl, kDescBytes + t->ctrl_stride * S));
  BT_CUDA(cudaMallocHost(&t->h_res, t->L.stride * S));
  BT_CUDA(cudaMallocHost(&t->h_resB, t->L.strideB * S));
  BT_CUDA(cudaMallocHost(&t->h_flags, sizeof(uint32_t) * BT_MAX_BATCH));
  memset(t->h_flags, 0, sizeof(uint32_t) * BT_MAX_BATCH);
  BT_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void**>(&t->h_ctrl_dev), t->h_ctrl, 0));
  BT_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void**>(&t->h_res_dev), t->h_res, 0));
  BT_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void**>(&t->h_flags_dev), t->h_flags, 0));
  BT_CUDA(cudaMallocHost(&t->h_in_boxes, sizeof(int32_t) * 2 * ND * 4));
  BT_CUDA(cudaMallocHost(&t->h_in_scores, sizeof(float) * 2 * ND));
  BT_CUDA(cudaMallocHost(&t->h_birth, sizeof(int32_t) * (size_t)S * (2 * md + cap)));
  BT_CUDA(cudaMallocHost(&t->h_pairs, sizeof(int32_t) * 2 * (size_t)kPairCap));
  BT_CUDA(cudaMallocHost(&t->h_bpairs, sizeof(int32_t) * 2 * (size_t)t->bpair_cap));
  BT_CUDA(cudaMallocHost(&t->h_bpair_count, 64));
  BT_CUDA(cudaMallocHost(&t->h_list, sizeof(int32_t) * cap));
  t->streams.resize(S);
  for (auto& s : t->streams) reset_stream(t, s, nullptr);
  BT_CUDA(cudaEventCreateWithFlags(&t->ev_x, cudaEventDisableTiming));
  BT_CUDA(cudaEventCreateWithFlags(&t->ev_tail, cudaEventDisableTiming));
  BT_CUDA(cudaEventCreateWithFlags(&t->ev_fork, cudaEventDisableTiming));
  BT_CUDA(cudaEventCreateWithFlags(&t->ev_join, cudaEventDisableTiming));
  t->in_events.resize(2 * (size_t)S + 2);
  for (auto& e : t->in_events) BT_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  t->host_debug = getenv("BT_HOST_DEBUG") != nullptr;
  t->no_refine = getenv("BT_NO_REFINE") != nullptr;
  t->ctrl_by_copy = getenv("BT_CTRL_COPY") != nullptr;
  t->no_prebuild = getenv("BT_NO_PREBUILD") != nullptr;
  t->copy_inline = getenv("BT_COPY_INLINE") != nullptr;
  t->no_l2_prefetch = getenv("BT_NO_L2_PREFETCH") != nullptr;
  t->direct_results = getenv("BT_DIRECT_RESULT") != nullptr;
  // measured at C3 (profiles/README.md): replaying the captured frame costs one ~23 us cudaGraphLaunch before the GPU
  // starts, the plain enqueue ~70 us of driver calls of which only the first ~25 us delay the GPU -- the plain
  // enqueue wins on this driver, and clearly so when a copy stream is busy next to it (pipelined ingest).  The graph
  // path stays available (BT_GRAPH=1) for hosts with slower launches.
  t->use_graph = getenv("BT_GRAPH") != nullptr && getenv("BT_NO_GRAPH") == nullptr;
  return BT_OK;
}

void bt_tracker_destroy(bt_ctx* ctx) {
  bt_tracker* t = ctx->trk;
  if (!t) return;
  bt_store& st = t->st;
  void* ptrs[] = {st.mean, st.cov, st.tlbr, st.tlbr_f32, st.feat16, st.norm, st.curr32, st.smooth32, st.slot_f32,
                  st.det_boxes, st.det_scores, st.det16, st.det32, st.det_norm, st.det_tlbr, st.det_xywh, st.det_xywh32,
                  st.col_kind, st.col_pk, st.ctrl, st.res, st.resB, st.y, st.pairs, t->d_birth, t->d_bpairs,
                  t->d_bpair_count, t->d_list, t->d_gather, t->d_gather32};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  for (auto& s : t->streams)
    if (s.face_dev) cudaFree(s.face_dev);
  void* hptrs[] = {t->h_ctrl, t->h_res, t->h_resB, t->h_flags, t->h_in_boxes, t->h_in_scores, t->h_birth, t->h_pairs, t->h_bpairs, t->h_bpair_count, t->h_list};
  for (void* p : hptrs)
    if (p) cudaFreeHost(p);
  if (t->ev_x) cudaEventDestroy(t->ev_x);
  if (t->ev_tail) cudaEventDestroy(t->ev_tail);
  if (t->ev_fork) cudaEventDestroy(t->ev_fork);
  if (t->ev_join) cudaEventDestroy(t->ev_join);
  for (auto& e : t->in_events)
    if (e) cudaEventDestroy(e);
  for (int s = 0; s < BT_SEG_HOST_ENQUEUE1; ++s)
    for (cudaEvent_t e : t->ev[s])
      if (e) cudaEventDestroy(e);
  for (auto& g : t->graphs) {
    if (g.exec) cudaGraphExecDestroy(g.exec);
    if (g.graph) cudaGraphDestroy(g.graph);
  }
  delete t;
  ctx->trk = nullptr;
}

// ------------------------------------------------------------------------------------------------
// submit: inputs of one frame -> the stream's free input half (copy stream)
// ------------------------------------------------------------------------------------------------
static int32_t check_batch(bt_ctx* ctx, int32_t count, const int32_t* sids) {
  bt_tracker* t = ctx->trk;
  BT_CHECK(count >= 1 && count <= BT_MAX_BATCH && sids, BT_ERR_INVALID, "batch of %d streams (1..%d)", count, BT_MAX_BATCH);
  for (int k = 0; k < count; ++k) {
    BT_CHECK(sids[k] >= 0 && sids[k] < t->S, BT_ERR_INVALID, "stream id %d out of range (ctx has %d)", sids[k], t->S);
    for (int j = 0; j < k; ++j) BT_CHECK(sids[j] != sids[k], BT_ERR_INVALID, "stream id %d listed twice", sids[k]);
  }
  return BT_OK;
}

static int32_t submit_batch(bt_ctx* ctx, int32_t count, const int32_t* sids, const int32_t* const* boxes,
                            const float* const* scores, const void* const* feats, const int32_t* m_arr,
                            int32_t dtype, const float* const* face_sims, int32_t loc) {
  bt_tracker* t = ctx->trk;
  bt_store& st = t->st;
  BT_TRY(check_batch(ctx, count, sids));
  BT_CHECK(loc == BT_HOST || loc == BT_DEVICE, BT_ERR_INVALID, "bad loc");
  BT_CHECK(dtype == BT_F32 || dtype == BT_F16, BT_ERR_INVALID, "feat_dtype must be BT_F32 or BT_F16");
  BT_CHECK(boxes && scores && m_arr, BT_ERR_INVALID, "NULL argument arrays");
  const size_t D = t->D, ND = (size_t)t->S * t->md;
  // validate everything before touching any state
  for (int k = 0; k < count; ++k) {
    StreamState& s = t->streams[sids[k]];
    const int m = m_arr[k];
    BT_CHECK(m >= 0 && m <= t->md, BT_ERR_CAPACITY, "%d detections exceed ctx max_dets %d", m, t->md);
    BT_CHECK(m == 0 || (boxes[k] && scores[k]), BT_ERR_INVALID, "NULL boxes/scores");
    const bool reid = s.cfg.with_reid != 0;
    BT_CHECK(!reid || m == 0 || (feats && feats[k]), BT_ERR_INVALID, "with_reid is set but feats is NULL");
    BT_CHECK(s.n_pending < 2, BT_ERR_STATE, "stream %d already has two submitted frames waiting for bt_step_streams", sids[k]);
    const bool has_tracks = !s.tracked.empty() || !s.lost.empty();
    BT_CHECK(!reid || m == 0 || s.feat_dtype < 0 || !has_tracks || s.feat_dtype == dtype, BT_ERR_STATE,
             "stream %d holds %s features: the feature dtype cannot change while tracks exist", sids[k],
             s.feat_dtype == BT_F16 ? "fp16" : "fp32");
    BT_CHECK(!(face_sims && face_sims[k] && m > 0) || s.n_pending == 0, BT_ERR_STATE,
             "face similarities need the previous frame of stream %d stepped first", sids[k]);
  }
  bool any_f32 = false;
  for (int k = 0; k < count; ++k)
    any_f32 = any_f32 || (dtype == BT_F32 && t->streams[sids[k]].cfg.with_reid && m_arr[k] > 0);
  if (any_f32 && !st.det32) {   // fp32 ingest staging + the fp32 current-feature bank: allocated on first use
    BT_CUDA(cudaStreamSynchronize(ctx->stream));
    BT_TRY(dev_alloc(ctx, &st.det32, 2 * ND * D));
    BT_TRY(dev_alloc(ctx, &st.curr32, (size_t)t->S * t->cap * D));
    BT_CUDA(cudaMemset(st.curr32, 0, sizeof(float) * (size_t)t->S * t->cap * D));
  }
  cudaStream_t cs = ctx->copy_stream;
  if (t->tail_pending) {    // births of the last step may still be reading the input half we are about to refill
    BT_CUDA(cudaStreamWaitEvent(cs, t->ev_tail, 0));
    t->tail_pending = false;
  }
  bool copied = false;
  for (int k = 0; k < count; ++k) {
    const int sid = sids[k];
    StreamState& s = t->streams[sid];
    const int m = m_arr[k];
    const int parity = s.next_parity;
    s.next_parity ^= 1;
    if (parity == s.parity) s.boxes_valid = false;   // the half the last stepped frame came from is being refilled
    FrameIn& in = s.in[parity];
    in = FrameIn();
    in.m = m; in.dtype = dtype; in.loc = loc;
    const size_t g0 = (size_t)parity * ND + (size_t)sid * t->md;
    const bool reid = s.cfg.with_reid != 0;
    const cudaMemcpyKind kind = loc == BT_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
    if (m > 0) {
      int32_t* db = st.det_boxes + g0 * 4;
      float* ds = st.det_scores + g0;
      const int32_t* src_b = boxes[k];
      const float* src_s = scores[k];
      if (loc == BT_HOST) {   // pinned copies: truly asynchronous H2D, and the host-side bookkeeping owns what it reads
        memcpy(t->h_in_boxes + g0 * 4, boxes[k], sizeof(int32_t) * 4 * m);
        memcpy(t->h_in_scores + g0, scores[k], sizeof(float) * m);
        src_b = t->h_in_boxes + g0 * 4;
        src_s = t->h_in_scores + g0;
      }
      if (src_b != db) { BT_CUDA(cudaMemcpyAsync(db, src_b, sizeof(int32_t) * 4 * m, kind, cs)); copied = true; }
      if (src_s != ds) { BT_CUDA(cudaMemcpyAsync(ds, src_s, sizeof(float) * m, kind, cs)); copied = true; }
      if (reid) {
        if (dtype == BT_F16) {
          __half* d16 = st.det16 + g0 * D;
          if (feats[k] != d16) { BT_CUDA(cudaMemcpyAsync(d16, feats[k], sizeof(__half) * (size_t)m * D, kind, cs)); copied = true; }
        } else {
          BT_CUDA(cudaMemcpyAsync(st.det32 + g0 * D, feats[k], sizeof(float) * (size_t)m * D, kind, cs));
          copied = true;
        }
      }
      if (loc == BT_HOST) { in.host_boxes = t->h_in_boxes + g0 * 4; in.host_scores = t->h_in_scores + g0; }
    }
    if (face_sims && face_sims[k] && m > 0) {
      // rows in pool order: activated tracked tracks then lost tracks (the pool of THIS step; no frame of
      // this stream may be pending, or the pool would not be known yet)
      int n_pool = (int)s.lost.size();
      for (int slot : s.tracked) n_pool += s.meta[slot].activated ? 1 : 0;
      const size_t need = (size_t)n_pool * m;
      if (need > 0) {
        if (need > s.face_cap) {
          BT_CUDA(cudaStreamSynchronize(ctx->stream));
          if (s.face_dev) BT_CUDA(cudaFree(s.face_dev));
          s.face_dev = nullptr;
          BT_CUDA(cudaMalloc(&s.face_dev, sizeof(float) * need * 2));
          s.face_cap = need * 2;
        }
        BT_CUDA(cudaMemcpyAsync(s.face_dev, face_sims[k], sizeof(float) * need, kind, cs));
        copied = true;
        in.face_sim = s.face_dev;
      }
    }
    s.pending[s.n_pending++] = parity;
  }
  if (copied) {
    const int e = t->in_seq;
    t->in_seq = (t->in_seq + 1) % (int)t->in_events.size();
    BT_CUDA(cudaEventRecord(t->in_events[e], cs));
    for (int k = 0; k < count; ++k) {
      StreamState& s = t->streams[sids[k]];
      s.in[s.pending[s.n_pending - 1]].ev = e;
    }
  }
  return BT_OK;
}

// ------------------------------------------------------------------------------------------------
// A steady-state frame issues ~14 driver calls (copies, 6-7 kernel launches, events); on the host that is
// 60-90 us -- longer than the GPU needs for the work.  The same sequence is therefore captured ONCE per
// launch shape into a CUDA graph (programmatic-dependent-launch edges included); every later frame records
// its kernels' geometry + arguments (bt_launch_rec), patches the instantiated graph's kernel nodes with them
// and launches the graph: one launch + one cheap update per kernel.
// ------------------------------------------------------------------------------------------------
template <typename Enqueue>
static int32_t run_enqueue(bt_ctx* ctx, bt_tracker* t, Enqueue& enqueue, const GraphKey& key, bool eligible,
                           bool& part_a_sent, bool& ema_pending) {
  cudaStream_t st = ctx->stream;
  if (!eligible) return enqueue(0);
  FrameGraph* g = nullptr;
  for (auto& cand : t->graphs)
    if (cand.key == key) g = &cand;
  if (!g) {
    bool seen = false;
    for (const auto& k : t->seen_keys) seen = seen || k == key;
    if (!seen) {       // first frame of this shape: the plain way (first-launch work must not happen under capture)
      t->seen_keys.push_back(key);
      return enqueue(0);
    }
    FrameGraph fg;
    fg.key = key;
    const int64_t launches0 = ctx->launches;
    cudaError_t e = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
    int32_t rc = BT_OK;
    if (e == cudaSuccess) {
      rc = enqueue(1);
      cudaError_t e2 = cudaStreamEndCapture(st, &fg.graph);
      if (rc == BT_OK && e2 != cudaSuccess) e = e2;
    }
    fg.n_kernels = (int)(ctx->launches - launches0);
    ctx->launches = launches0;
    if (e == cudaSuccess && rc == BT_OK) e = cudaGraphInstantiate(&fg.exec, fg.graph, 0);
    if (e != cudaSuccess || rc != BT_OK) {
      // graphs are an optimisation: fall back to the plain enqueue for good
      (void)cudaGetLastError();
      if (fg.exec) cudaGraphExecDestroy(fg.exec);
      if (fg.graph) cudaGraphDestroy(fg.graph);
      t->use_graph = false;
      if (getenv("BT_HOST_DEBUG")) fprintf(stderr, "botsort_b200: frame graph capture failed (%s), using plain launches\n", cudaGetErrorString(e));
      part_a_sent = false; ema_pending = false;
      return enqueue(0);
    }
    t->graphs.push_back(fg);
    g = &t->graphs.back();
  }
  // the kernels read this frame's sizes / offsets from the launch description that the graph's first node copies
  // up: nothing to patch
  BT_CUDA(cudaGraphLaunch(g->exec, st));
  ctx->launches += g->n_kernels;
  part_a_sent = true;
  ema_pending = false;
  return BT_OK;
}

// ------------------------------------------------------------------------------------------------
// the frame step of a batch of video streams
// ------------------------------------------------------------------------------------------------
static int32_t step_batch(bt_ctx* ctx, int32_t count, const int32_t* sids, bt_frame_info* infos) {
  bt_tracker* t = ctx->trk;
  bt_store& dst = t->st;
  const bt_res_layout& L = t->L;
  BT_TRY(check_batch(ctx, count, sids));
  for (int k = 0; k < count; ++k)
    BT_CHECK(t->streams[sids[k]].n_pending > 0, BT_ERR_STATE, "stream %d has no submitted frame", sids[k]);
  cudaStream_t st = ctx->stream;
  const int D = t->D;
  double t_host = now_ms();
  double hphase[5] = {0, 0, 0, 0, 0};
  const char* fm_name[40]; double fm_t[40]; int fm_n = 0;
  auto fmark = [&](const char* nm) { if (t->host_debug && fm_n < 40) { fm_name[fm_n] = nm; fm_t[fm_n++] = now_ms(); } };
  fmark("start");

  // ---- per-stream set-up: pop the oldest frame, split lists (demo:1415-1423), control segments ----
  FrameDesc* hd = reinterpret_cast<FrameDesc*>(t->h_ctrl);
  const FrameDesc* dd = reinterpret_cast<const FrameDesc*>(dst.ctrl);
  bt_batch& B = hd->B;
  memset(&B, 0, sizeof(B));
  B.count = count;
  bool any_reid = false, any_dev_inputs = false, any_f32 = false, any_f16 = false, any_face = false;
  int waited[BT_MAX_BATCH];
  int n_waited = 0;
  size_t ctrl_off = kDescBytes;
  for (int k = 0; k < count; ++k) {
    const int sid = sids[k];
    StreamState& s = t->streams[sid];
    const int parity = s.pending[0];
    s.pending[0] = s.pending[1];
    s.n_pending -= 1;
    const FrameIn& in = s.in[parity];
    s.parity = parity;
    s.m = in.m;
    s.device_inputs = in.loc == BT_DEVICE;
    s.frame_id += 1;  // demo:1292
    s.tlbr_cache_valid = false;
    t->region_gen[k] += 1;
    const bool reid = s.cfg.with_reid != 0;
    if (reid && in.m > 0) {
      s.feat_dtype = in.dtype;
      (in.dtype == BT_F16 ? any_f16 : any_f32) = true;
    }
    any_reid = any_reid || (reid && in.m > 0);
    any_dev_inputs = any_dev_inputs || s.device_inputs;
    if (in.ev >= 0) {
      bool seen = false;
      for (int j = 0; j < n_waited; ++j) seen = seen || waited[j] == in.ev;
      if (!seen) { BT_CUDA(cudaStreamWaitEvent(st, t->in_events[in.ev], 0)); waited[n_waited++] = in.ev; }
    }
    std::vector<SlotMeta>& meta = s.meta;
    std::vector<int>& unconfirmed = s.v_unconfirmed; unconfirmed.clear();
    // pool = activated tracked tracks (list order) then lost tracks (joint_stracks, demo:1423: ids are unique per
    // slot, no overlap): written straight into the control segment, which the list bookkeeping reads back later
    if (s.pre_valid) s.n_pool = s.pre_n_pool;
    else {
      int n_act = 0;
      for (int slot : s.tracked) n_act += meta[slot].activated ? 1 : 0;
      s.n_pool = n_act + (int)s.lost.size();
    }
    s.n_rows = s.high_water;
    const bool face = in.face_sim != nullptr && s.n_pool > 0;
    any_face = any_face || face;
    char* seg = t->h_ctrl + ctrl_off;
    int32_t* h_idx = reinterpret_cast<int32_t*>(seg);
    int32_t* h_state = h_idx + s.n_pool;
    uint8_t* h_kind = reinterpret_cast<uint8_t*>(h_state + s.n_pool);
    int32_t* h_pos = reinterpret_cast<int32_t*>(h_kind + (((size_t)s.n_rows + 3) & ~size_t(3)));
    bool all_f32 = s.n_pool > 0;
    if (s.pre_valid && !face && s.pre_n_rows == s.n_rows && s.pre_n_pool == s.n_pool) {
      // built at the end of the previous step (no pooled track has float32 state there: all_f32 is false)
      if (s.n_pool > 0) {
        memcpy(h_idx, s.pre_idx.data(), sizeof(int32_t) * (size_t)s.n_pool);
        memcpy(h_state, s.pre_state.data(), sizeof(int32_t) * (size_t)s.n_pool);
      }
      if (s.n_rows > 0) memcpy(h_kind, s.pre_kind.data(), ((size_t)s.n_rows + 3) & ~size_t(3));
      unconfirmed.swap(s.pre_unconfirmed);
      all_f32 = false;
    } else {
      memset(h_kind, BT_ROW_NONE, ((size_t)s.n_rows + 3) & ~size_t(3));
      if (face) for (int r = 0; r < s.n_rows; ++r) h_pos[r] = -1;
      int np = 0;
      auto add_pool = [&](int slot) {
        SlotMeta& tm = meta[slot];
        h_idx[np] = slot;
        h_state[np] = tm.state;
        h_kind[slot] = (tm.state == BT_STATE_TRACKED) ? BT_ROW_POOL_TRACKED : BT_ROW_POOL_OTHER;
        if (face) h_pos[slot] = np;
        all_f32 = all_f32 && tm.f32_state;
        tm.f32_state = 0;     // predicted below: float64 from now on
        ++np;
      };
      for (int slot : s.tracked) {
        if (!meta[slot].activated) { unconfirmed.push_back(slot); h_kind[slot] = BT_ROW_UNCONFIRMED; }
        else add_pool(slot);
      }
      for (int slot : s.lost) add_pool(slot);
    }
    s.pre_valid = false;
    s.pool = h_idx;
    s.n_unc = (int)unconfirmed.size();
    B.sid[k] = sid;
    B.m[k] = s.m;
    B.n_rows[k] = s.n_rows;
    B.n_pool[k] = s.n_pool;
    B.ctrl_off[k] = (int32_t)ctrl_off;
    B.noise_f32[k] = all_f32 ? 1 : 0;
    B.parity[k] = (uint8_t)parity;
    B.want_norm[k] = (reid && s.n_unc > 0 && s.m > 0 && in.dtype == BT_F16) ? 1 : 0;
    ctrl_off += (ctrl_bytes(s.n_pool, s.n_rows, face) + 15) & ~size_t(15);
    // host views of this frame's scores / boxes and result regions
    const int32_t* hresA = reinterpret_cast<const int32_t*>(t->h_res + (size_t)k * L.stride);
    for (int q = 0; q < 3; ++q) s.hx[q] = hresA + L.o_x + (size_t)q * t->cap;
    s.sc = s.device_inputs ? reinterpret_cast<const float*>(hresA + L.o_sc) : in.host_scores;
    s.host_boxes = s.device_inputs ? hresA + L.o_bx : in.host_boxes;
  }
  fmark("streams_setup");
  BT_CHECK(!(any_f32 && any_f16), BT_ERR_INVALID, "one batch cannot mix fp16 and fp32 feature streams");
  const bool f16 = any_f16;
  const uint32_t flags = ctx->flags;
  const bool tensor_path = !(flags & BT_FLAG_SIMT_SIM) && (D % 64 == 0) && D >= 512;
  const StreamState& s0 = t->streams[sids[0]];
  const bt_config& cfg = s0.cfg;     // thresholds of a batch are those of its first stream (bt_tracker_reset applies one cfg to all by default)
  bt_frame_cfg fc;
  fc.high = cfg.track_high_thresh; fc.low = cfg.track_low_thresh;
  fc.alpha = (float)cfg.ema_alpha; fc.one_minus_alpha = (float)(1.0 - cfg.ema_alpha);
  fc.dup_limit = cfg.duplicate_iou_dist;
  fc.device_inputs = any_dev_inputs ? 1 : 0;
  fc.f16_inputs = f16 ? 1 : 0;
  fc.keep_smooth = dst.smooth32 != nullptr;
  fc.prefetch_pairs = kPairPrefetch;
  fc.with_reid = any_reid ? 1 : 0;
  fc.l2_prefetch = (any_reid && tensor_path && !t->no_l2_prefetch) ? 1 : 0;

  const int mx_rows = bt_batch_max(B.n_rows, count), mx_m = bt_batch_max(B.m, count);
  // Direct results (opt-in, BT_DIRECT_RESULT=1): the LAP kernel stores the assignment vectors into the pinned result
  // block itself and raises a flag the host polls -- no event record / cross-stream wait / D2H copy / event between
  // the LAP kernel and the host.  Measured on the C3 step: 0.103 ms against 0.098 ms with the side-stream copy (the
  // SM-issued PCIe stores + system fence at the end of the single LAP CTA cost more than the copy engine's round trip).
  const bool direct = t->direct_results && !t->use_graph && mx_rows > 0;
  dst.res_host = direct ? t->h_res_dev : nullptr;
  if (direct) for (int k = 0; k < count; ++k) t->h_flags[k] = 0u;
  const int assoc_bn = (any_reid && tensor_path) ? btk_assoc_pick_bn(ctx, B.n_rows, B.m, count) : 256;
  const int assoc_precision = (any_reid && tensor_path) ? 0 : 1;
  const size_t ND = (size_t)t->S * t->md;
  bt_cand cand = *bt_lap_own_cand(ctx);
  cand.seg = assoc_bn / 2;    // one epilogue thread owns one (row, segment) pair; the LAP gather follows

  // ---- the frame's launch description: association problems + LAP batch (host copy; goes up with the control block) ----
  bt_lap_batch& LB = hd->LB;
  memset(&LB, 0, sizeof(LB));
  LB.count = count;
  LB.y_stride = t->md;
  bt_assoc_params p;
  memset(&p, 0, sizeof(p));
  p.count = count;
  for (int k = 0; k < count; ++k) {
    const int sid = sids[k];
    const StreamState& s = t->streams[sid];
    const FrameIn& in = s.in[s.parity];
    const bool face = in.face_sim != nullptr && s.n_pool > 0;
    const int pos_off = B.ctrl_off[k] + 8 * s.n_pool + (int)(((size_t)s.n_rows + 3) & ~size_t(3));
    p.n[k] = s.n_rows; p.m[k] = s.m;
    p.a_row0[k] = sid * t->cap;
    p.b_row0[k] = (int32_t)((size_t)s.parity * ND + (size_t)sid * t->md);
    p.row0[k] = sid * t->cap; p.col0[k] = sid * t->md;
    p.kind_off[k] = B.ctrl_off[k] + 8 * s.n_pool;
    p.pos_off[k] = pos_off;
    p.cand_sid[k] = sid;
    p.face_sim[k] = face ? in.face_sim : nullptr;
    LB.sid[k] = sid; LB.n[k] = s.n_rows; LB.m[k] = s.m;
    LB.row0[k] = p.row0[k]; LB.col0[k] = p.col0[k]; LB.in0[k] = p.b_row0[k];
    LB.pos_off[k] = pos_off;
    LB.face_sim[k] = p.face_sim[k];
    int32_t* dresA = reinterpret_cast<int32_t*>(dst.res + (size_t)k * L.stride);
    LB.x[k] = dresA + L.o_x; LB.x_stride[k] = t->cap;
    LB.y[k] = dst.y + (size_t)k * 3 * t->md;
    LB.zero_word[k] = reinterpret_cast<int32_t*>(dst.resB + (size_t)k * L.strideB) + L.o_hdr;
    LB.hx[k] = direct ? reinterpret_cast<int32_t*>(t->h_res_dev + (size_t)k * L.stride) + L.o_x : nullptr;
    LB.hflag[k] = direct ? t->h_flags_dev + k : nullptr;
  }
  p.d = any_reid ? D : 0;
  p.a_rows_alloc = t->S * t->cap; p.b_rows_alloc = (int32_t)(2 * ND);
  p.bn = assoc_bn;
  // both operands are complete before the prep kernel starts (the bank since the last frame, the
  // detection features since the copy / cast ahead of it): the main loop may run under it
  p.operands_early = 1;
  if (assoc_precision == 0 || (any_reid && f16)) {
    p.a16 = dst.feat16; p.b16 = dst.det16; p.row_norm = dst.norm;
  } else if (any_reid) {
    p.a32 = dst.curr32; p.b32 = dst.det32;      // normalised fp32 current features x raw fp32 detections
  }
  p.row_tlbr = dst.tlbr; p.row_tlbr_f32 = dst.tlbr_f32; p.row_kind_base = dst.ctrl;
  p.col_tlbr = dst.det_tlbr; p.col_kind = dst.col_kind; p.col_pk = dst.col_pk;
  // detection feature norms (stage 3 compares normalised detections): always there with fp32 ingest, computed on
  // demand (streams with unconfirmed rows) with fp16 ingest; nobody else looks at them
  p.col_norm = any_reid ? dst.det_norm : nullptr;
  p.match_thresh = cfg.match_thresh; p.second_thresh = cfg.second_thresh;
  p.unconf_thresh = cfg.unconfirmed_thresh; p.proximity = cfg.proximity_thresh;
  p.appearance = (float)cfg.appearance_thresh;
  // fp32 ingest rounds the operands to fp16 (3e-5 on unit 2048-d rows, SURVEY hard part 2); fp16 ingest
  // multiplies the reference's own values exactly and only the accumulation order differs
  p.gate_band = f16 ? 1.0e-4f : 5.0e-4f;
  p.cand = cand;
  btk_assoc_fill(p, assoc_precision, &hd->AF);
  bt_refine rf;
  memset(&rf, 0, sizeof(rf));
  if (any_reid && tensor_path && !t->no_refine) {
    rf.enabled = 1; rf.d = D; rf.f16 = f16 ? 1 : 0;
    rf.a16 = dst.feat16; rf.a_norm = dst.norm; rf.a32 = dst.curr32;
    rf.b16 = dst.det16; rf.b32 = dst.det32;
    rf.b_norm = p.col_norm;
    rf.row_tlbr = dst.tlbr; rf.col_tlbr = dst.det_tlbr; rf.ctrl = dst.ctrl;
    rf.proximity = cfg.proximity_thresh; rf.appearance = (float)cfg.appearance_thresh;
  }
  if (mx_rows > 0 && mx_m > 0) {
    t->last_assoc = p;
    t->last_AF = hd->AF;
    t->last_assoc_precision = assoc_precision;
    t->last_assoc_valid = true;
  }
  fmark("setup");

  bool part_a_sent = false, ema_pending = false;
  // fixed == 0: plain enqueue with the frame's own launch geometry; fixed == 1: the same calls under stream
  // capture, with capacity-sized geometry and copies (what a graph that is replayed for every frame needs) and the
  // side stream joined back
  auto enqueue = [&](const int fixed) -> int32_t {
    const size_t ctrl_bytes_now = fixed ? kDescBytes + t->ctrl_stride * (size_t)count : ctrl_off;
    if (t->ctrl_by_copy) BT_CUDA(cudaMemcpyAsync(dst.ctrl, t->h_ctrl, ctrl_bytes_now, cudaMemcpyHostToDevice, st));
    else BT_TRY(btk_ctrl_upload(ctx, t->h_ctrl_dev, dst.ctrl, ctrl_bytes_now));
    fmark("ctrl_h2d");
    SEG_BEGIN(BT_SEG_PREP);
    if (any_f32) BT_TRY(btk_frame_cast(ctx, dst, B, &dd->B, fixed));
    SEG_END(BT_SEG_PREP);
    SEG_BEGIN(BT_SEG_PREDICT);
    BT_TRY(btk_frame_prep(ctx, dst, B, &dd->B, fc, fixed));
    SEG_END(BT_SEG_PREDICT);
    fmark("prep");
    if (mx_rows > 0 || fixed) {
      // the candidate counters are left zeroed by the previous frame's LAP kernel (no memset here)
      SEG_BEGIN(BT_SEG_ASSOC);
      if (mx_m > 0 || fixed) BT_TRY(btk_assoc_launch(ctx, p, assoc_precision, hd->AF, &dd->AF, fixed, t->cap, t->md));
      SEG_END(BT_SEG_ASSOC);
      fmark("assoc");
      SEG_BEGIN(BT_SEG_LAP);
      const double th[3] = {cfg.match_thresh, cfg.second_thresh, cfg.unconfirmed_thresh};
      BT_TRY(btk_lap_solve3(ctx, cand, LB, &dd->LB, th, rf, assoc_precision != 0));   // also zeroes the pair counters
      SEG_END(BT_SEG_LAP);
      fmark("lap");
    }
    // part A of the result regions: the assignment vectors (+ echoed scores / boxes of device inputs)
    const size_t widthA = sizeof(int32_t) * (any_dev_inputs ? L.o_endA : L.o_sc);
    if (!(direct && !fixed) && (mx_rows > 0 || (any_dev_inputs && mx_m > 0) || fixed)) {
      // plain enqueue: the read-back runs on the side stream, so that the Kalman update behind the LAP kernel does not
      // queue behind a copy-engine round trip (it becomes a programmatic dependent of the LAP kernel instead)
      const bool aside = !fixed && !t->copy_inline;
      cudaStream_t cs = aside ? ctx->side_stream : st;
      if (aside) {
        BT_CUDA(cudaEventRecord(t->ev_fork, st));
        BT_CUDA(cudaStreamWaitEvent(cs, t->ev_fork, 0));
      }
      if (count == 1) BT_CUDA(cudaMemcpyAsync(t->h_res, dst.res, widthA, cudaMemcpyDeviceToHost, cs));
      else BT_CUDA(cudaMemcpy2DAsync(t->h_res, L.stride, dst.res, L.stride, widthA, count, cudaMemcpyDeviceToHost, cs));
      // under capture: an event RECORD NODE the host can wait on (a plain record would only order captured work)
      if (fixed) BT_CUDA(cudaEventRecordWithFlags(t->ev_x, st, cudaEventRecordExternal));
      else BT_CUDA(cudaEventRecord(t->ev_x, cs));
      part_a_sent = true;
      fmark("copyA+ev");
    }
    if (mx_rows > 0 || fixed) {
      // The matched tracks' Kalman update and feature EMA are a pure function of the three assignment
      // vectors, so they run on the device straight away (STrack.update / re_activate arithmetic,
      // demo:570-610) while the host is still waiting for / digesting the assignments.
      SEG_BEGIN(BT_SEG_UPDATE);
      BT_TRY(btk_frame_post(ctx, dst, B, &dd->B, fc, fixed, (!fixed && (direct || (!t->copy_inline && part_a_sent))) ? 1 : 0));
      SEG_END(BT_SEG_UPDATE);
      fmark("post");
      // duplicate candidates among all live slots (superset of tracked x lost) + every slot's box
      SEG_BEGIN(BT_SEG_DUP);
      // ... and, in the same launch, the matched tracks' feature update (it only needs the assignment vectors: its
      // CTAs run beside the Kalman update and the duplicate test)
      BT_TRY(btk_frame_dup(ctx, dst, B, &dd->B, fc, fixed, (any_reid && (mx_m > 0 || fixed)) ? 1 : 0));
      SEG_END(BT_SEG_DUP);
      fmark("dup");
      // part B: pair count + first pairs, boxes of all slots
      const size_t widthB = L.o_tlbr_bytes + sizeof(double) * 4 * (size_t)(fixed ? t->cap : mx_rows);
      if (count == 1) BT_CUDA(cudaMemcpyAsync(t->h_resB, dst.resB, widthB, cudaMemcpyDeviceToHost, st));
      else BT_CUDA(cudaMemcpy2DAsync(t->h_resB, L.strideB, dst.resB, L.strideB, widthB, count, cudaMemcpyDeviceToHost, st));
      if (fixed && ema_pending) {     // a captured fork must rejoin its origin stream
        BT_CUDA(cudaStreamWaitEvent(st, t->ev_join, 0));
        ema_pending = false;
      }
    }
    return BT_OK;
  };
  {
    GraphKey key = {count, any_f32 ? 1 : 0, any_reid ? 1 : 0, any_dev_inputs ? 1 : 0, assoc_bn, assoc_precision, f16 ? 1 : 0};
    const bool eligible = t->use_graph && !t->prof && mx_rows > 0 && mx_m > 0;
    BT_TRY(run_enqueue(ctx, t, enqueue, key, eligible, part_a_sent, ema_pending));
  }
  fmark("copyB");
  HOST_MARK(BT_SEG_HOST_ENQUEUE1);
  // wait only for the assignments (+ scores / boxes): the list bookkeeping below overlaps the GPU's
  // update / EMA / duplicate-test tail
  fmark("launch");
  if (part_a_sent) BT_CUDA(cudaEventSynchronize(t->ev_x));
  else if (direct) {
    // poll the streams' flags; every few thousand polls make sure the stream is still alive (a faulting kernel must
    // not leave the host spinning)
    for (int k = 0; k < count; ++k) {
      volatile uint32_t* f = t->h_flags + k;
      uint32_t polls = 0;
      int finished_checks = 0;
      while (*f == 0u) {
#if defined(__x86_64__) || defined(__i386__)
        __builtin_ia32_pause();
#endif
        if ((++polls & 4095u) == 0u) {
          const cudaError_t q = cudaStreamQuery(st);
          if (q == cudaSuccess) {
            BT_CHECK(++finished_checks < 64 || *f != 0u, BT_ERR_STATE, "the frame's kernels finished without publishing stream %d's assignments", sids[k]);
          } else if (q != cudaErrorNotReady) {
            BT_CUDA(q);
          }
        }
      }
    }
    std::atomic_thread_fence(std::memory_order_acquire);
  }
  HOST_MARK(BT_SEG_HOST_WAIT1);
  fmark("wait_x");

  // ---- per-stream list bookkeeping (demo:1493-1636) -------------------------------------------------
  for (int k = 0; k < count; ++k) {
    StreamState& s = t->streams[sids[k]];
    SlotMeta* meta = s.meta.data();
    const bt_config& c = s.cfg;
    const int m = s.m, n_pool = s.n_pool, n_unc = s.n_unc, n_rows = s.n_rows, frame_id = s.frame_id;
    const float* sc = s.sc;
    const int32_t* pool = s.pool;
    const std::vector<int>& unconfirmed = s.v_unconfirmed;
    std::vector<uint8_t>& det_taken = s.v_det_taken; det_taken.assign(m, 0);
    std::vector<int>& refind = s.v_refind; refind.clear();
    std::vector<int>& lost_now = s.v_lost_now; lost_now.clear();
    std::vector<int>& removed_now = s.v_removed_now; removed_now.clear();
    std::vector<int>& r_tracked = s.v_r_tracked; r_tracked.clear();
    const int32_t* hx0 = s.hx[0]; const int32_t* hx1 = s.hx[1]; const int32_t* hx2 = s.hx[2];
    s.n_m1 = s.n_m2 = s.n_m3 = 0;
    // STrack.update (demo:586-610) / STrack.re_activate (demo:570-584) bookkeeping of a matched slot.  The order
    // of the reference's activated / refind lists only matters for tracks that are not in tracked_stracks yet:
    // re-found lost tracks (first association only: ascending pool index) and births.
    auto apply_match = [&](int slot, int det) {
      SlotMeta& tm = meta[slot];
      tm.f32_state = 0;
      if (tm.state == BT_STATE_TRACKED) tm.tracklet_len += 1;
      else { tm.tracklet_len = 0; refind.push_back(slot); }
      tm.frame_id = frame_id;
      tm.state = BT_STATE_TRACKED;
      tm.activated = 1;
      tm.score = sc[det];
      tm.det_index = det;
      det_taken[det] = 1;
    };
    // first association (demo:1556-1566).  Stage 2's candidates (demo:1569: unmatched pool tracks in state Tracked)
    // are collected in the same pass: a stage-1 update never touches an unmatched track.
    if (n_rows > 0) {
      for (int i = 0; i < n_pool; ++i) {
        const int slot = pool[i];
        const int j = hx0[slot];
        if (j >= 0) { apply_match(slot, j); s.n_m1 += 1; }
        else if (meta[slot].state == BT_STATE_TRACKED) r_tracked.push_back(slot);
      }
    }
    if (k == 0) fmark("L:stage1");
    // second association (demo:1568-1586)
    for (int slot : r_tracked) {
      const int j = hx1[slot];
      if (j >= 0) { apply_match(slot, j); s.n_m2 += 1; }
      else { meta[slot].state = BT_STATE_LOST; lost_now.push_back(slot); }   // mark_lost
    }
    // unconfirmed (demo:1588-1612)
    for (int slot : unconfirmed) {
      const int j = hx2[slot];
      if (j >= 0) { apply_match(slot, j); s.n_m3 += 1; }
      else { meta[slot].state = BT_STATE_REMOVED; removed_now.push_back(slot); }   // mark_removed, demo:1609-1612
    }
    if (k == 0) fmark("L:stage23");
    // births (demo:1614-1621, STrack.activate demo:556-568): unmatched high detections (demo:1501: float(score) >
    // 0.40 as Python doubles) with score >= new_track_thresh, in detection order.  The reference's track store is
    // unbounded; here a full store drops the birth (the detection stays unmatched, as if its score were below
    // new_track_thresh) and the frame stays consistent: nothing else depends on it.
    s.v_birth_slot.clear(); s.v_birth_det.clear();
    s.n_births_skipped = 0;
    int n_high = 0, n_low = 0;
    for (int j = 0; j < m; ++j) {
      const double sd = (double)sc[j];
      if (sd > c.track_high_thresh) {
        n_high += 1;
        if (det_taken[j] || sd < c.new_track_thresh) continue;
        const int slot = alloc_slot(s, t->cap);
        if (slot < 0) { s.n_births_skipped += 1; continue; }
        SlotMeta& tm = meta[slot];
        tm = SlotMeta();
        tm.used = 1;
        tm.track_id = ++s.id_count;
        tm.state = BT_STATE_TRACKED;
        tm.activated = (frame_id == 1) ? 1 : 0;
        tm.frame_id = frame_id;
        tm.start_frame = frame_id;
        tm.tracklet_len = 0;
        tm.score = sc[j];
        tm.det_index = j;
        tm.f32_state = 1;
        s.v_birth_slot.push_back(slot);
        s.v_birth_det.push_back(j);
      } else if (sd >= c.track_low_thresh) {
        n_low += 1;
      }
    }
    s.n_high = n_high; s.n_low = n_low;
    s.n_births = (int)s.v_birth_slot.size();
    // expiry (demo:1623-1627)
    for (int slot : s.lost) {
      if (frame_id - meta[slot].frame_id > s.max_time_lost) {
        meta[slot].state = BT_STATE_REMOVED;
        removed_now.push_back(slot);
      }
    }
    if (k == 0) fmark("L:births");
    // merge lists (demo:1629-1636): tracked tracks that are still Tracked keep their positions, then births
    // (detection order), then re-found tracks (pool order)
    std::vector<int>& new_tracked = s.v_new_tracked; new_tracked.clear();
    for (int slot : s.tracked)
      if (meta[slot].state == BT_STATE_TRACKED) new_tracked.push_back(slot);
    for (int slot : s.v_birth_slot) new_tracked.push_back(slot);
    for (int slot : refind) { new_tracked.push_back(slot); meta[slot].mark = 1; }
    std::vector<int>& new_lost = s.v_new_lost; new_lost.clear();
    for (int slot : s.lost)
      if (!meta[slot].mark && !meta[slot].in_removed) new_lost.push_back(slot);   // sub_stracks(lost, tracked), then sub_stracks(lost, removed) BEFORE this frame's removals
    for (int slot : lost_now)
      if (!meta[slot].in_removed) new_lost.push_back(slot);                        // extend(lost_stracks)
    for (int slot : refind) meta[slot].mark = 0;
    for (int slot : removed_now) meta[slot].in_removed = 1; // removed_stracks.extend
    s.n_removed_total += (int)removed_now.size();
    s.v_pool.assign(pool, pool + n_pool);   // the control block is rewritten by the next step; diagnostics may ask later
    s.matches_valid = true;
  }
  fmark("L:merge");
  HOST_MARK(BT_SEG_HOST_LISTS);
  // ---- the next frame's pool lists, while the GPU finishes this frame (Kalman update / EMA / duplicate test) ----
  for (int k = 0; k < count && !t->no_prebuild; ++k) {
    StreamState& s = t->streams[sids[k]];
    const SlotMeta* meta = s.meta.data();
    const int n_rows = s.high_water;
    s.pre_kind.assign(((size_t)n_rows + 3) & ~size_t(3), (uint8_t)BT_ROW_NONE);
    s.pre_idx.clear(); s.pre_state.clear(); s.pre_unconfirmed.clear();
    bool any_f32 = false;
    auto add_pool = [&](int slot) {
      const SlotMeta& tm = meta[slot];
      s.pre_idx.push_back(slot);
      s.pre_state.push_back(tm.state);
      s.pre_kind[slot] = (tm.state == BT_STATE_TRACKED) ? BT_ROW_POOL_TRACKED : BT_ROW_POOL_OTHER;
      any_f32 = any_f32 || tm.f32_state;
    };
    for (int slot : s.v_new_tracked) {
      if (!meta[slot].activated) { s.pre_unconfirmed.push_back(slot); s.pre_kind[slot] = BT_ROW_UNCONFIRMED; }
      else add_pool(slot);
    }
    for (int slot : s.v_new_lost) add_pool(slot);
    s.pre_n_pool = (int)s.pre_idx.size();
    s.pre_n_rows = n_rows;
    s.pre_valid = !any_f32;
  }
  fmark("prebuild");
  BT_CUDA(cudaStreamSynchronize(st));    // the frame's device work is complete, part B is on the host
  if (ema_pending) BT_CUDA(cudaEventSynchronize(t->ev_join));   // the side stream's EMA (usually done already)
  HOST_MARK(BT_SEG_HOST_WAIT2);
  BT_TRY(prof_collect(ctx, t));

  // ---- per-stream finish: births on the device, remove_duplicate_stracks, slot recycling ----------------
  for (int k = 0; k < count; ++k) {
    const int sid = sids[k];
    StreamState& s = t->streams[sid];
    std::vector<SlotMeta>& meta = s.meta;
    const bt_config& c = s.cfg;
    const int frame_id = s.frame_id, n_births = s.n_births;
    std::vector<int>& new_tracked = s.v_new_tracked;
    std::vector<int>& new_lost = s.v_new_lost;
    const int nt = (int)new_tracked.size(), nl = (int)new_lost.size();
    const int32_t* hresB = reinterpret_cast<const int32_t*>(t->h_resB + (size_t)k * L.strideB);
    const double* hres_tlbr = reinterpret_cast<const double*>(t->h_resB + (size_t)k * L.strideB + L.o_tlbr_bytes);
    int n_pairs = (s.n_rows > 1) ? hresB[L.o_hdr] : 0;   // live-slot duplicate candidates found by the device
    const int32_t* live_pairs = hresB + L.o_pairs;
    bool pairs_overflow = false;
    if (n_pairs > kPairCap) { pairs_overflow = true; n_pairs = 0; }
    else if (n_pairs > kPairPrefetch) {
      BT_CUDA(cudaMemcpyAsync(t->h_pairs, dst.pairs + (size_t)k * 2 * kPairCap, sizeof(int32_t) * 2 * (size_t)n_pairs,
                              cudaMemcpyDeviceToHost, st));
      BT_CUDA(cudaStreamSynchronize(st));
      live_pairs = t->h_pairs;
    }
    int n_bpairs = 0;
    if (n_births > 0) {
      const bool reid = s.cfg.with_reid != 0;
      const bool birth_dup = nl > 0;       // a newborn can duplicate a lost track: test births x lost
      int32_t* hB = t->h_birth + (size_t)sid * (2 * t->md + t->cap);
      int32_t* dB = t->d_birth + (size_t)sid * (2 * t->md + t->cap);
      memcpy(hB, s.v_birth_slot.data(), sizeof(int32_t) * n_births);
      memcpy(hB + n_births, s.v_birth_det.data(), sizeof(int32_t) * n_births);
      // global slot indices for the births x lost duplicate test
      int32_t* hBa = hB + 2 * n_births;
      if (birth_dup) {
        // reuse the tail: [birth global slots n_births][lost global slots nl] cannot exceed md + cap entries... births <= md
        for (int i = 0; i < nl; ++i) hBa[i] = sid * t->cap + new_lost[i];
      }
      const size_t bytesB = sizeof(int32_t) * (2 * (size_t)n_births + (birth_dup ? nl : 0));
      BT_CUDA(cudaMemcpyAsync(dB, hB, bytesB, cudaMemcpyHostToDevice, st));
      BT_TRY(btk_frame_births(ctx, dst, sid, s.parity, dB, dB + n_births, n_births, fc, (reid && s.m > 0) ? 1 : 0));
      if (birth_dup) {
        // births by global slot: a second tiny list built on the host next to the first
        std::vector<int>& tmp = s.v_tmp; tmp.resize(n_births);
        for (int i = 0; i < n_births; ++i) tmp[i] = sid * t->cap + s.v_birth_slot[i];
        BT_CUDA(cudaMemcpyAsync(t->d_list, tmp.data(), sizeof(int32_t) * n_births, cudaMemcpyHostToDevice, st));
        BT_CUDA(cudaMemsetAsync(t->d_bpair_count, 0, sizeof(int32_t), st));
        BT_TRY(btk_iou_pairs_below(ctx, dst.tlbr, t->d_list, n_births, dB + 2 * n_births, nl, c.duplicate_iou_dist,
                                   t->d_bpairs, t->d_bpair_count, t->bpair_cap));
        BT_CUDA(cudaMemcpyAsync(t->h_bpair_count, t->d_bpair_count, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        BT_CUDA(cudaStreamSynchronize(st));  // rare second sync: births next to lost tracks
        n_bpairs = *t->h_bpair_count;
        if (n_bpairs > t->bpair_cap) { pairs_overflow = true; n_bpairs = 0; }
        else if (n_bpairs > 0) {
          BT_CUDA(cudaMemcpyAsync(t->h_bpairs, t->d_bpairs, sizeof(int32_t) * 2 * n_bpairs, cudaMemcpyDeviceToHost, st));
          BT_CUDA(cudaStreamSynchronize(st));
        }
      }
    }
    // box of a track of the new tracked list (born-now tracks: the detection's own box, demo:624-648 on initiate's mean)
    auto box_of = [&](int slot, double* out4) {
      const bool born_now = meta[slot].f32_state && meta[slot].start_frame == frame_id;
      if (!born_now) memcpy(out4, hres_tlbr + 4 * (size_t)slot, 4 * sizeof(double));
      else {
        const int32_t* bx = s.host_boxes + 4 * (size_t)meta[slot].det_index;
        for (int q = 0; q < 4; ++q) out4[q] = (double)bx[q];
      }
    };
    // ---- remove_duplicate_stracks (demo:1637, demo:1665-1680) ----
    std::vector<uint8_t>& dupa = s.scratch_a; dupa.assign(nt, 0);
    std::vector<uint8_t>& dupb = s.scratch_b; dupb.assign(nl, 0);
    auto resolve = [&](int p, int q) {   // p: position in tracked, q: position in lost (demo:1669-1677)
      const int timep = meta[new_tracked[p]].frame_id - meta[new_tracked[p]].start_frame;
      const int timeq = meta[new_lost[q]].frame_id - meta[new_lost[q]].start_frame;
      if (timep > timeq) dupb[q] = 1;
      else dupa[p] = 1;
    };
    if (pairs_overflow) {
      // more candidates than the device keeps (thousands of coincident boxes): exact test on the host
      for (int p = 0; p < nt; ++p) {
        double a[4]; box_of(new_tracked[p], a);
        for (int q = 0; q < nl; ++q) {
          const double* b = hres_tlbr + 4 * (size_t)new_lost[q];
          const double ix1 = std::max(a[0], b[0]), iy1 = std::max(a[1], b[1]), ix2 = std::min(a[2], b[2]), iy2 = std::min(a[3], b[3]);
          double iou = 0.0;
          if (!(ix2 <= ix1 || iy2 <= iy1)) {
            const double inter = (ix2 - ix1) * (iy2 - iy1);
            iou = inter / ((a[2] - a[0]) * (a[3] - a[1]) + (b[2] - b[0]) * (b[3] - b[1]) - inter);
          }
          if (1.0 - iou < c.duplicate_iou_dist) resolve(p, q);
        }
      }
    } else if (n_pairs > 0 || n_bpairs > 0) {
      std::vector<int>& pos_t = s.scratch_pos_t; pos_t.assign(s.high_water, -1);
      std::vector<int>& pos_l = s.scratch_pos_l; pos_l.assign(s.high_water, -1);
      for (int i = 0; i < nt; ++i) pos_t[new_tracked[i]] = i;
      for (int i = 0; i < nl; ++i) pos_l[new_lost[i]] = i;
      for (int q = 0; q < n_pairs; ++q) {
        const int i = live_pairs[2 * q], j = live_pairs[2 * q + 1];
        if (i >= s.high_water || j >= s.high_water) continue;
        if (pos_t[i] >= 0 && pos_l[j] >= 0) resolve(pos_t[i], pos_l[j]);
        if (pos_t[j] >= 0 && pos_l[i] >= 0) resolve(pos_t[j], pos_l[i]);
      }
      for (int q = 0; q < n_bpairs; ++q)     // (birth index, lost position)
        resolve(pos_t[s.v_birth_slot[t->h_bpairs[2 * q]]], t->h_bpairs[2 * q + 1]);
    }
    if (pairs_overflow || n_pairs > 0 || n_bpairs > 0) s.pre_valid = false;   // the lists may change below
    s.tracked.clear();
    s.lost.clear();
    for (int i = 0; i < nt; ++i)
      if (!dupa[i]) s.tracked.push_back(new_tracked[i]);
    for (int i = 0; i < nl; ++i)
      if (!dupb[i]) s.lost.push_back(new_lost[i]);
    // the boxes of the returned list are on the host already (result part B / this frame's detections): bt_get_tracks
    // assembles them on demand from there (valid until this stream's next step)
    s.tlbr_src = hres_tlbr;
    s.tlbr_n_rows = s.n_rows;
    s.tlbr_region = k;
    s.tlbr_gen = t->region_gen[k];
    s.boxes_valid = true;
    s.tlbr_cache_valid = true;
    // ---- recycle slots that left both lists ----
    for (int slot : s.tracked) meta[slot].mark = 1;
    for (int slot : s.lost) meta[slot].mark = 1;
    bool freed = false;
    for (int slot = 0; slot < s.high_water; ++slot) {
      if (meta[slot].used && !meta[slot].mark) {
        meta[slot] = SlotMeta();
        s.free_slots.push_back(slot);
        freed = true;
      }
      meta[slot].mark = 0;
    }
    if (freed) std::sort(s.free_slots.begin(), s.free_slots.end(), std::greater<int>());
    if (infos) {
      bt_frame_info& info = infos[k];
      memset(&info, 0, sizeof(info));
      info.frame_id = frame_id;
      info.n_tracked = (int)s.tracked.size();
      info.n_lost = (int)s.lost.size();
      info.n_removed_total = s.n_removed_total;
      info.n_pool = s.n_pool;
      info.n_high = s.n_high;
      info.n_low = s.n_low;
      info.n_unconfirmed = s.n_unc;
      info.n_matches1 = s.n_m1;
      info.n_matches2 = s.n_m2;
      info.n_matches3 = s.n_m3;
      info.n_births = n_births;
      info.n_births_skipped = s.n_births_skipped;
    }
  }
  {
    bool any_births = false;
    for (int k = 0; k < count; ++k) any_births = any_births || t->streams[sids[k]].n_births > 0;
    if (any_births) { BT_CUDA(cudaEventRecord(t->ev_tail, st)); t->tail_pending = true; }
  }
  HOST_MARK(BT_SEG_HOST_FINAL);
  if (t->host_debug) {
    fprintf(stderr, "  enqueue detail (us):");
    for (int i = 1; i < fm_n; ++i) fprintf(stderr, " %s %.1f", fm_name[i], 1e3 * (fm_t[i] - fm_t[i - 1]));
    fprintf(stderr, "\n");
  }
  if (t->host_debug)
    fprintf(stderr, "step of %d stream(s), host phases (us): enqueue %.1f wait_x %.1f lists %.1f wait_end %.1f final %.1f\n", count,
            1e3 * hphase[0], 1e3 * hphase[1], 1e3 * hphase[2], 1e3 * hphase[3], 1e3 * hphase[4]);
  return BT_OK;
}

extern "C" {

int32_t bt_tracker_reset_stream(bt_ctx* ctx, int32_t stream_id, const bt_config* cfg) {
  if (!ctx) return BT_ERR_INVALID;
  BT_CUDA(cudaSetDevice(ctx->device));
  bt_tracker* t = ctx->trk;
  BT_CHECK(stream_id < t->S, BT_ERR_INVALID, "stream id %d out of range (ctx has %d)", stream_id, t->S);
  BT_CUDA(cudaStreamSynchronize(ctx->stream));
  BT_CUDA(cudaStreamSynchronize(ctx->copy_stream));
  const bt_cand& cand = *bt_lap_own_cand(ctx);
  const size_t D = t->D;
  // captured frame graphs carry the thresholds of the configuration they were captured under
  for (auto& g : t->graphs) {
    if (g.exec) cudaGraphExecDestroy(g.exec);
    if (g.graph) cudaGraphDestroy(g.graph);
  }
  t->graphs.clear();
  for (int sid = (stream_id < 0 ? 0 : stream_id); sid < (stream_id < 0 ? t->S : stream_id + 1); ++sid) {
    reset_stream(t, t->streams[sid], cfg);
    BT_CUDA(cudaMemsetAsync(t->st.feat16 + (size_t)sid * t->cap * D, 0, sizeof(__half) * (size_t)t->cap * D, ctx->stream));
    BT_CUDA(cudaMemsetAsync(t->st.norm + (size_t)sid * t->cap, 0, sizeof(float) * t->cap, ctx->stream));
    BT_CUDA(cudaMemsetAsync(reinterpret_cast<char*>(cand.cnt) + (size_t)sid * cand.s_cnt, 0, cand.clear_bytes, ctx->stream));
  }
  return BT_OK;
}

int32_t bt_tracker_reset(bt_ctx* ctx, const bt_config* cfg) { return bt_tracker_reset_stream(ctx, -1, cfg); }

int32_t bt_submit_streams(bt_ctx* ctx, int32_t count, const int32_t* stream_ids, const int32_t* const* boxes,
                          const float* const* scores, const void* const* feats, const int32_t* m, int32_t feat_dtype,
                          const float* const* face_sims, int32_t loc) {
  if (!ctx) return BT_ERR_INVALID;
  BT_CUDA(cudaSetDevice(ctx->device));
  return submit_batch(ctx, count, stream_ids, boxes, scores, feats, m, feat_dtype, face_sims, loc);
}

int32_t bt_step_streams(bt_ctx* ctx, int32_t count, const int32_t* stream_ids, bt_frame_info* infos) {
  if (!ctx) return BT_ERR_INVALID;
  BT_CUDA(cudaSetDevice(ctx->device));
  return step_batch(ctx, count, stream_ids, infos);
}

int32_t bt_update_streams(bt_ctx* ctx, int32_t count, const int32_t* stream_ids, const int32_t* const* boxes,
                          const float* const* scores, const void* const* feats, const int32_t* m, int32_t feat_dtype,
                          const float* const* face_sims, int32_t loc, bt_frame_info* infos) {
  if (!ctx) return BT_ERR_INVALID;
  BT_CUDA(cudaSetDevice(ctx->device));
  BT_TRY(submit_batch(ctx, count, stream_ids, boxes, scores, feats, m, feat_dtype, face_sims, loc));
  return step_batch(ctx, count, stream_ids, infos);
}

int32_t bt_update_arrays(bt_ctx* ctx, const int32_t* boxes, const float* scores, const float* feats, int32_t m,
                         int32_t loc, bt_frame_info* info) {
  const int32_t sid = 0;
  const void* f = feats;
  return bt_update_streams(ctx, 1, &sid, &boxes, &scores, &f, &m, BT_F32, nullptr, loc, info);
}

int32_t bt_input_buffers(bt_ctx* ctx, int32_t stream_id, int32_t** boxes, float** scores, void** feats16) {
  if (!ctx) return BT_ERR_INVALID;
  bt_tracker* t = ctx->trk;
  BT_CHECK(stream_id >= 0 && stream_id < t->S, BT_ERR_INVALID, "stream id %d out of range (ctx has %d)", stream_id, t->S);
  const StreamState& s = t->streams[stream_id];
  const size_t g0 = (size_t)s.next_parity * t->S * t->md + (size_t)stream_id * t->md;
  if (boxes) *boxes = t->st.det_boxes + g0 * 4;
  if (scores) *scores = t->st.det_scores + g0;
  if (feats16) *feats16 = t->st.det16 + g0 * t->D;
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
  auto clear_lists = [&]() -> int32_t {
    BT_CUDA(cudaMemsetAsync(cand.cnt, 0, cand.s_cnt * (size_t)t->S, st));
    return BT_OK;
  };
  double sum = 0.0;
  if (idempotent) {
    // the frame description of the last step is still in the device control block
    const bt_assoc_frame* dAF = &reinterpret_cast<const FrameDesc*>(t->st.ctrl)->AF;
    BT_TRY(btk_assoc_launch(ctx, t->last_assoc, 0, t->last_AF, dAF, 0, 0, 0));   // warm
    BT_CUDA(cudaEventRecord(e0, st));
    for (int i = 0; i < iters; ++i) BT_TRY(btk_assoc_launch(ctx, t->last_assoc, 0, t->last_AF, dAF, 0, 0, 0));
    BT_CUDA(cudaEventRecord(e1, st));
    BT_CUDA(cudaStreamSynchronize(st));
    float ms = 0.f;
    BT_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    sum = ms;
  } else {
    for (int i = 0; i < iters; ++i) {
      BT_TRY(clear_lists());
      BT_CUDA(cudaEventRecord(e0, st));
      BT_TRY(btk_assoc(ctx, t->last_assoc, t->last_assoc_precision));
      BT_CUDA(cudaEventRecord(e1, st));
      BT_CUDA(cudaStreamSynchronize(st));
      float ms = 0.f;
      BT_CUDA(cudaEventElapsedTime(&ms, e0, e1));
      sum += ms;
    }
  }
  BT_TRY(clear_lists());
  BT_CUDA(cudaStreamSynchronize(st));
  if (getenv("BT_ASSOC_DEBUG") && (atoi(getenv("BT_ASSOC_DEBUG")) & 32768)) btk_assoc_stamps_report();
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

int32_t bt_get_tracks_stream(bt_ctx* ctx, int32_t stream_id, int32_t which, int32_t cap, int32_t* n, int32_t* ids,
                             int32_t* state, int32_t* activated, int32_t* frame_id, int32_t* start_frame,
                             int32_t* tracklet_len, int32_t* det_index, float* score, double* tlbr, double* mean,
                             double* cov) {
  if (!ctx) return BT_ERR_INVALID;
  BT_CUDA(cudaSetDevice(ctx->device));
  bt_tracker* t = ctx->trk;
  BT_CHECK(stream_id >= 0 && stream_id < t->S, BT_ERR_INVALID, "stream id %d out of range (ctx has %d)", stream_id, t->S);
  BT_CHECK(which == 0 || which == 1, BT_ERR_INVALID, "which must be 0 (tracked) or 1 (lost)");
  StreamState& s = t->streams[stream_id];
  const std::vector<int>& lst = which == 0 ? s.tracked : s.lost;
  const int cnt = (int)lst.size();
  if (n) *n = cnt;
  BT_CHECK(cnt <= cap || !(ids || state || activated || frame_id || start_frame || tracklet_len || det_index ||
                           score || tlbr || mean || cov),
           BT_ERR_CAPACITY, "list has %d tracks, buffers hold %d", cnt, cap);
  for (int i = 0; i < cnt; ++i) {
    const SlotMeta& tm = s.meta[lst[i]];
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
  if (tlbr && which == 0 && s.tlbr_cache_valid && s.boxes_valid && t->region_gen[s.tlbr_region] == s.tlbr_gen) {
    for (int i = 0; i < cnt; ++i) {
      const int slot = lst[i];
      const SlotMeta& tm = s.meta[slot];
      const bool born_now = tm.f32_state && tm.start_frame == s.frame_id;   // its box is the detection's (demo:624-648 on initiate's mean)
      if (!born_now && slot < s.tlbr_n_rows) memcpy(tlbr + 4 * (size_t)i, s.tlbr_src + 4 * (size_t)slot, 4 * sizeof(double));
      else {
        const int32_t* bx = s.host_boxes + 4 * (size_t)tm.det_index;
        for (int q = 0; q < 4; ++q) tlbr[4 * (size_t)i + q] = (double)bx[q];
      }
    }
    tlbr = nullptr;
  }
  if (tlbr || mean || cov) {
    cudaStream_t st = ctx->stream;
    for (int i = 0; i < cnt; ++i) t->h_list[i] = stream_id * t->cap + lst[i];
    BT_CUDA(cudaMemcpyAsync(t->d_list, t->h_list, sizeof(int32_t) * cnt, cudaMemcpyHostToDevice, st));
    struct Job { const double* src; double* dst; int width; };
    const Job jobs[3] = {{t->st.tlbr, tlbr, 4}, {t->st.mean, mean, 8}, {t->st.cov, cov, 64}};
    for (const Job& j : jobs) {
      if (!j.dst) continue;
      BT_TRY(btk_gather_rows_f64(ctx, j.src, t->d_list, cnt, j.width, t->d_gather));
      BT_CUDA(cudaMemcpyAsync(j.dst, t->d_gather, sizeof(double) * (size_t)cnt * j.width, cudaMemcpyDeviceToHost, st));
      BT_CUDA(cudaStreamSynchronize(st));
    }
  }
  return BT_OK;
}

int32_t bt_get_tracks(bt_ctx* ctx, int32_t which, int32_t cap, int32_t* n, int32_t* ids, int32_t* state,
                      int32_t* activated, int32_t* frame_id, int32_t* start_frame, int32_t* tracklet_len,
                      int32_t* det_index, float* score, double* tlbr, double* mean, double* cov) {
  return bt_get_tracks_stream(ctx, 0, which, cap, n, ids, state, activated, frame_id, start_frame, tracklet_len,
                              det_index, score, tlbr, mean, cov);
}

int32_t bt_get_track_features_stream(bt_ctx* ctx, int32_t stream_id, int32_t which, int32_t cap, float* curr,
                                     float* smooth) {
  if (!ctx) return BT_ERR_INVALID;
  BT_CUDA(cudaSetDevice(ctx->device));
  bt_tracker* t = ctx->trk;
  BT_CHECK(stream_id >= 0 && stream_id < t->S, BT_ERR_INVALID, "stream id %d out of range (ctx has %d)", stream_id, t->S);
  BT_CHECK(which == 0 || which == 1, BT_ERR_INVALID, "which must be 0 (tracked) or 1 (lost)");
  BT_CHECK(!smooth || t->st.smooth32 != nullptr, BT_ERR_STATE, "ctx was created with BT_FLAG_NO_F32_FEATURES");
  StreamState& s = t->streams[stream_id];
  const std::vector<int>& lst = which == 0 ? s.tracked : s.lost;
  const int cnt = (int)lst.size();
  BT_CHECK(cnt <= cap, BT_ERR_CAPACITY, "list has %d tracks, buffers hold %d", cnt, cap);
  if (cnt == 0) return BT_OK;
  cudaStream_t st = ctx->stream;
  if (!t->d_gather32) BT_CUDA(cudaMalloc(&t->d_gather32, sizeof(float) * (size_t)t->cap * t->D));
  for (int i = 0; i < cnt; ++i) t->h_list[i] = stream_id * t->cap + lst[i];
  BT_CUDA(cudaMemcpyAsync(t->d_list, t->h_list, sizeof(int32_t) * cnt, cudaMemcpyHostToDevice, st));
  if (curr) {
    // the current feature is the adopted detection row divided by its norm (demo:497-502): stored like that
    // with fp32 ingest, derived exactly from the raw fp16 row + norm with fp16 ingest
    if (s.feat_dtype == BT_F32 && t->st.curr32) BT_TRY(btk_gather_rows_f32(ctx, t->st.curr32, t->d_list, cnt, t->D, t->d_gather32));
    else BT_TRY(btk_gather_curr_f16(ctx, t->st.feat16, t->st.norm, t->d_list, cnt, t->D, t->d_gather32));
    BT_CUDA(cudaMemcpyAsync(curr, t->d_gather32, sizeof(float) * (size_t)cnt * t->D, cudaMemcpyDeviceToHost, st));
    BT_CUDA(cudaStreamSynchronize(st));
  }
  if (smooth) {
    BT_TRY(btk_gather_rows_f32(ctx, t->st.smooth32, t->d_list, cnt, t->D, t->d_gather32));
    BT_CUDA(cudaMemcpyAsync(smooth, t->d_gather32, sizeof(float) * (size_t)cnt * t->D, cudaMemcpyDeviceToHost, st));
    BT_CUDA(cudaStreamSynchronize(st));
  }
  return BT_OK;
}

int32_t bt_get_track_features(bt_ctx* ctx, int32_t which, int32_t cap, float* curr, float* smooth) {
  return bt_get_track_features_stream(ctx, 0, which, cap, curr, smooth);
}

// The matches of the last step in the reference's index spaces (demo:1556, demo:1571, demo:1604), assembled on
// demand from what the step left on the host (assignment vectors, pool order, scores) -- the hot path does not pay
// for a diagnostic read-back.
static int32_t build_matches(bt_ctx* ctx, bt_tracker* t, StreamState& s) {
  BT_CHECK(s.matches_valid && s.boxes_valid && t->region_gen[s.tlbr_region] == s.tlbr_gen, BT_ERR_STATE,
           "the matches of the last step are no longer available (another step or submit has reused its buffers)");
  for (auto& mm : s.matches) mm.clear();
  const int m = s.m;
  const float* sc = s.sc;
  const bt_config& c = s.cfg;
  std::vector<int>& hi_pos = s.v_hi_pos; hi_pos.assign(m, -1);
  std::vector<int>& lo_pos = s.v_lo_pos; lo_pos.assign(m, -1);
  int nh = 0, nl = 0;
  for (int j = 0; j < m; ++j) {
    const double sd = (double)sc[j];
    if (sd > c.track_high_thresh) hi_pos[j] = nh++;
    else if (sd >= c.track_low_thresh) lo_pos[j] = nl++;
  }
  std::vector<uint8_t>& taken = s.scratch_a; taken.assign(m, 0);
  if (s.n_rows > 0) {
    for (int i = 0; i < s.n_pool; ++i) {
      const int j = s.hx[0][s.v_pool[i]];
      if (j >= 0) { s.matches[0].push_back(i); s.matches[0].push_back(hi_pos[j]); taken[j] = 1; }
    }
    for (int i = 0; i < (int)s.v_r_tracked.size(); ++i) {
      const int j = s.hx[1][s.v_r_tracked[i]];
      if (j >= 0) { s.matches[1].push_back(i); s.matches[1].push_back(lo_pos[j]); taken[j] = 1; }
    }
    std::vector<int>& u_det_pos = s.v_u_det_pos; u_det_pos.assign(m, -1);
    int q = 0;
    for (int j = 0; j < m; ++j)
      if (hi_pos[j] >= 0 && !taken[j]) u_det_pos[j] = q++;
    for (int i = 0; i < s.n_unc; ++i) {
      const int j = s.hx[2][s.v_unconfirmed[i]];
      if (j >= 0) { s.matches[2].push_back(i); s.matches[2].push_back(u_det_pos[j]); }
    }
  }
  return BT_OK;
}

int32_t bt_get_matches_stream(bt_ctx* ctx, int32_t stream_id, int32_t stage, int32_t cap, int32_t* n, int32_t* pairs) {
  if (!ctx) return BT_ERR_INVALID;
  bt_tracker* t = ctx->trk;
  BT_CHECK(stream_id >= 0 && stream_id < t->S, BT_ERR_INVALID, "stream id %d out of range (ctx has %d)", stream_id, t->S);
  BT_CHECK(stage >= 1 && stage <= 3, BT_ERR_INVALID, "stage must be 1, 2 or 3");
  StreamState& s = t->streams[stream_id];
  if (s.frame_id > 0) BT_TRY(build_matches(ctx, t, s));
  const std::vector<int32_t>& mm = s.matches[stage - 1];
  const int cnt = (int)mm.size() / 2;
  if (n) *n = cnt;
  if (pairs) {
    BT_CHECK(cnt <= cap, BT_ERR_CAPACITY, "%d matches, buffer holds %d", cnt, cap);
    if (cnt) memcpy(pairs, mm.data(), sizeof(int32_t) * 2 * cnt);
  }
  return BT_OK;
}

int32_t bt_get_matches(bt_ctx* ctx, int32_t stage, int32_t cap, int32_t* n, int32_t* pairs) {
  return bt_get_matches_stream(ctx, 0, stage, cap, n, pairs);
}

}  // extern "C"
