// Internal declarations shared by the translation units of libbotsort_b200.so.
// Not part of the ABI (that is include/botsort_b200.h).
#pragma once

#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <string.h>
#include <vector>

#include "../../include/botsort_b200.h"

struct bt_tracker;  // track_step.cu

constexpr size_t kBtCropLutOffset = 16384;
struct bt_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  int max_tracks = 0, max_dets = 0, feat_dim = 0;
  int n_streams = 1;             // video streams (trackers) of this ctx: a leading batch dimension of every kernel
  cudaStream_t copy_stream = nullptr;   // input staging of bt_submit_streams (overlaps the frame step)
  cudaStream_t side_stream = nullptr;   // work of a frame that is independent of its main chain (feature EMA)
  uint32_t flags = 0;
  int num_sms = 148;
  int pdl = 1;   // programmatic dependent launch between the frame step's kernels (BT_NO_PDL=1 turns it off)
  std::string err;
  int64_t launches = 0;
  char* d_desc = nullptr;   // device scratch for the launch descriptors of the stand-alone entry points (16 KB), then the
                            // crop normalisation table float[3][256] at kBtCropLutOffset (written once at create)
  // bump arena for the stand-alone entry points (device) and a pinned mirror for small results
  char* arena = nullptr;
  size_t arena_cap = 0, arena_off = 0;
  char* pinned = nullptr;
  size_t pinned_cap = 0;
  // LAP workspaces (lap.cu), sized for max_tracks x max_dets
  struct bt_lap_ws* lap = nullptr;
  // ReID similarity GEMM workspaces (reid_gemm.cu)
  struct bt_gemm_ws* gemm = nullptr;
  bt_tracker* trk = nullptr;
};

int32_t bt_fail(bt_ctx* ctx, int32_t code, const char* fmt, ...);

#ifdef __CUDACC__
// Launch on `stream`, optionally as a programmatic dependent of the stream's previous kernel: the grid
// may then be scheduled before that kernel has completed, and must execute griddepcontrol.wait
// (bt_grid_dependency_wait) before it touches anything the earlier kernels of the stream produce -- and at
// the latest before it exits, so that completion stays ordered for its own dependents.
template <typename... KArgs, typename... Args>
static inline cudaError_t bt_launch_on(bt_ctx* ctx, cudaStream_t stream, bool dependent, void (*kern)(KArgs...), dim3 grid,
                                       dim3 block, size_t smem, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = (dependent && ctx->pdl) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
template <typename... KArgs, typename... Args>
static inline cudaError_t bt_launch(bt_ctx* ctx, bool dependent, void (*kern)(KArgs...), dim3 grid, dim3 block,
                                    size_t smem, Args... args) {
  return bt_launch_on(ctx, ctx->stream, dependent, kern, grid, block, smem, args...);
}
__device__ __forceinline__ void bt_grid_dependency_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void bt_grid_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// Box -> two 32-bit words of 15-bit integer corners rounded OUTWARD, the operands of the association
// kernel's two-subtraction overlap screen:
//   .x = (x1 + 1) | (y1 + 1) << 16        .y = x2 | y2 << 16 | 0x80008000
// Saturation keeps the screen conservative: lowering a lower corner or raising an upper corner only adds
// overlaps; upper corners above 32767 meet lower corners saturated to 32766 (test passes); a ROW's negative
// corners clamp to 0 because detection corners are >= 0 -- a detection with a negative corner (outside the
// reference's domain: YOLOX._postprocess clamps at 0, demo:1009) and NaNs fall back to the whole range, so
// the packed test never rejects a pair whose exact IoU is positive.
__device__ __forceinline__ uint2 bt_pack16_corners(int lx, int ly, int hx, int hy, bool whole) {
  const uint32_t ix1 = whole ? 0u : (uint32_t)min(max(lx, 0), 32766), iy1 = whole ? 0u : (uint32_t)min(max(ly, 0), 32766);
  const uint32_t ix2 = whole ? 32767u : (uint32_t)min(max(hx, 0), 32767), iy2 = whole ? 32767u : (uint32_t)min(max(hy, 0), 32767);
  return make_uint2((ix1 + 1u) | ((iy1 + 1u) << 16), ix2 | (iy2 << 16) | 0x80008000u);
}
// from an fp32 interval (lower corners rounded down, upper corners rounded up): integer conversions and
// integer clamps only -- fp64 rounding / compares crawl next to a running tcgen05 main loop
__device__ __forceinline__ uint2 bt_pack16_f32(float x1, float y1, float x2, float y2, bool is_col) {
  const bool nan = (x1 != x1) || (y1 != y1) || (x2 != x2) || (y2 != y2);
  const bool whole = nan || (is_col && (x1 < 0.0f || y1 < 0.0f));
  return bt_pack16_corners(__float2int_rd(x1), __float2int_rd(y1), __float2int_ru(x2), __float2int_ru(y2), whole);
}
#endif
extern thread_local std::string g_bt_create_error;

#define BT_CUDA(expr)                                                                              \
  do {                                                                                             \
    cudaError_t e__ = (expr);                                                                      \
    if (e__ != cudaSuccess)                                                                        \
      return bt_fail(ctx, BT_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #expr,               \
                     cudaGetErrorString(e__));                                                     \
  } while (0)

#define BT_CHECK(cond, code, ...)                                                                  \
  do {                                                                                             \
    if (!(cond)) return bt_fail(ctx, code, __VA_ARGS__);                                           \
  } while (0)

#define BT_TRY(expr)                                                                               \
  do {                                                                                             \
    int32_t s__ = (expr);                                                                          \
    if (s__ != BT_OK) return s__;                                                                  \
  } while (0)

#define BT_LAUNCHED(ctx)                                                                           \
  do {                                                                                             \
    (ctx)->launches++;                                                                             \
    BT_CUDA(cudaGetLastError());                                                                   \
  } while (0)

// ---- arena ----------------------------------------------------------------------------------
int32_t bt_arena_reset(bt_ctx* ctx);
int32_t bt_arena_reserve(bt_ctx* ctx, size_t total);
int32_t bt_arena_alloc(bt_ctx* ctx, size_t bytes, void** out);
template <typename T>
static inline int32_t bt_arena(bt_ctx* ctx, size_t count, T** out) {
  void* p = nullptr;
  int32_t s = bt_arena_alloc(ctx, count * sizeof(T), &p);
  *out = reinterpret_cast<T*>(p);
  return s;
}
// stage an input: loc==BT_DEVICE returns the pointer itself, BT_HOST copies into the arena
int32_t bt_stage_in(bt_ctx* ctx, const void* src, size_t bytes, int32_t loc, const void** dev);
// stage an output: loc==BT_DEVICE returns the pointer itself, BT_HOST an arena buffer
int32_t bt_stage_out(bt_ctx* ctx, void* dst, size_t bytes, int32_t loc, void** dev);
// finish an output: BT_HOST copies back (async) -- caller must bt_finish() afterwards
int32_t bt_unstage_out(bt_ctx* ctx, void* dst, const void* dev, size_t bytes, int32_t loc);
// BT_HOST: synchronise the stream so host buffers are valid; BT_DEVICE: nothing
int32_t bt_finish(bt_ctx* ctx, int32_t loc);

template <typename T>
static inline int32_t bt_in(bt_ctx* ctx, const T* src, size_t count, int32_t loc, const T** dev) {
  const void* p = nullptr;
  int32_t s = bt_stage_in(ctx, src, count * sizeof(T), loc, &p);
  *dev = reinterpret_cast<const T*>(p);
  return s;
}
template <typename T>
static inline int32_t bt_out(bt_ctx* ctx, T* dst, size_t count, int32_t loc, T** dev) {
  void* p = nullptr;
  int32_t s = bt_stage_out(ctx, dst, count * sizeof(T), loc, &p);
  *dev = reinterpret_cast<T*>(p);
  return s;
}

// ---- constants of the reference's Kalman filter (demo:163-164) -------------------------------
#define BT_STD_POS (1.0 / 20)
#define BT_STD_VEL (1.0 / 160)

// ---- reference cost fusion (shared by the association epilogue and the LAP's exact re-costing) ----
#ifdef __CUDACC__
// bbox_iou, demo:1695-1713, as a distance
__device__ __forceinline__ double bt_iou_dist_f64(const double* __restrict__ a, const double* __restrict__ b) {
  const double ixmin = fmax(a[0], b[0]), iymin = fmax(a[1], b[1]);
  const double ixmax = fmin(a[2], b[2]), iymax = fmin(a[3], b[3]);
  if (ixmax <= ixmin || iymax <= iymin) return 1.0;
  const double inter = (ixmax - ixmin) * (iymax - iymin);
  const double area1 = (a[2] - a[0]) * (a[3] - a[1]);
  const double area2 = (b[2] - b[0]) * (b[3] - b[1]);
  return 1.0 - inter / (area1 + area2 - inter);
}
// first association, demo:1539-1554
__device__ __forceinline__ double bt_fuse_stage1(double iou_d, float sim, float face, float appearance) {
  float emb = 1.0f - sim;
  const float face_emb = 1.0f - face;
  if (fminf(emb, face_emb) > appearance) emb = 1.0f;
  return fmin(iou_d, (double)emb);
}
// unconfirmed tracks, demo:1599-1602
__device__ __forceinline__ double bt_fuse_stage3(double iou_d, float sim, float appearance, double proximity) {
  float emb = 1.0f - fmaxf(0.0f, sim);
  if (emb > appearance) emb = 1.0f;
  if (iou_d > proximity) emb = 1.0f;
  return fmin(iou_d, (double)emb);
}
#endif

// ---- batch of video streams served by one launch ------------------------------------------------
// Per-slot arrays of the track store are indexed by the global slot gs = stream * max_tracks + slot,
// per-detection arrays by gd = stream * max_dets + j; the frame's control segment (pool list, states,
// row kinds) of batch entry k starts ctrl_off[k] bytes into the packed control block and its results
// land in result region k.
struct bt_batch {
  int32_t count;
  int32_t sid[BT_MAX_BATCH];       // video stream of batch entry k
  int32_t m[BT_MAX_BATCH];         // detections this frame
  int32_t n_rows[BT_MAX_BATCH];    // slots in use (high-water mark) = rows of the association problem
  int32_t n_pool[BT_MAX_BATCH];    // pool tracks (Kalman predict)
  int32_t ctrl_off[BT_MAX_BATCH];  // byte offset of the control segment: [pool_idx n_pool][pool_state n_pool][row_kind n_rows (pad 4)][pool_pos n_rows]
  uint8_t noise_f32[BT_MAX_BATCH]; // every pooled mean is still float32 (NumPy >= 2 quirk, frame 2)
  uint8_t parity[BT_MAX_BATCH];    // which half of the double-buffered inputs holds this frame
  uint8_t want_norm[BT_MAX_BATCH]; // detection feature norms are needed (unconfirmed rows exist)
};
static inline int bt_batch_max(const int32_t* v, int count) {
  int mx = 0;
  for (int k = 0; k < count; ++k) mx = v[k] > mx ? v[k] : mx;
  return mx;
}

// Device-side track store + per-frame buffers of a ctx (track_step.cu owns the allocations).
struct bt_store {
  int32_t S, cap, md, D;
  // per slot (gs)
  double *mean, *cov, *tlbr;
  float* tlbr_f32;
  __half* feat16;      // raw fp16 feature of the track's latest detection = A operand of the similarity GEMM
  float* norm;         // its L2 norm: body_curr_feature = feat16 / norm (demo:497-502)
  float* curr32;       // fp32 ingest only: the normalised fp32 current feature (exact re-costing + exposed state)
  float* smooth32;     // EMA feature (A10), optional
  uint8_t* slot_f32;
  // per detection (gd), per input parity where noted
  int32_t* det_boxes;  // [2][S*md][4]
  float* det_scores;   // [2][S*md]
  __half* det16;       // [2][S*md][D] raw fp16 features = B operand
  float* det32;        // [2][S*md][D] fp32 ingest staging (allocated on first use)
  float* det_norm;     // [S*md] L2 norms of the detection rows (when computed)
  double *det_tlbr, *det_xywh;
  float* det_xywh32;
  uint8_t* col_kind;
  uint2* col_pk;
  // per batch entry
  char* ctrl;          // packed control block
  char* res;           // result regions part A (read back right after the LAP): x1,x2,x3 [cap each] | scores[md] | boxes[4 md]
  char* res_host;      // the pinned host copy of part A when the kernels publish it themselves (direct results), else null
  char* resB;          // result regions part B (read back at the end): pair count (2 ints) | pairs[2*prefetch] | tlbr[4*cap] f64
  int32_t* y;          // [BT_MAX_BATCH][3][md]
  int32_t* pairs;      // [BT_MAX_BATCH][2*pair_cap] duplicate candidates beyond the prefetch
  int32_t pair_cap;
};

// ---- kernel launchers (device pointers only; enqueue on ctx->stream) --------------------------
// kalman.cu (stand-alone entry points)
int32_t btk_kalman_initiate(bt_ctx* ctx, const float* xywh, const int32_t* src_idx, double* mean,
                            double* cov, double* tlbr, float* tlbr_f32, const int32_t* dst_idx, int32_t k,
                            uint8_t* slot_f32 = nullptr);
int32_t btk_kalman_predict(bt_ctx* ctx, double* mean, double* cov, double* tlbr, float* tlbr_f32,
                           const int32_t* state, const int32_t* idx, int32_t n, int32_t noise_f32,
                           uint8_t* slot_f32 = nullptr);
int32_t btk_kalman_update(bt_ctx* ctx, double* mean, double* cov, double* tlbr, float* tlbr_f32,
                          const double* meas, const int32_t* track_idx, const int32_t* meas_idx,
                          const uint8_t* noise_f32, int32_t k);
int32_t btk_kalman_project(bt_ctx* ctx, const double* mean, const double* cov, double* pmean,
                           double* pcov, int32_t n);
// iou.cu
int32_t btk_iou_distance(bt_ctx* ctx, const double* a, int32_t n, const double* b, int32_t m,
                         double* out);
int32_t btk_fuse_score(bt_ctx* ctx, const double* d, const double* s, int32_t n, int32_t m, double* out);
// emits (i, j) pairs with 1-IoU < limit between lists given by slot index; count via atomic
int32_t btk_iou_pairs_below(bt_ctx* ctx, const double* tlbr, const int32_t* a_idx, int32_t n,
                            const int32_t* b_idx, int32_t m, double limit, int32_t* pairs,
                            int32_t* pair_count, int32_t pair_cap);
// features.cu (stand-alone entry points)
int32_t btk_feature_prep(bt_ctx* ctx, const float* feat, int32_t m, int32_t d, float* out_f32,
                         __half* out_f16, int32_t normalise, int32_t dependent = 0);
int32_t btk_feature_ema(bt_ctx* ctx, float* smooth, float* curr, const float* feat,
                        const int32_t* track_idx, const int32_t* feat_idx, const uint8_t* first,
                        int32_t k, int32_t d, float alpha, float one_minus_alpha);

// result-block layout of one batch entry's regions (part A: int32 offsets; part B: int32 offsets + the byte
// offset of the float64 boxes)
struct bt_res_layout {
  size_t o_x, o_sc, o_bx, o_endA, stride;           // part A, regions `stride` bytes apart
  size_t o_hdr, o_pairs, o_tlbr_bytes, strideB;     // part B, regions `strideB` bytes apart
};
// frame_kernels.cu: the per-frame launches of the tracker, all with a leading video-stream dimension
struct bt_frame_cfg {
  double high, low;            // score classes (demo:1501, demo:1531)
  float alpha, one_minus_alpha;
  double dup_limit;
  int32_t device_inputs;       // scores / boxes are echoed into the result block
  int32_t f16_inputs;          // the frame's features came as fp16 (det16 holds them as they are)
  int32_t keep_smooth;
  int32_t prefetch_pairs;      // duplicate pairs kept in the result block
  int32_t with_reid;           // feature rows are part of the frame
  int32_t l2_prefetch;         // the tensor-core association kernel follows: prefetch its operands into L2 beside the prep work
};
// The per-frame kernels read the batch description (`db`: bt_batch in DEVICE memory, uploaded with the frame's
// control block) themselves, so their kernel arguments never change from frame to frame and the whole frame can
// be replayed as a CUDA graph; `b` is the host copy, used for the launch geometry only.  fixed != 0: geometry
// from the ctx capacities instead (what a captured graph needs; surplus blocks exit at once).
// fp32 ingest: det32 rows -> det16 (raw, round to nearest) + L2 norms
// control block: pinned host -> device by a kernel (no copy-engine command on the frame's critical path)
int32_t btk_ctrl_upload(bt_ctx* ctx, const void* h_src, void* d_dst, size_t bytes);
int32_t btk_frame_cast(bt_ctx* ctx, const bt_store& st, const bt_batch& b, const bt_batch* db, int fixed);
// detection prep (boxes -> tlbr / xywh / score class / packed corners) + batched Kalman predict of every
// stream's pool + (optional) detection feature norms: one launch
int32_t btk_frame_prep(bt_ctx* ctx, const bt_store& st, const bt_batch& b, const bt_batch* db, const bt_frame_cfg& fc, int fixed);
// Kalman update of every slot matched by one of the three stages; feature EMA of the same slots (independent of
// it: launched on `stream`, the ctx's side stream)
int32_t btk_frame_post(bt_ctx* ctx, const bt_store& st, const bt_batch& b, const bt_batch* db, const bt_frame_cfg& fc, int fixed,
                       int dependent);
int32_t btk_frame_ema(bt_ctx* ctx, const bt_store& st, const bt_batch& b, const bt_batch* db, const bt_frame_cfg& fc,
                      cudaStream_t stream, int fixed);
// duplicate candidates among all live slots (superset of tracked x lost) + every slot's box
int32_t btk_frame_dup(bt_ctx* ctx, const bt_store& st, const bt_batch& b, const bt_batch* db, const bt_frame_cfg& fc, int fixed,
                      int with_ema);
// births of one stream: Kalman initiate + feature adoption from the frame's detections
int32_t btk_frame_births(bt_ctx* ctx, const bt_store& st, int32_t sid, int32_t parity, const int32_t* d_slot,
                         const int32_t* d_det, int32_t n_births, const bt_frame_cfg& fc, int32_t with_feat);
bt_res_layout bt_res_layout_for(int cap, int md, int prefetch_pairs);
// gathers of list read-backs
int32_t btk_gather_rows_f64(bt_ctx* ctx, const double* src, const int32_t* idx, int32_t n, int32_t width, double* dst);
int32_t btk_gather_rows_f32(bt_ctx* ctx, const float* src, const int32_t* idx, int32_t n, int32_t width, float* dst);
// curr feature of fp16 stores: dst[i] = float(feat16[idx[i]]) / norm[idx[i]]
int32_t btk_gather_curr_f16(bt_ctx* ctx, const __half* feat16, const float* norm, const int32_t* idx, int32_t n,
                            int32_t d, float* dst);

// ---- association (reid_gemm.cu) ----------------------------------------------------------------
// Row kinds / column kinds of the fused association kernel.
enum { BT_ROW_NONE = 0, BT_ROW_POOL_TRACKED = 1, BT_ROW_POOL_OTHER = 2, BT_ROW_UNCONFIRMED = 3 };
enum { BT_COL_NONE = 0, BT_COL_HIGH = 1, BT_COL_LOW = 2 };
// flag bits in the column field of an emitted candidate edge (and of rowcol):
//   SIM   the cost comes from the (fp16 tensor-core) similarity: the LAP re-costs it exactly when the row
//         competes with others (SURVEY hard part 2)
//   AMBIG the appearance gate is within the tensor-core error of opening / closing: always re-costed
#define BT_EDGE_SIM 0x40000000
#define BT_EDGE_AMBIG 0x20000000
#define BT_EDGE_COLMASK 0x1fffffff

// Candidate lists ("ragged dense" adjacency) of ONE video stream: list s in {0,1,2} = association stage 1,2,3.
// Row r owns entries [r*stride, r*stride + cnt[s*rows_cap + r]) of col/cost.
// A row's region of `stride` entries is cut into segments of `seg` columns: the
// edges of row r found in columns [g*seg, (g+1)*seg) are written to
// [r*stride + g*seg, r*stride + g*seg + cnt[list][r][g]).  The tensor-core epilogue owns one
// (row, segment) pair per thread, so it appends with plain stores and a register counter -- no
// atomics; the LAP set-up gathers the segments of the rows it has to look at.
#define BT_CAND_MAXSEG 64   // segments per row (bits of segmask); cnt row pitch
struct bt_cand {
  int32_t* cnt;    // [3][rows_cap][nseg]
  int32_t* total;  // [4] per list: non-zero when any edge was emitted (lets the LAP skip empty stages)
  unsigned long long* segmask;  // [3][rows_cap] bit g: segment g of the row is non-empty (nseg <= 64)
  int32_t* rowdeg;  // [3][rows_cap] edges emitted for the row      } maintained by the emitters with fire-and-forget
  int32_t* indeg;   // [3][cols_cap] edges emitted into the column  } atomics: the LAP classifies rows without
  int32_t* rowcol;  // [3][rows_cap] a column of the row (THE column when rowdeg == 1), with flag bits
  int32_t cols_cap;
  int32_t* deg;    // [3][rows_cap]  row degree after compaction (large-problem path of the LAP)
  int32_t* col;    // [3][rows_cap*stride]
  double* cost;    // [3][rows_cap*stride]
  int32_t rows_cap;
  int32_t stride;  // entries per row >= max_dets + one segment of slack
  int32_t nseg;    // = BT_CAND_MAXSEG (row pitch of cnt)
  int32_t seg;     // columns per segment THIS frame: half the association tile width (128 or 112), 128 otherwise
  size_t clear_bytes;  // cnt, total, segmask, rowdeg, indeg live in one allocation: one memset of this many bytes at cnt
  // element strides from one video stream's lists to the next (0: single set)
  size_t s_cnt, s_rowcol, s_deg, s_edges;
};
// the lists of video stream `sid` (all streams' lists are slices of the same allocations)
static __host__ __device__ __forceinline__ bt_cand bt_cand_of(const bt_cand& base, int sid) {
  bt_cand c = base;
  char* p = reinterpret_cast<char*>(base.cnt) + (size_t)sid * base.s_cnt;
  c.cnt = reinterpret_cast<int32_t*>(p);
  c.total = reinterpret_cast<int32_t*>(reinterpret_cast<char*>(base.total) + (size_t)sid * base.s_cnt);
  c.segmask = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(base.segmask) + (size_t)sid * base.s_cnt);
  c.rowdeg = reinterpret_cast<int32_t*>(reinterpret_cast<char*>(base.rowdeg) + (size_t)sid * base.s_cnt);
  c.indeg = reinterpret_cast<int32_t*>(reinterpret_cast<char*>(base.indeg) + (size_t)sid * base.s_cnt);
  c.rowcol = base.rowcol + (size_t)sid * base.s_rowcol;
  c.deg = base.deg + (size_t)sid * base.s_deg;
  c.col = base.col + (size_t)sid * base.s_edges;
  c.cost = base.cost + (size_t)sid * base.s_edges;
  return c;
}

struct bt_assoc_params {
  // operands (row-major, K contiguous).  Tensor-core path: fp16; CUDA-core path: a32/b32 (fp32) or a16/b16
  const __half* a16;   // [a_rows_alloc, d] track-side features (bank rows = global slots)
  const __half* b16;   // [b_rows_alloc, d] detection features
  const float* a32;
  const float* b32;
  int32_t d;
  int32_t bn;                          // association tile width (256 or 224); 0 = 256.  Candidate emission
                                       // requires cand.seg == bn / 2 (btk_assoc_pick_bn)
  int32_t a_rows_alloc, b_rows_alloc;  // rows the operand buffers really hold: lets the TMA descriptors be
                                       // cached across frames; rows past a stream's n / m are masked
  int32_t operands_early;              // the operands were complete before the previous kernel of the stream
                                       // started: the TMA / MMA warps need not wait for it (PDL)
  // the batch: problem k has n[k] rows starting at operand row a_row0[k] and m[k] columns at b_row0[k]
  int32_t count;
  int32_t n[BT_MAX_BATCH], m[BT_MAX_BATCH];
  int32_t a_row0[BT_MAX_BATCH], b_row0[BT_MAX_BATCH];
  int32_t row0[BT_MAX_BATCH], col0[BT_MAX_BATCH];   // first entry of the problem in the row / column side arrays
  int32_t kind_off[BT_MAX_BATCH];                   // byte offset of the problem's row_kind array in row_kind_base
  int32_t cand_sid[BT_MAX_BATCH];                   // which slice of `cand`
  const float* face_sim[BT_MAX_BATCH];              // [n_pool, m] or null
  int32_t pos_off[BT_MAX_BATCH];                    // byte offset of pool_pos (row -> face_sim row) in row_kind_base
  // epilogue inputs (null => not used), indexed by row0[k] + r / col0[k] + c
  const double* row_tlbr;   // [.,4]
  const float* row_tlbr_f32;// [.,4] conservative fp32 interval (lo down, hi up)
  const float* row_norm;    // [.] sim = acc / row_norm (null: 1)
  const char* row_kind_base;
  const double* col_tlbr;   // [.,4]
  const uint8_t* col_kind;  // [.]
  const uint2* col_pk;      // [.] packed 15-bit integer corners (bt_pack16_*), or null: packed in the kernel
  const float* col_norm;    // [.] detection feature norms: unconfirmed rows use sim / col_norm (demo:1593-1599); null: 1
  // thresholds
  double match_thresh, second_thresh, unconf_thresh, proximity;
  float appearance;
  float gate_band;          // half-width of the ambiguous band around the appearance gate (similarity units)
  // outputs
  bt_cand cand;             // candidate emission (cnt==null => off)
  float* out_emb;           // [n,m] 1-max(0,sim)   (dense dump of problem 0, null => off)
  double* out_dists;        // [n,m] fused cost      (dense dump of problem 0, null => off)
  int32_t dense_stage;      // 1 or 3: which fusion rule the dense dump uses
};
// the per-frame part of an association launch, as the kernels read it from DEVICE memory
struct bt_assoc_frame {
  int32_t count;
  int32_t tile_start[BT_MAX_BATCH + 1];   // prefix sums of the problems' tile counts (tensor-core kernel)
  int32_t n[BT_MAX_BATCH], m[BT_MAX_BATCH];
  int32_t a_row0[BT_MAX_BATCH], b_row0[BT_MAX_BATCH];
  int32_t row0[BT_MAX_BATCH], col0[BT_MAX_BATCH];
  int32_t kind_off[BT_MAX_BATCH], pos_off[BT_MAX_BATCH];
  int32_t cand_sid[BT_MAX_BATCH];
  const float* face_sim[BT_MAX_BATCH];
};
// stand-alone: fills a frame description from p, uploads it to the ctx's scratch and launches
int32_t btk_assoc(bt_ctx* ctx, const bt_assoc_params& p, int32_t precision);
// tracker: fill the host copy (goes up with the control block) ...
void btk_assoc_fill(const bt_assoc_params& p, int32_t precision, bt_assoc_frame* out);
// ... and launch against its device copy.  fixed != 0: grid from (max_rows x max_cols) per problem instead of the
// frame's sizes (captured graphs).
int32_t btk_assoc_launch(bt_ctx* ctx, const bt_assoc_params& p, int32_t precision, const bt_assoc_frame& hf,
                         const bt_assoc_frame* df, int fixed, int max_rows, int max_cols);
// tile width that minimises (waves x MMA time per tile) for a batch of n[k] x m[k] problems on this GPU
int32_t btk_assoc_pick_bn(const bt_ctx* ctx, const int32_t* n, const int32_t* m, int32_t count);
void btk_assoc_stamps_report(void);   // BT_ASSOC_DEBUG bit 32768
int32_t bt_gemm_ws_create(bt_ctx* ctx);
void bt_gemm_ws_destroy(bt_ctx* ctx);

// ---- LAP (lap.cu) -------------------------------------------------------------------------------
int32_t bt_lap_ws_create(bt_ctx* ctx);
void bt_lap_ws_destroy(bt_ctx* ctx);
// dense cost -> candidate list `list` of ctx->lap's own bt_cand (video stream 0's slice)
int32_t btk_lap_compact_dense(bt_ctx* ctx, const double* cost, int32_t n, int32_t m, double thresh,
                              const bt_cand& cand, int32_t list);
// Solve list `list`: rows [0,n), cols [0,m).  x[n], y[m] outputs (device).
int32_t btk_lap_solve(bt_ctx* ctx, const bt_cand& cand, int32_t list, int32_t n, int32_t m,
                      double thresh, int32_t* x, int32_t* y);
const bt_cand* bt_lap_own_cand(bt_ctx* ctx);
// Exact re-costing of flagged edges inside the LAP (SURVEY hard part 2): what it needs to evaluate a
// (row, column) pair from scratch in fp32 / fp64.
struct bt_refine {
  int32_t enabled, d, f16;                 // f16: features are the raw fp16 rows + norms; else fp32 rows
  const __half* a16; const float* a_norm;  // per global slot
  const float* a32;                        // per global slot (normalised fp32 current feature)
  const __half* b16; const float* b32;     // per global detection of the frame's parity
  const float* b_norm;                     // per global detection (stage 3 only)
  const double* row_tlbr; const double* col_tlbr;
  const char* ctrl;                        // control block (pool_pos for the face term)
  double proximity; float appearance;
};
// the three chained association stages of a frame, one CTA per video stream, ONE launch (lists 0,1,2)
struct bt_lap_batch {
  int32_t count;
  int32_t sid[BT_MAX_BATCH], n[BT_MAX_BATCH], m[BT_MAX_BATCH];
  int32_t row0[BT_MAX_BATCH], col0[BT_MAX_BATCH];       // global slot / detection of row 0 / column 0
  int32_t in0[BT_MAX_BATCH];                            // row of column 0 in the (double-buffered) detection feature buffers
  int32_t pos_off[BT_MAX_BATCH];
  const float* face_sim[BT_MAX_BATCH];
  int32_t* x[BT_MAX_BATCH];         // x1,x2,x3 [n_x_stride each]
  int32_t x_stride[BT_MAX_BATCH];
  int32_t* y[BT_MAX_BATCH];         // y1,y2,y3 [y_stride each]
  int32_t y_stride;
  int32_t* zero_word[BT_MAX_BATCH]; // device word to clear (the frame's duplicate-pair counter)
  // direct results: the LAP kernel copies the stream's assignment vectors into pinned host memory (same layout as x)
  // and then sets the stream's flag word there -- the host polls it instead of waiting for a D2H copy + event
  int32_t* hx[BT_MAX_BATCH];
  uint32_t* hflag[BT_MAX_BATCH];
};
// atomic_emitter != 0: the lists were filled by the CUDA-core kernel (atomic appends: its segment counters must be
// left zeroed for the next frame)
int32_t btk_lap_solve3(bt_ctx* ctx, const bt_cand& cand, const bt_lap_batch& b, const bt_lap_batch* db,
                       const double thresh[3], const bt_refine& rf, int atomic_emitter);

// ---- detector side ------------------------------------------------------------------------------
int32_t btk_yolox_postprocess(bt_ctx* ctx, const float* raw, const bt_yolox_config& cfg,
                              double* out_boxes, int32_t max_out, int32_t* out_count);
int32_t btk_reid_crop_gather(bt_ctx* ctx, const uint8_t* frame, int32_t h, int32_t w,
                             const int32_t* boxes, int32_t n, int32_t out_h, int32_t out_w, float* out);

// ---- tracker ------------------------------------------------------------------------------------
int32_t bt_tracker_create(bt_ctx* ctx);
void bt_tracker_destroy(bt_ctx* ctx);
