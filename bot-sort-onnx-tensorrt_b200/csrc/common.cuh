// Internal declarations shared by the translation units of libbotsort_b200.so.
// Not part of the ABI (that is include/botsort_b200.h).
#pragma once

#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/botsort_b200.h"

struct bt_tracker;  // track_step.cu

struct bt_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  int max_tracks = 0, max_dets = 0, feat_dim = 0;
  uint32_t flags = 0;
  int num_sms = 148;
  int pdl = 1;   // programmatic dependent launch between the frame step's kernels (BT_NO_PDL=1 turns it off)
  std::string err;
  int64_t launches = 0;
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
// Launch on ctx->stream, optionally as a programmatic dependent of the stream's previous kernel: the grid
// may then be scheduled before that kernel has completed, and must execute griddepcontrol.wait
// (bt_grid_dependency_wait) before it touches anything the earlier kernels of the stream produce -- and at
// the latest before it exits, so that completion stays ordered for its own dependents.
template <typename... KArgs, typename... Args>
static inline cudaError_t bt_launch(bt_ctx* ctx, bool dependent, void (*kern)(KArgs...), dim3 grid, dim3 block,
                                    size_t smem, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = ctx->stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = (dependent && ctx->pdl) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
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

// ---- kernel launchers (device pointers only; enqueue on ctx->stream) --------------------------
// kalman.cu
int32_t btk_kalman_initiate(bt_ctx* ctx, const float* xywh, const int32_t* src_idx, double* mean,
                            double* cov, double* tlbr, float* tlbr_f32, const int32_t* dst_idx, int32_t k,
                            uint8_t* slot_f32 = nullptr);
int32_t btk_kalman_predict(bt_ctx* ctx, double* mean, double* cov, double* tlbr, float* tlbr_f32,
                           const int32_t* state, const int32_t* idx, int32_t n, int32_t noise_f32,
                           uint8_t* slot_f32 = nullptr);
int32_t btk_kalman_update(bt_ctx* ctx, double* mean, double* cov, double* tlbr, float* tlbr_f32,
                          const double* meas, const int32_t* track_idx, const int32_t* meas_idx,
                          const uint8_t* noise_f32, int32_t k);
// tracker mode: slot g takes the detection x1[g] / x2[g] / x3[g] (first non-negative) as measurement
int32_t btk_kalman_update_x(bt_ctx* ctx, double* mean, double* cov, double* tlbr, float* tlbr_f32,
                            const double* meas, const int32_t* x1, const int32_t* x2, const int32_t* x3,
                            uint8_t* slot_f32, int32_t n_slots, double* res_tlbr = nullptr);
int32_t btk_kalman_project(bt_ctx* ctx, const double* mean, const double* cov, double* pmean,
                           double* pcov, int32_t n);
// features.cu: EMA of every slot matched by one of the three stages (x arrays), fp16 bank refresh
int32_t btk_feature_ema_x(bt_ctx* ctx, float* smooth, float* curr, const float* feat, __half* bank16,
                          const __half* det16, const int32_t* x1, const int32_t* x2, const int32_t* x3,
                          int32_t n_slots, int32_t d, float alpha);
// iou.cu: pairs (i < j) of live slots (kind != 0) whose IoU distance is below `limit`
// pair_count and the first small_cap pairs also land in the frame's result block (one D2H per frame)
int32_t btk_iou_pairs_live(bt_ctx* ctx, const double* tlbr, const float* tlbr_f32, const uint8_t* kind, int32_t n,
                           double limit, int32_t* pairs, int32_t* pair_count, int32_t pair_cap,
                           int32_t* pairs_small = nullptr, int32_t small_cap = 0);
// iou.cu
int32_t btk_iou_distance(bt_ctx* ctx, const double* a, int32_t n, const double* b, int32_t m,
                         double* out);
int32_t btk_fuse_score(bt_ctx* ctx, const double* d, const double* s, int32_t n, int32_t m, double* out);
// emits (i, j) pairs with 1-IoU < limit between lists given by slot index; count via atomic
int32_t btk_iou_pairs_below(bt_ctx* ctx, const double* tlbr, const int32_t* a_idx, int32_t n,
                            const int32_t* b_idx, int32_t m, double limit, int32_t* pairs,
                            int32_t* pair_count, int32_t pair_cap);

// features.cu
// det_prep: normalise rows, write fp32 normalised copy (optional) + fp16 copy; also boxes -> tlbr/xywh
int32_t btk_feature_prep(bt_ctx* ctx, const float* feat, int32_t m, int32_t d, float* out_f32,
                         __half* out_f16, int32_t normalise, int32_t dependent = 0);
int32_t btk_feature_ema(bt_ctx* ctx, float* smooth, float* curr, const float* feat,
                        const int32_t* track_idx, const int32_t* feat_idx, const uint8_t* first,
                        int32_t k, int32_t d, float alpha);
// same, and also refreshes the fp16 bank row (GEMM A operand): bank16[track_idx[i]] = det16[feat_idx[i]]
int32_t btk_feature_ema16(bt_ctx* ctx, float* smooth, float* curr, const float* feat, __half* bank16,
                          const __half* det16, const int32_t* track_idx, const int32_t* feat_idx,
                          const uint8_t* first, int32_t k, int32_t d, float alpha);

// ---- association (reid_gemm.cu) ----------------------------------------------------------------
// Row kinds / column kinds of the fused association kernel.
enum { BT_ROW_NONE = 0, BT_ROW_POOL_TRACKED = 1, BT_ROW_POOL_OTHER = 2, BT_ROW_UNCONFIRMED = 3 };
enum { BT_COL_NONE = 0, BT_COL_HIGH = 1, BT_COL_LOW = 2 };

// Candidate lists ("ragged dense" adjacency): list s in {0,1,2} = association stage 1,2,3.
// Row r owns entries [r*stride, r*stride + cnt[s*rows_cap + r]) of col/cost.
// A row's region of `stride` entries is cut into segments of `seg` columns: the
// edges of row r found in columns [g*seg, (g+1)*seg) are written to
// [r*stride + g*seg, r*stride + g*seg + cnt[list][r][g]).  The tensor-core epilogue owns one
// (row, segment) pair per thread, so it appends with plain stores and a register counter -- no
// atomics; the LAP set-up compacts the segments of a row to the front of its region.
#define BT_CAND_MAXSEG 64   // segments per row (bits of segmask); cnt row pitch
struct bt_cand {
  int32_t* cnt;    // [3][rows_cap][nseg]
  int32_t* total;  // [4] per list: non-zero when any edge was emitted (lets the LAP skip empty stages)
  unsigned long long* segmask;  // [3][rows_cap] bit g: segment g of the row is non-empty (nseg <= 64)
  int32_t* rowdeg;  // [3][rows_cap] edges emitted for the row      } maintained by the emitters with fire-and-forget
  int32_t* indeg;   // [3][cols_cap] edges emitted into the column  } atomics: the LAP classifies rows without
  int32_t* rowcol;  // [3][rows_cap] a column of the row (THE column when rowdeg == 1)   touching the edge lists
  int32_t cols_cap;
  int32_t* deg;    // [3][rows_cap]  row degree after compaction (written by the LAP set-up)
  int32_t* col;    // [3][rows_cap*stride]
  double* cost;    // [3][rows_cap*stride]
  int32_t rows_cap;
  int32_t stride;  // entries per row >= max_dets + one segment of slack
  int32_t nseg;    // = BT_CAND_MAXSEG (row pitch of cnt)
  int32_t seg;     // columns per segment THIS frame: half the association tile width (128 or 112), 128 otherwise
  size_t clear_bytes;  // cnt, total and segmask live in one allocation: one memset of this many bytes at cnt
};

struct bt_assoc_params {
  // operands
  const __half* a16;   // [n, d] track-side features (bank rows = slots)
  const __half* b16;   // [m, d] detection features
  const float* a32;    // fp32 variants for the SIMT kernel (may be null when tensor path is used)
  const float* b32;
  int32_t n, m, d;
  int32_t bn;                          // association tile width (256 or 224); 0 = 256.  Candidate emission
                                       // requires cand.seg == bn / 2 (btk_assoc_pick_bn)
  int32_t a_rows_alloc, b_rows_alloc;  // rows the operand buffers really hold (0 = n / m): lets the TMA
                                       // descriptors be cached across frames; rows >= n / m are masked
  // epilogue inputs (null => not used)
  const double* row_tlbr;   // [n,4]
  const float* row_tlbr_f32;// [n,4] conservative fp32 interval (lo down, hi up)
  const uint8_t* row_kind;  // [n]
  const double* col_tlbr;   // [m,4]
  const uint8_t* col_kind;  // [m]
  const uint2* col_pk;      // [m] packed 15-bit integer corners (bt_pack16_*), or null: packed in the kernel
  const float* face_sim;    // [n,m] or null
  // thresholds
  double match_thresh, second_thresh, unconf_thresh, proximity;
  float appearance;
  // outputs
  bt_cand cand;             // candidate emission (cnt==null => off)
  float* out_emb;           // [n,m] 1-max(0,sim)   (dense dump, null => off)
  double* out_dists;        // [n,m] fused cost      (dense dump, null => off)
  int32_t dense_stage;      // 1 or 3: which fusion rule the dense dump uses
};
int32_t btk_assoc(bt_ctx* ctx, const bt_assoc_params& p, int32_t precision);
// tile width that minimises (waves x MMA time per tile) for an n x m problem on this GPU
int32_t btk_assoc_pick_bn(const bt_ctx* ctx, int32_t n, int32_t m);
int32_t bt_gemm_ws_create(bt_ctx* ctx);
void bt_gemm_ws_destroy(bt_ctx* ctx);

// ---- LAP (lap.cu) -------------------------------------------------------------------------------
int32_t bt_lap_ws_create(bt_ctx* ctx);
void bt_lap_ws_destroy(bt_ctx* ctx);
// dense cost -> candidate list `list` of ctx->lap's own bt_cand
int32_t btk_lap_compact_dense(bt_ctx* ctx, const double* cost, int32_t n, int32_t m, double thresh,
                              const bt_cand& cand, int32_t list);
// Solve list `list`: rows [0,n), cols [0,m).  row_block/col_block: optional int32 arrays, an edge
// is valid only if row_block[r] < 0 and col_block[c] < 0 (results of an earlier stage).
// x[n], y[m] outputs (device).
int32_t btk_lap_solve(bt_ctx* ctx, const bt_cand& cand, int32_t list, int32_t n, int32_t m,
                      double thresh, const int32_t* row_block, const int32_t* col_block, int32_t* x,
                      int32_t* y);
const bt_cand* bt_lap_own_cand(bt_ctx* ctx);
// the three chained association stages of a frame in ONE launch (lists 0,1,2)
int32_t btk_lap_solve3(bt_ctx* ctx, const bt_cand& cand, int32_t n, int32_t m, const double thresh[3],
                       int32_t* const x[3], int32_t* const y[3], int32_t* zero_word = nullptr);

// ---- detector side ------------------------------------------------------------------------------
int32_t btk_yolox_postprocess(bt_ctx* ctx, const float* raw, const bt_yolox_config& cfg,
                              double* out_boxes, int32_t max_out, int32_t* out_count);
int32_t btk_reid_crop_gather(bt_ctx* ctx, const uint8_t* frame, int32_t h, int32_t w,
                             const int32_t* boxes, int32_t n, int32_t out_h, int32_t out_w, float* out);

// ---- tracker ------------------------------------------------------------------------------------
int32_t bt_tracker_create(bt_ctx* ctx);
void bt_tracker_destroy(bt_ctx* ctx);
