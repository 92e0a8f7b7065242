// Fused association kernel: ReID similarity GEMM (tracks x detections x 2048-d) on the 5th-gen
// tensor cores with the whole cost fusion of the reference in its epilogue.
//
// Reference arithmetic (demo = /root/reference/demo_bottrack_onnx_tflite.py):
//   sim   = f_trk . f_det^T            in-graph cosine, README.md:185-195, consumed demo:1453-1460
//   stage 1 (demo:1539-1554): emb = 1 - sim; emb[min(emb, face_emb) > 0.25] = 1;
//                             dists = min(iou_dist, emb)            -> linear_assignment(0.8)
//   stage 2 (demo:1568-1571): dists = iou_dist (Tracked rows x low-score dets) -> thresh 0.5
//   stage 3 (demo:1593-1604): emb = 1 - max(0, sim); emb[emb > 0.25] = 1; emb[iou_d > 0.5] = 1;
//                             dists = min(iou_dist, emb)            -> thresh 0.7
//   iou_dist = 1 - bbox_iou (demo:1695-1713), float64.
//
// B200 design: A = per-slot fp16 feature bank [n, d] (K-major), B = per-frame fp16 detection
// features [m, d] (K-major).  One persistent CTA per SM walks 128 x BN output tiles:
//   warp 0   : TMA producer (cp.async.bulk.tensor, 128B swizzle, 4-stage mbarrier ring); first copies the
//              launch description to shared memory with one batch of loads
//   warp 1   : tcgen05.mma issue -- the warp stays converged, elect.sync predicates the MMA / commit
//              (see tcgen05_mma_f16), fp32 accumulators in TMEM, 2 stages
//   warps 2-9: epilogue -- integer box pass in the shadow of the main loop, then tcgen05.ld the accumulator,
//              apply the fusion rules above and emit only the *candidate edges* (cost < stage threshold,
//              flagged when their cost carries the tensor-core error) into per-row lists.
// The N x M matrices never go to HBM on the tracker path (they are written only for the
// stand-alone entry points / parity dumps): a 2000 x 2000 frame reads 2*8.2 MB of fp16
// features and writes a few thousand edges.  Almost every pair is rejected by an fp32 test
// (no box overlap possible AND appearance gate closed), the exact float64 IoU runs only for
// the survivors.
#include "common.cuh"

#include <cuda.h>
#include <math.h>
#include <type_traits>
#include <mutex>
#include <string.h>
#include <stdlib.h>

namespace {

// ------------------------------------------------------------------------------------------------
// epilogue element logic shared by the tensor-core and the CUDA-core kernels
// ------------------------------------------------------------------------------------------------
struct EpiParams {
  // the batch (one entry per video stream / stand-alone problem): sizes, operand / side-array offsets, tile
  // prefix sums -- read from DEVICE memory, so that the kernel arguments are the same every frame
  const bt_assoc_frame* F;
  // side inputs
  const double* row_tlbr;
  const float* row_tlbr_f32;
  const float* row_norm;
  const char* row_kind_base;
  const double* col_tlbr;
  const uint8_t* col_kind;
  const uint2* col_pk;
  const float* col_norm;
  double match_thresh, second_thresh, unconf_thresh, proximity;
  float appearance;
  bt_cand cand;
  float* out_emb;
  double* out_dists;
  int dense_stage;
  int operands_early;
  // single-problem launches outside graph capture: the first tile's coordinates as kernel arguments, so that the
  // producer's first TMA requests do not wait for the launch description's round trip (0 tiles_n = no hint)
  int hint_tiles_n, hint_num_tiles, hint_a_row0, hint_b_row0;
  // BT_ASSOC_DEBUG bits (profiling / bisection aids of the tensor-core kernel, device printf):
  //   1 sampled CTA timestamps   2 similarity pass off   4 every CTA's timestamps   16 epilogue phase times
  //   32 no box pass (every pair evaluated by the similarity pass)   64 no early TMA prologue
  //   1024 globaltimer go/end of every epilogue warp   4096 no operand traffic after the prologue (the producer
  //   marks stages full without a TMA: stale tiles, WRONG results -- measures what the loads cost the main loop)
  //   32768 per-SM globaltimer stamps at CTA entry / exit, summarised by bt_profile_replay_assoc
  int debug;
  float sim_gate;  // smallest similarity for which the appearance gate is open
  float gate_band; // |sim - sim_gate| <= gate_band: the gate decision is within the tensor-core error (AMBIG)
  float iou_gate;  // a pair without open appearance gate needs IoU > 1 - max(stage thresholds); slightly lowered
};

#define iou_dist_f64 bt_iou_dist_f64
#define fuse_stage1 bt_fuse_stage1
#define fuse_stage3 bt_fuse_stage3


// ---- candidate-list addressing for video stream `sid` (all slices of the same allocations) --------
__device__ __forceinline__ int32_t* c_cnt(const EpiParams& p, int sid) {
  return reinterpret_cast<int32_t*>(reinterpret_cast<char*>(p.cand.cnt) + (size_t)sid * p.cand.s_cnt);
}
__device__ __forceinline__ int32_t* c_total(const EpiParams& p, int sid) {
  return reinterpret_cast<int32_t*>(reinterpret_cast<char*>(p.cand.total) + (size_t)sid * p.cand.s_cnt);
}
__device__ __forceinline__ unsigned long long* c_segmask(const EpiParams& p, int sid) {
  return reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(p.cand.segmask) + (size_t)sid * p.cand.s_cnt);
}
__device__ __forceinline__ int32_t* c_rowdeg(const EpiParams& p, int sid) {
  return reinterpret_cast<int32_t*>(reinterpret_cast<char*>(p.cand.rowdeg) + (size_t)sid * p.cand.s_cnt);
}
__device__ __forceinline__ int32_t* c_indeg(const EpiParams& p, int sid) {
  return reinterpret_cast<int32_t*>(reinterpret_cast<char*>(p.cand.indeg) + (size_t)sid * p.cand.s_cnt);
}
__device__ __forceinline__ int32_t* c_rowcol(const EpiParams& p, int sid) { return p.cand.rowcol + (size_t)sid * p.cand.s_rowcol; }
__device__ __forceinline__ int32_t* c_col(const EpiParams& p, int sid) { return p.cand.col + (size_t)sid * p.cand.s_edges; }
__device__ __forceinline__ double* c_cost(const EpiParams& p, int sid) { return p.cand.cost + (size_t)sid * p.cand.s_edges; }

// face similarity of (row, col) of problem k (0 when the stream has no face term)
__device__ __forceinline__ float face_sim_of(const EpiParams& p, int k, int row, int col) {
  const float* fs = p.F->face_sim[k];
  if (!fs) return 0.0f;
  int pr = row;
  if (p.row_kind_base) pr = reinterpret_cast<const int32_t*>(p.row_kind_base + p.F->pos_off[k])[row];
  return pr >= 0 ? fs[(size_t)pr * p.F->m[k] + col] : 0.0f;
}

// atomic append into the (row, column-segment) sub-list (CUDA-core kernel: several threads share a row)
__device__ __forceinline__ void emit(const EpiParams& p, int sid, int list, int row, int col, double cost) {
  const bt_cand& c = p.cand;
  const int seg = col / c.seg;
  const size_t ri = (size_t)list * c.rows_cap + row;
  const int k = atomicAdd(c_cnt(p, sid) + ri * c.nseg + seg, 1);
  const size_t base = ri * (size_t)c.stride + (size_t)seg * c.seg + k;
  c_col(p, sid)[base] = col;
  c_cost(p, sid)[base] = cost;
  atomicAdd(&c_total(p, sid)[list], 1);
  if (k == 0) atomicOr(&c_segmask(p, sid)[ri], 1ull << seg);
  atomicAdd(&c_rowdeg(p, sid)[ri], 1);
  atomicAdd(&c_indeg(p, sid)[(size_t)list * c.cols_cap + col], 1);
  c_rowcol(p, sid)[ri] = col;
}
// tensor-core kernel: the owner of (row, segment) adds the row degree once, at the end of its tile
__device__ __forceinline__ void emit_owned_tc(const EpiParams& p, int sid, int list, int row, int seg, int k, int col,
                                              int flags, double cost) {
  const bt_cand& c = p.cand;
  const size_t base = ((size_t)list * c.rows_cap + row) * (size_t)c.stride + (size_t)seg * c.seg + k;
  c_col(p, sid)[base] = col | flags;
  c_cost(p, sid)[base] = cost;
  atomicAdd(&c_indeg(p, sid)[(size_t)list * c.cols_cap + col], 1);
}

// exact path of the CUDA-core kernel for one (row, col) pair that survived the cheap rejection test
__device__ __noinline__ void assoc_exact(const EpiParams& p, int k, int row, int col, float sim, int rkind, int ckind) {
  const int sid = p.F->cand_sid[k];
  const double iou_d = iou_dist_f64(p.row_tlbr + ((size_t)p.F->row0[k] + row) * 4, p.col_tlbr + ((size_t)p.F->col0[k] + col) * 4);
  if (rkind == BT_ROW_UNCONFIRMED) {
    if (ckind == BT_COL_HIGH) {
      const double c3 = fuse_stage3(iou_d, sim, p.appearance, p.proximity);
      if (c3 < p.unconf_thresh) emit(p, sid, 2, row, col, c3);
    }
  } else {
    if (ckind == BT_COL_HIGH) {
      const double c1 = fuse_stage1(iou_d, sim, face_sim_of(p, k, row, col), p.appearance);
      if (c1 < p.match_thresh) emit(p, sid, 0, row, col, c1);
    } else if (ckind == BT_COL_LOW && rkind == BT_ROW_POOL_TRACKED) {
      if (iou_d < p.second_thresh) emit(p, sid, 1, row, col, iou_d);
    }
  }
}

// Edge flags of a pair whose similarity has been looked at (tensor-core kernel).  SIM: the fused cost may be
// the embedding distance, i.e. it carries the fp16 tensor-core error and is re-costed exactly by the LAP when
// the row competes with others.  AMBIG: the appearance gate itself is within that error of flipping.
__device__ __forceinline__ int sim_flags(const EpiParams& p, int list, float sim, float face) {
  if (list == 1) return 0;
  const bool face_open = (list == 0) && !((1.0f - face) > p.appearance);
  const bool open = sim >= p.sim_gate;
  int f = 0;
  if (open || face_open) f |= BT_EDGE_SIM;
  if (!face_open && fabsf(sim - p.sim_gate) <= p.gate_band) f |= BT_EDGE_AMBIG | BT_EDGE_SIM;
  return f;
}

// Similarity pass of the tensor-core kernel for a pair that is not in the warp's shared-memory slots:
// evaluates it from scratch; if the box pass spilled edges of this row straight to the list (more than
// kPreSlots of them) the pair is looked up there first and re-costed in place.
// Returns col | flags when a new edge was appended at position `k_next` of the caller's (row, seg) sub-list, else -1.
__device__ __noinline__ int open_pair_slow(const EpiParams& p, int k, double r0, double r1, double r2, double r3,
                                           const double* __restrict__ cbox, int list, int row, int seg, int col,
                                           float sim, int k_next, int scan_from) {
  const int sid = p.F->cand_sid[k];
  const double rb[4] = {r0, r1, r2, r3};
  const size_t eb = ((size_t)list * p.cand.rows_cap + row) * (size_t)p.cand.stride + (size_t)seg * p.cand.seg;
  int32_t* ecol = c_col(p, sid);
  double* ecost = c_cost(p, sid);
  int k_found = -1;
  if (scan_from >= 0)
    for (int j = scan_from; j < k_next; ++j)
      if ((ecol[eb + j] & BT_EDGE_COLMASK) == col) k_found = j;
  const double iou_d = iou_dist_f64(rb, cbox);
  double cost = iou_d;
  const float face = (list == 0) ? face_sim_of(p, k, row, col) : 0.0f;
  if (list == 0) cost = fuse_stage1(iou_d, sim, face, p.appearance);
  else if (list == 2) cost = fuse_stage3(iou_d, sim, p.appearance, p.proximity);
  const int flags = sim_flags(p, list, sim, face);
  if (k_found >= 0) {
    if (cost != iou_d) ecost[eb + k_found] = cost;
    if (flags) ecol[eb + k_found] = col | flags;
    return -1;
  }
  const double thr = list == 0 ? p.match_thresh : (list == 1 ? p.second_thresh : p.unconf_thresh);
  if (!(cost < thr) && !(flags & BT_EDGE_AMBIG)) return -1;
  emit_owned_tc(p, sid, list, row, seg, k_next, col, flags, cost);
  return col | flags;
}

// dense dumps (stand-alone entry points): problem 0
__device__ __forceinline__ void assoc_dense(const EpiParams& p, int row, int col, float sim) {
  const int m = p.F->m[0];
  if (p.out_emb) p.out_emb[(size_t)row * m + col] = 1.0f - fmaxf(0.0f, sim);
  if (p.out_dists) {
    const double iou_d = iou_dist_f64(p.row_tlbr + (size_t)row * 4, p.col_tlbr + (size_t)col * 4);
    const float face = p.F->face_sim[0] ? p.F->face_sim[0][(size_t)row * m + col] : 0.0f;
    p.out_dists[(size_t)row * m + col] = (p.dense_stage == 3)
                                             ? fuse_stage3(iou_d, sim, p.appearance, p.proximity)
                                             : fuse_stage1(iou_d, sim, face, p.appearance);
  }
}

// ------------------------------------------------------------------------------------------------
// PTX wrappers (sm_100a)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tmap, int x, int y,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(x), "r"(y)
      : "memory");
}
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
// The MMA warp stays CONVERGED: every lane executes the main loop and one elected lane (always the same one: the
// lowest) issues.  Under a divergent `if (lane == 0)` ptxas wraps every UTCHMMA / UTCBAR in an ELECT / R2UR.BROADCAST
// / BRA.U.ANY loop (it cannot prove the operands uniform): ~60 cycles of issue-thread time per MMA, more than the
// tensor pipe needs for the MMA itself once the loop's barrier wait and commit are added -- the main loop was
// bound by its own issue thread (profiles/README.md, round 2).
__device__ __forceinline__ void tcgen05_commit(uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tcgen05_mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                                uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
// float64 box -> packed corners (common.cuh) through a directed-rounding fp32 interval: 4 conversions
__device__ __forceinline__ uint2 pack16_box(double x1, double y1, double x2, double y2, bool is_col) {
  return bt_pack16_f32(__double2float_rd(x1), __double2float_rd(y1), __double2float_ru(x2), __double2float_ru(y2), is_col);
}
// tcgen05.wait::ld with the loaded registers as in/out operands so that the compiler cannot move a use of them
// above the wait when loads are software-pipelined.
__device__ __forceinline__ void tmem_ld_wait_dep(uint32_t (&v)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                 "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]),
                 "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]),
                 "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
               :
               : "memory");
}

// K-major operand tile in shared memory, 128-byte swizzle (as written by TMA SWIZZLE_128B):
// rows of 64 fp16 = 128 B, 8-row groups 1024 B apart.  sm_100 descriptor: start>>4 [0,14),
// LBO>>4 [16,30) (unused for swizzled K-major, 1), SBO>>4 [32,46) = 1024>>4, version 1 at [46,48),
// layout SWIZZLE_128B = 2 at [61,64).
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// ------------------------------------------------------------------------------------------------
// tensor-core kernel
// ------------------------------------------------------------------------------------------------
// BT_ASSOC_DEBUG bit 32768: every CTA appends (globaltimer at entry, at exit) to its SM's list -- one CTA per SM at a
// time, so no races; bt_profile_replay_assoc prints per-SM lifetimes and hand-over gaps between consecutive launches.
constexpr int kStampSMs = 160, kStampDepth = 64;
__device__ unsigned long long g_assoc_stamps[kStampSMs][kStampDepth][2];
__device__ int g_assoc_stamp_cnt[kStampSMs];
constexpr int BM = 128;
constexpr int BK = 64;       // 64 fp16 = one 128 B swizzle row
constexpr int UMMA_K = 16;
constexpr int kStages = 4;
constexpr int kAccStages = 2;
constexpr int kEpiWarps = 8;     // two per SM sub-partition: warp w and w+4 share TMEM lane quarter w%4
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kPreSlots = 4;     // pairs per row and tile half whose box-pass result is remembered in shared memory
constexpr int kTcThreads = 64 + kEpiThreads;  // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue

template <int BN>
struct TcSmem {
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kColPkOff = kStages * kStageBytes;             // uint2[BN]  packed integer det corners
  static constexpr int kColKindOff = kColPkOff + BN * 8;              // uint8[BN]
  static constexpr int kColInvOff = kColKindOff + 256;                // float[BN] 1 / detection feature norm (unconfirmed rows)
  static constexpr int kColF64Off = kColInvOff + BN * 4;              // double[BN][4] det box float64 (exact path)
  static constexpr int kPreIouOff = kColF64Off + BN * 32;             // double[kEpiWarps][kPreSlots][32] IoU distance of box-pass edges
  static constexpr int kPreKeyOff = kPreIouOff + 8 * 4 * 32 * 8;      // uint32[kEpiWarps][kPreSlots][32] their (column, list, flags)
  static constexpr int kBarOff = kPreKeyOff + 8 * 4 * 32 * 4;         // barriers (8B aligned)
  static constexpr int kNumBars = 2 * kStages + 2 * kAccStages;
  static constexpr int kTmemPtrOff = kBarOff + kNumBars * 8;
  static constexpr int kFrameOff = kTmemPtrOff + 16;                   // bt_assoc_frame: the launch description, copied once
  static constexpr int kTotal = kFrameOff + (int)sizeof(bt_assoc_frame);
  static constexpr int kDyn = kTotal + 1024;  // slack for manual 1024 B alignment
};

// tile t of the batch -> problem k and its tile coordinates
__device__ __forceinline__ void tile_of(const bt_assoc_frame* F, int t, int bn, int& k, int& m0, int& n0) {
  k = 0;
  while (k + 1 < F->count && t >= F->tile_start[k + 1]) ++k;
  const int local = t - F->tile_start[k];
  const int tiles_n = (F->m[k] + bn - 1) / bn;
  m0 = (local / tiles_n) * BM;
  n0 = (local % tiles_n) * bn;
}

// One persistent CTA per SM walks the 128 x BN output tiles of every problem of the batch (video streams
// are a leading dimension: tile -> (stream, tile row, tile column)); with more tiles than SMs the second
// TMEM accumulator stage lets tile i's epilogue run under tile i+1's main loop.
template <int BN, bool kDense>
// 10 warps = 3 + 3 + 2 + 2 per SM sub-partition of 16 K registers: 168 registers per thread is the ceiling
__global__ void __launch_bounds__(kTcThreads, 1)
assoc_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                const __grid_constant__ EpiParams p, int d) {   // grid constant: &p (open_pair_slow) needs no local copy
  using L = TcSmem<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOff);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full = empty_bar + kStages;
  uint64_t* tmem_empty = tmem_full + kAccStages;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + L::kTmemPtrOff);
  uint2* s_colpk = reinterpret_cast<uint2*>(smem + L::kColPkOff);
  uint8_t* s_colkind = smem + L::kColKindOff;
  float* s_colinv = reinterpret_cast<float*>(smem + L::kColInvOff);
  double* s_preiou = reinterpret_cast<double*>(smem + L::kPreIouOff);
  uint32_t* s_prekey = reinterpret_cast<uint32_t*>(smem + L::kPreKeyOff);
  double* s_col64 = reinterpret_cast<double*>(smem + L::kColF64Off);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long t_start = clock64();
  unsigned long long g_start = 0;
  if (p.debug) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g_start));
  int stamp_sm = -1, stamp_i = 0;
  if ((p.debug & 32768) && threadIdx.x == 0) {
    asm volatile("mov.u32 %0, %%smid;" : "=r"(stamp_sm));
    if (stamp_sm < kStampSMs) { stamp_i = g_assoc_stamp_cnt[stamp_sm]++ % kStampDepth; g_assoc_stamps[stamp_sm][stamp_i][0] = g_start; } else stamp_sm = -1;
  }
  const int num_kb = d / BK;
  const int first_tile = blockIdx.x, tile_step = gridDim.x;
  // The launch description lives in device memory (same kernel arguments every frame): warp 0 copies it to
  // shared memory with ONE batch of independent loads -- walking it in place cost the producer four dependent
  // global round trips (count -> tile_start -> m -> operand rows) before the first TMA request.
  bt_assoc_frame* sF = reinterpret_cast<bt_assoc_frame*>(smem + L::kFrameOff);
  // Programmatic dependent launch: the next kernel of the stream may be scheduled as soon as SMs free
  // up (it blocks in its own griddepcontrol.wait until this grid has completed and flushed).
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  int pro_kb = 0;   // k-blocks of the first tile already requested by the prologue below
  // the pipeline's first kStages loads do not wait for anybody: they are requested before the TMEM allocation and
  // the CTA-wide sync so that their latency overlaps the rest of the ramp
  auto prologue_loads = [&](const int ya, const int yb) {
    pro_kb = num_kb < kStages ? num_kb : kStages;
    for (int kb = 0; kb < pro_kb; ++kb) {
      mbar_expect_tx(&full_bar[kb], L::kStageBytes);
      uint8_t* sa = smem + kb * L::kStageBytes;
      tma_load_2d(sa, &tmap_a, kb * BK, ya, &full_bar[kb]);
      tma_load_2d(sa + L::kABytes, &tmap_b, kb * BK, yb, &full_bar[kb]);
    }
  };
  const bool hinted = p.hint_tiles_n > 0 && !(p.debug & 64);
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_a)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_b)) : "memory");
    for (int i = 0; i < kStages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    // The operands (feature bank, detection features) are complete before the previous kernel of the
    // stream even starts when the host says so (operands_early): the main loop then runs under that
    // kernel; only the epilogue (boxes, kinds, candidate lists) waits for it.
    if (!p.operands_early) asm volatile("griddepcontrol.wait;" ::: "memory");
    if (hinted && first_tile < p.hint_num_tiles)        // tile coordinates from the kernel arguments: no global round trip
      prologue_loads(p.hint_a_row0 + (first_tile / p.hint_tiles_n) * BM, p.hint_b_row0 + (first_tile % p.hint_tiles_n) * BN);
  }
  if (warp == 0) {
    __syncwarp();
    constexpr int kWords = (int)(sizeof(bt_assoc_frame) / 8);
    static_assert(sizeof(bt_assoc_frame) % 8 == 0, "frame description is copied in 8-byte words");
    const uint2* src = reinterpret_cast<const uint2*>(p.F);
    uint2 v[(kWords + 31) / 32];
#pragma unroll
    for (int i = 0; i < (kWords + 31) / 32; ++i) v[i] = (lane + 32 * i < kWords) ? src[lane + 32 * i] : make_uint2(0u, 0u);
#pragma unroll
    for (int i = 0; i < (kWords + 31) / 32; ++i)
      if (lane + 32 * i < kWords) reinterpret_cast<uint2*>(sF)[lane + 32 * i] = v[i];
    __syncwarp();
    if (lane == 0 && !hinted && first_tile < sF->tile_start[sF->count] && !(p.debug & 64)) {
      int k, m0, n0;
      tile_of(sF, first_tile, BN, k, m0, n0);
      prologue_loads(sF->a_row0[k] + m0, sF->b_row0[k] + n0);
    }
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < kAccStages; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], kEpiWarps); }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    // whole warp: allocate all 512 TMEM columns (2 accumulator stages x BN<=256 fp32 columns)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const int num_tiles = sF->tile_start[sF->count];

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int stage = pro_kb % kStages;
      uint32_t phase = pro_kb == kStages ? 1u : 0u;
      for (int t = first_tile; t < num_tiles; t += tile_step) {
        int k, m0, n0;
        tile_of(sF, t, BN, k, m0, n0);
        const int ya = sF->a_row0[k] + m0, yb = sF->b_row0[k] + n0;
        for (int kb = (t == first_tile) ? pro_kb : 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (p.debug & 4096) {                 // experiment: no operand traffic after the prologue (stale tiles)
            mbar_arrive(&full_bar[stage]);
          } else {
            mbar_expect_tx(&full_bar[stage], L::kStageBytes);
            uint8_t* sa = smem + stage * L::kStageBytes;
            tma_load_2d(sa, &tmap_a, kb * BK, ya, &full_bar[stage]);
            tma_load_2d(sa + L::kABytes, &tmap_b, kb * BK, yb, &full_bar[stage]);
          }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: the whole warp walks the loop, one elected lane issues (see tcgen05_mma_f16) =====
    {
      // instruction descriptor: D fp32 (1<<4), A/B fp16 K-major, N>>3 at [17,23), M>>4 at [24,29)
      const uint32_t idesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      const uint64_t adesc0 = make_kmajor_sw128_desc(smem_u32(smem));
      const uint64_t bdesc0 = make_kmajor_sw128_desc(smem_u32(smem) + L::kABytes);
      uint32_t g = 0;                 // running k-block count: stage g % kStages, barrier parity (g / kStages) & 1
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int t = first_tile; t < num_tiles; t += tile_step) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < num_kb; ++kb, ++g) {
          const uint32_t stage = g % kStages;
          mbar_wait(&full_bar[stage], (g / kStages) & 1u);
          tcgen05_fence_after();
          const uint64_t adesc = adesc0 + (uint64_t)(stage * (uint32_t)(L::kStageBytes >> 4));
          const uint64_t bdesc = bdesc0 + (uint64_t)(stage * (uint32_t)(L::kStageBytes >> 4));
#pragma unroll
          for (int kk = 0; kk < BK / UMMA_K; ++kk) {
            // advance 16 fp16 = 32 B along K inside the swizzle atom: +2 in the (addr>>4) field
            tcgen05_mma_f16(tmem_d, adesc + (uint64_t)(kk * 2), bdesc + (uint64_t)(kk * 2), idesc,
                            (uint32_t)((kb | kk) != 0));
          }
          tcgen05_commit(&empty_bar[stage]);   // frees the smem slot when these MMAs retire
        }
        tcgen05_commit(&tmem_full[acc]);      // accumulator complete -> epilogue
        if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===== epilogue: 8 warps; warp w may only touch TMEM lanes [32*(w%4), +32); the two warps of a
    // lane quarter split the tile's columns in halves of BN/2 =====
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;
    const int et = threadIdx.x - 64;  // 0..kEpiThreads-1
    int acc = 0;
    uint32_t acc_phase = 0;
    asm volatile("griddepcontrol.wait;" ::: "memory");   // boxes / kinds / candidate lists belong to earlier kernels
    const long long t_go = clock64();
    unsigned long long g_go = 0;
    if (p.debug & 1024) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g_go));
    for (int t = first_tile; t < num_tiles; t += tile_step) {
      int k, m0, n0;
      tile_of(sF, t, BN, k, m0, n0);
      const int pn = sF->n[k], pm = sF->m[k];
      const int sid = sF->cand_sid[k];
      const size_t R0 = (size_t)sF->row0[k], C0 = (size_t)sF->col0[k];
      const int row = m0 + quarter * 32 + lane;
      int rkind = BT_ROW_NONE;
      double rbox[4] = {0.0, 0.0, 0.0, 0.0};
      uint32_t r2h = 0, r1p = 0;   // packed conservative integer row box (see pack16 below)
      int rix1 = 0, riy1 = 0, rix2 = 0, riy2 = 0;
      float r_area_lb = 0.f;
      float rnorm = 1.0f;
      long long t_s1 = 0, t_s2 = 0, t_s3 = 0, t_rl = 0;
      if (!kDense) {
        // All global loads of the tile first, in one batch and without dependent chains: while the TMA
        // pipeline is streaming, an ordinary load queues behind ~4 stages of operand traffic (measured
        // 3-6 k cycles per round trip), so the row kind -> row box dependency of the obvious code and the
        // separate detection-box round trip cost ~14 k cycles (profiles/README.md).
        const int c = et, col = n0 + et;          // BN <= kEpiThreads: one detection per thread
        double2 rlo = make_double2(0.0, 0.0), rhi = rlo, clo = rlo, chi = rlo;
        float4 r32 = make_float4(0.f, 0.f, 0.f, 0.f);
        uint2 cpk = make_uint2(0u, 0u);
        uint8_t ck = BT_COL_NONE;
        float cn = 1.0f;
        if (row < pn) {
          const double2* s = reinterpret_cast<const double2*>(p.row_tlbr + (R0 + row) * 4);
          rkind = reinterpret_cast<const uint8_t*>(p.row_kind_base + sF->kind_off[k])[row];
          rlo = s[0]; rhi = s[1];
          if (p.row_tlbr_f32) r32 = *reinterpret_cast<const float4*>(p.row_tlbr_f32 + (R0 + row) * 4);
          if (p.row_norm) rnorm = p.row_norm[R0 + row];
        }
        if (c < BN && col < pm) {
          const double2* s = reinterpret_cast<const double2*>(p.col_tlbr + (C0 + col) * 4);
          ck = p.col_kind[C0 + col];
          clo = s[0]; chi = s[1];
          if (p.col_pk) cpk = p.col_pk[C0 + col];
          if (p.col_norm) cn = p.col_norm[C0 + col];
        }
        // stage the detection boxes: packed 15-bit integer corners for the screen, float64 for the exact
        // path, and the score class
        asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");  // previous tile's readers are done
        t_s1 = clock64();
        if (c < BN) {
          uint2 pk = make_uint2(0x7fff7fffu, 0x80008000u);   // x1+1 = y1+1 = 32767, x2 = y2 = 0: overlaps nothing
          if (ck != BT_COL_NONE) pk = p.col_pk ? cpk : pack16_box(clo.x, clo.y, chi.x, chi.y, true);
          s_colpk[c] = pk;
          s_colkind[c] = ck;
          s_colinv[c] = cn > 0.f ? 1.0f / cn : 0.f;
          reinterpret_cast<double2*>(s_col64 + c * 4)[0] = clo;
          reinterpret_cast<double2*>(s_col64 + c * 4)[1] = chi;
        }
        t_s2 = clock64();
        asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
        t_s3 = clock64();
        if (p.debug & 16) {
          asm volatile("" ::"d"(rlo.x), "d"(rhi.y), "r"(rkind) : "memory");   // the row loads have landed
          t_rl = clock64();
        }
        if (rkind != BT_ROW_NONE) {
          rbox[0] = rlo.x; rbox[1] = rlo.y; rbox[2] = rhi.x; rbox[3] = rhi.y;
          const uint2 pk = p.row_tlbr_f32 ? bt_pack16_f32(r32.x, r32.y, r32.z, r32.w, false)
                                          : pack16_box(rlo.x, rlo.y, rhi.x, rhi.y, false);
          r1p = pk.x;   // (x1+1) | (y1+1) << 16
          r2h = pk.y;   // x2 | y2 << 16 | 0x80008000
          rix1 = (int)(pk.x & 0xffffu) - 1; riy1 = (int)(pk.x >> 16) - 1;
          rix2 = (int)(pk.y & 0x7fffu); riy2 = (int)((pk.y >> 16) & 0x7fffu);
          // the integer corners are rounded outward by < 1 px per side: a lower bound of the true area
          r_area_lb = (float)max(rix2 - rix1 - 2, 0) * (float)max(riy2 - riy1 - 2, 0);
        }
      }
      const long long t_rows = clock64();
      const int seg = n0 / (BN / 2) + half;   // the warp's rows x this column half = one candidate segment per row
      double* my_preiou = s_preiou + (warp - 2) * (kPreSlots * 32);     // [slot][lane]
      uint32_t* my_prekey = s_prekey + (warp - 2) * (kPreSlots * 32);
      // This lane owns (row, seg) of every candidate list: counts live in registers, appends need no atomics.
      int cnt_a = 0, cnt_b = 0;   // edges appended so far: list 0 (or 2 for an unconfirmed row) | list 1
      my_prekey[lane] = 0u; my_prekey[32 + lane] = 0u;   // slots 0 / 1 are read unconditionally by the exact pass
      int npre = 0;               // box-pass candidates waiting in shared memory (<= kPreSlots)
      int last_a = 0, last_b = 0; // column (| flags) of the latest edge per list (the row's only column if its degree stays 1)
      int spill_a = -1, spill_b = -1;   // 0 once the box pass wrote edges of a crowded row straight to the list
      int dbg_open = 0;
      const bool unc = rkind == BT_ROW_UNCONFIRMED;
      const int la = unc ? 2 : 0;
      const double thr_a = unc ? p.unconf_thresh : p.match_thresh;
      // the raw accumulator of this row opens the appearance gate (or comes within the tensor-core error of
      // it) at acc >= gate_lo: sim = acc / ||track feature|| (the bank holds raw rows)
      const float gate_lo = (rnorm > 0.f) ? (p.sim_gate - p.gate_band) * rnorm : 3.0e38f;
      // candidate list a (row, detection class) pair belongs to: demo:1539-1556 (0), demo:1569-1571 (1), demo:1593-1604 (2)
      auto list_of = [&](const int ckind) -> int {
        if (ckind == BT_COL_HIGH) return la;
        if (ckind == BT_COL_LOW && rkind == BT_ROW_POOL_TRACKED) return 1;
        return -1;
      };
      constexpr int kHalfCols = BN / 2, kFullChunks = kHalfCols / 32, kTailCols = kHalfCols % 32;
      constexpr int kChunks = kFullChunks + (kTailCols ? 1 : 0);
      static_assert(kTailCols == 0 || kTailCols == 16, "column half must be a multiple of 16");
      static_assert(BN <= kEpiThreads, "one staged detection per epilogue thread");
      // ---- Box pass, in the shadow of the MMA main loop (it needs no accumulator).  With the
      // appearance gate closed the cost of a pair is its IoU distance in every stage, so every edge that
      // exists without the similarity is found, evaluated in float64 and appended here:
      //  1. integer screen on outward-rounded 15-bit corners, two corners per 32-bit word:
      //       r.x2 > c.x1 & r.y2 > c.y1  <=>  both half-words of (r2 | 0x80008000) - (c1 + 0x00010001) keep bit 15
      //       c.x2 > r.x1 & c.y2 > r.y1  <=>  likewise with the roles swapped     (false => the exact IoU is 0)
      //  2. for the few overlapping pairs an integer upper bound of the intersection against lower bounds
      //     of the areas (no division): the IoU must be able to reach 1 - max(stage threshold)
      //  3. exact float64 IoU distance against the stage threshold.
      // (With real face similarities the gate is not a function of the body similarity alone: everything
      // is left to the similarity pass.)
      const bool all_open = sF->face_sim[k] != nullptr || (p.debug & 32);   // no box pass: every pair goes through open_pair
      if (!kDense && !all_open && rkind != BT_ROW_NONE) {
#pragma unroll
        for (int ch = 0; ch < kChunks; ++ch) {
          const int W = (ch < kFullChunks) ? 32 : kTailCols;
          const int c0 = half * kHalfCols + ch * 32;
          const uint32_t colbase = smem_u32(s_colpk) + (uint32_t)c0 * 8u;
          uint32_t hot_ov = 0;
#pragma unroll
          for (int c = 0; c < W; c += 2) {
            // two detections per 16-byte load; volatile + memory clobber: must stay below the staging
            // barrier (bar.sync 1) -- a plain asm was hoisted above it by the compiler
            uint32_t a1, a2, b1, b2;
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a1), "=r"(a2), "=r"(b1), "=r"(b2)
                         : "r"(colbase + (uint32_t)c * 8u) : "memory");
            if (((r2h - a1) & (a2 - r1p) & 0x80008000u) == 0x80008000u) hot_ov |= 1u << c;
            if (((r2h - b1) & (b2 - r1p) & 0x80008000u) == 0x80008000u) hot_ov |= 2u << c;
          }
          while (hot_ov) {
            const int lc = c0 + __ffs(hot_ov) - 1;
            hot_ov &= hot_ov - 1;
            const uint2 pk = s_colpk[lc];
            const int cx1 = (int)(pk.x & 0xffffu) - 1, cy1 = (int)(pk.x >> 16) - 1;
            const int cx2 = (int)(pk.y & 0x7fffu), cy2 = (int)((pk.y >> 16) & 0x7fffu);
            const float inter_ub = (float)(min(rix2, cx2) - max(rix1, cx1)) * (float)(min(riy2, cy2) - max(riy1, cy1));
            const float c_area_lb = (float)max(cx2 - cx1 - 2, 0) * (float)max(cy2 - cy1 - 2, 0);
            // IoU <= inter_ub / (areas_lb - inter_ub) < gate  <=>  inter_ub * (1 + gate) < gate * areas_lb
            if (inter_ub * (1.0f + p.iou_gate) * 1.0001f < p.iou_gate * (r_area_lb + c_area_lb)) continue;
            const int list = list_of(s_colkind[lc]);
            if (list < 0) continue;
            if (npre < kPreSlots) {
              // remembered in shared memory; the float64 IoU waits until the main loop is over (FP64
              // instructions crawl next to a running tcgen05 pipeline: ~5 k cycles per evaluation here,
              // ~300 afterwards, and they slow the MMAs down -- profiles/README.md)
              my_prekey[npre * 32 + lane] = (uint32_t)lc | ((uint32_t)(list == 1) << 16);
              ++npre;
              continue;
            }
            // crowded row (more than kPreSlots box candidates in this half tile): evaluate now, straight to the list
            const double iou_d = iou_dist_f64(rbox, s_col64 + lc * 4);
            if (!(iou_d < (list == 1 ? p.second_thresh : thr_a))) continue;
            const int kk = (list == 1) ? cnt_b++ : cnt_a++;
            if (list == 1) { last_b = n0 + lc; spill_b = 0; } else { last_a = n0 + lc; spill_a = 0; }
            emit_owned_tc(p, sid, list, row, seg, kk, n0 + lc, 0, iou_d);
          }
        }
      }
      __syncwarp();
      const long long t_shadow = clock64();
      mbar_wait(&tmem_full[acc], acc_phase);
      const long long t_acc = clock64();
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BN);
      uint32_t va[32], vb[32];
      auto issue = [&](const int ch, uint32_t (&v)[32]) {
        const uint32_t c0 = (uint32_t)(half * kHalfCols + ch * 32);
        if (ch < kFullChunks) tmem_ld_32x32b_x32(taddr + c0, v);
        else tmem_ld_32x32b_x16(taddr + c0, v);
      };
      // ---- Similarity pass.  A pair whose appearance gate is open (sim >= sim_gate, == !((1.0f - sim) >
      // appearance), sim_gate_for()) or within the tensor-core error of it gets the fused cost: an edge of
      // the box pass is overwritten in place, otherwise the pair is evaluated from scratch.  Lane-local, no
      // shuffles: the lane owns its row.
      auto open_pair = [&](const int lc, const float accv) {
        const int list = list_of(s_colkind[lc]);
        if (list < 0) return;
        float sim = accv;
        if (p.row_norm) sim = (rnorm > 0.f) ? accv / rnorm : 0.f;       // curr feature = raw row / norm (demo:497-502)
        if (unc && p.col_norm) sim *= s_colinv[lc];                     // stage 3 compares normalised detections (demo:1593-1599)
        const uint32_t want = (uint32_t)lc | ((uint32_t)(list == 1) << 16);
        bool found = false;
#pragma unroll
        for (int j = 0; j < kPreSlots; ++j) {
          if (j < npre && (my_prekey[j * 32 + lane] & 0x1ffffu) == want) {
            // an edge of the box pass whose gate turned out open (slots exist only without a face term)
            const double iou_d = my_preiou[j * 32 + lane];
            if (list == 0) my_preiou[j * 32 + lane] = fuse_stage1(iou_d, sim, 0.0f, p.appearance);
            if (list == 2) my_preiou[j * 32 + lane] = fuse_stage3(iou_d, sim, p.appearance, p.proximity);
            const int fl = sim_flags(p, list, sim, 0.0f);
            my_prekey[j * 32 + lane] = want | ((fl & BT_EDGE_SIM) ? (1u << 17) : 0u) | ((fl & BT_EDGE_AMBIG) ? (1u << 18) : 0u);
            found = true;
          }
        }
        if (found) return;
        const int cnt = (list == 1) ? cnt_b : cnt_a;
        const int added = open_pair_slow(p, k, rbox[0], rbox[1], rbox[2], rbox[3], s_col64 + lc * 4, list, row, seg, n0 + lc, sim,
                                         cnt, (list == 1) ? spill_b : spill_a);
        if (added >= 0) { if (list == 1) { ++cnt_b; last_b = added; } else { ++cnt_a; last_a = added; } }
      };
      long long t_l0 = 0, t_l1 = 0, t_pa = 0, t_ex = 0;
      if (kDense) {
#pragma unroll 1
        for (int ch = 0; ch < kChunks; ++ch) {
          issue(ch, va);
          tmem_ld_wait_dep(va);
          const int W = (ch < kFullChunks) ? 32 : kTailCols;
          if (row < pn) {
#pragma unroll
            for (int c = 0; c < 32; ++c) {
              const int col = n0 + half * kHalfCols + ch * 32 + c;
              if (c < W && col < pm) assoc_dense(p, row, col, __uint_as_float(va[c]));
            }
          }
        }
      } else {
        // phase A: chunk maxima, TMEM loads software-pipelined, no branches.  Warps that hold an unconfirmed
        // row while detection norms are in play (rare) scale every element by 1 / ||detection feature||.
        uint32_t gatebits = 0;
        const bool scaled = __any_sync(0xffffffffu, unc && p.col_norm != nullptr);
        if (!scaled) {
          issue(0, va);
#pragma unroll
          for (int ch = 0; ch < kChunks; ++ch) {
            uint32_t (&cur)[32] = (ch & 1) ? vb : va;
            uint32_t (&nxt)[32] = (ch & 1) ? va : vb;
            tmem_ld_wait_dep(cur);
            if (p.debug & 16) { if (ch == 0) t_l0 = clock64(); if (ch == 1) t_l1 = clock64(); }
            if (ch + 1 < kChunks) issue(ch + 1, nxt);
            const int G = ((ch < kFullChunks) ? 32 : kTailCols) / 4;
            float m4[8];
#pragma unroll
            for (int g = 0; g < G; ++g)
              m4[g] = fmaxf(fmaxf(__uint_as_float(cur[4 * g]), __uint_as_float(cur[4 * g + 1])),
                            fmaxf(__uint_as_float(cur[4 * g + 2]), __uint_as_float(cur[4 * g + 3])));
            float mx = m4[0];
            if (G == 8) mx = fmaxf(fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])), fmaxf(fmaxf(m4[4], m4[5]), fmaxf(m4[6], m4[7])));
            else mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
            if (mx >= gate_lo) gatebits |= 1u << ch;
          }
        } else {
#pragma unroll 1
          for (int ch = 0; ch < kChunks; ++ch) {
            tmem_ld_32x32b_x32(taddr + (uint32_t)(half * kHalfCols + ch * 32), va);
            tmem_ld_wait_dep(va);
            const int W = (ch < kFullChunks) ? 32 : kTailCols;
            float mx = -3.0e38f;
#pragma unroll
            for (int c = 0; c < 32; ++c) {
              const float f = unc ? s_colinv[half * kHalfCols + ch * 32 + (c < W ? c : 0)] : 1.0f;
              if (c < W) mx = fmaxf(mx, __uint_as_float(va[c]) * f);
            }
            if (mx >= gate_lo) gatebits |= 1u << ch;
          }
        }
        if (all_open) gatebits = (1u << kChunks) - 1u;
        if (p.debug & 16) t_pa = clock64();
        // exact pass: float64 IoU distance of the remembered candidates, all lanes at once; slots 0 and 1
        // are evaluated together (two independent dependency chains), slots 2.. only if some row has them
        {
          const uint32_t k0 = my_prekey[lane], k1 = my_prekey[32 + lane];
          const double d0 = iou_dist_f64(rbox, s_col64 + (k0 & 0xffu) * 4);
          const double d1 = iou_dist_f64(rbox, s_col64 + (k1 & 0xffu) * 4);
          if (npre > 0) my_preiou[lane] = d0;
          if (npre > 1) my_preiou[32 + lane] = d1;
          if (__any_sync(0xffffffffu, npre > 2)) {
#pragma unroll
            for (int j = 2; j < kPreSlots; ++j)
              if (j < npre) my_preiou[j * 32 + lane] = iou_dist_f64(rbox, s_col64 + (my_prekey[j * 32 + lane] & 0xffu) * 4);
          }
        }
        if (p.debug & 16) t_ex = clock64();
        if ((p.debug & 2) || rkind == BT_ROW_NONE) gatebits = 0;
        // phase B: only chunks in which some row of the warp has an open gate are looked at again (one rolled
        // loop; the tail chunk is loaded 32 wide and masked).  The chunk loop only COLLECTS the open pairs --
        // (local column, accumulator), two per lane in registers -- and the float64 fusion runs afterwards, once
        // for all lanes: rows of a warp have their open pairs in different chunks, and evaluating them chunk by
        // chunk ran the expensive part once per chunk with one or two lanes active.
        uint32_t need = __reduce_or_sync(0xffffffffu, gatebits);
        int e0c = -1, e1c = -1;
        float e0v = 0.f, e1v = 0.f;
        while (need) {                          // warp-uniform
          const int ch = __ffs(need) - 1;
          need &= need - 1;
          tmem_ld_32x32b_x32(taddr + (uint32_t)(half * kHalfCols + ch * 32), va);
          tmem_ld_wait_dep(va);
          if ((gatebits >> ch) & 1u) {
            uint32_t hot = 0;
            if (scaled && unc) {
#pragma unroll
              for (int c = 0; c < 32; ++c)
                if (__uint_as_float(va[c]) * s_colinv[half * kHalfCols + ch * 32 + ((ch < kFullChunks || c < kTailCols) ? c : 0)] >= gate_lo) hot |= 1u << c;
            } else {
#pragma unroll
              for (int c = 0; c < 32; ++c)
                if (__uint_as_float(va[c]) >= gate_lo) hot |= 1u << c;
            }
            if (all_open) hot = 0xffffffffu;
            if (ch >= kFullChunks) hot &= (1u << kTailCols) - 1u;
            while (hot) {
              const int c = __ffs(hot) - 1;
              hot &= hot - 1;
              // va[c] by a 5-level select tree on the bits of c
              uint32_t s16[16], s8[8], s4[4], s2[2];
#pragma unroll
              for (int i = 0; i < 16; ++i) s16[i] = (c & 16) ? va[i + 16] : va[i];
#pragma unroll
              for (int i = 0; i < 8; ++i) s8[i] = (c & 8) ? s16[i + 8] : s16[i];
#pragma unroll
              for (int i = 0; i < 4; ++i) s4[i] = (c & 4) ? s8[i + 4] : s8[i];
#pragma unroll
              for (int i = 0; i < 2; ++i) s2[i] = (c & 2) ? s4[i + 2] : s4[i];
              const float accv = __uint_as_float((c & 1) ? s2[1] : s2[0]);
              const int lc = half * kHalfCols + ch * 32 + c;
              if (e0c < 0) { e0c = lc; e0v = accv; }
              else if (e1c < 0) { e1c = lc; e1v = accv; }
              else open_pair(lc, accv);          // a third open pair of one row in this half tile: on the spot
              ++dbg_open;
            }
          }
          __syncwarp();
        }
#pragma unroll 1
        for (int e = 0; e < 2; ++e) {
          const int lc = e ? e1c : e0c;
          const float v = e ? e1v : e0v;
          if (!__any_sync(0xffffffffu, lc >= 0)) break;
          if (lc >= 0) open_pair(lc, v);
          __syncwarp();
        }
      }
      const long long t_chunks = clock64();
      if (!kDense && rkind != BT_ROW_NONE) {
        // the remembered edges go out with their final cost, then the (row, segment) bookkeeping the LAP
        // kernel classifies rows with: segment count, row degree, the row's column if it has only one
#pragma unroll
        for (int j = 0; j < kPreSlots; ++j) {
          if (j < npre) {
            const uint32_t key = my_prekey[j * 32 + lane];
            const double cost = my_preiou[j * 32 + lane];
            const int col = n0 + (int)(key & 0xffu);
            const int fl = ((key >> 17) & 1u ? BT_EDGE_SIM : 0) | ((key >> 18) & 1u ? BT_EDGE_AMBIG : 0);
            if ((key >> 16) & 1u) {
              if (cost < p.second_thresh) { emit_owned_tc(p, sid, 1, row, seg, cnt_b++, col, 0, cost); last_b = col; }
            } else if (cost < thr_a || (fl & BT_EDGE_AMBIG)) {
              emit_owned_tc(p, sid, la, row, seg, cnt_a++, col, fl, cost);
              last_a = col | fl;
            }
          }
        }
        if (cnt_a) {
          const size_t ri = (size_t)la * p.cand.rows_cap + row;
          c_cnt(p, sid)[ri * p.cand.nseg + seg] = cnt_a;
          atomicAdd(&c_total(p, sid)[la], cnt_a);
          atomicOr(&c_segmask(p, sid)[ri], 1ull << seg);
          atomicAdd(&c_rowdeg(p, sid)[ri], cnt_a);
          if (cnt_a == 1) c_rowcol(p, sid)[ri] = last_a;
        }
        if (cnt_b) {
          const size_t ri = (size_t)1 * p.cand.rows_cap + row;
          c_cnt(p, sid)[ri * p.cand.nseg + seg] = cnt_b;
          atomicAdd(&c_total(p, sid)[1], cnt_b);
          atomicOr(&c_segmask(p, sid)[ri], 1ull << seg);
          atomicAdd(&c_rowdeg(p, sid)[ri], cnt_b);
          if (cnt_b == 1) c_rowcol(p, sid)[ri] = last_b;
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if ((p.debug & 16) && lane == 0 && (blockIdx.x % 37) == 0)
        printf("EPI cta %d warp %d: go %lld bar1 +%lld staged +%lld bar2 +%lld rowloads +%lld rows +%lld box pass done +%lld acc +%lld ld0 +%lld ld1 +%lld phaseA +%lld exact +%lld sim pass +%lld tail +%lld (lane 0: box edges %d, open pairs %d)\n", blockIdx.x,
               warp, t_go - t_start, t_s1 - t_go, t_s2 - t_go, t_s3 - t_go, t_rl - t_go, t_rows - t_go, t_shadow - t_go, t_acc - t_go, t_l0 - t_acc, t_l1 - t_l0, t_pa - t_acc, t_ex - t_acc, t_chunks - t_acc, clock64() - t_chunks, npre, dbg_open);
      if ((p.debug & 1024) && lane == 0) {
        unsigned long long g_end;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g_end));
        printf("GT %d %d %llu %llu %lld\n", blockIdx.x, warp, g_go, g_end, clock64() - t_go);
      }
      if ((p.debug & 4) && et == 0)
        printf("ALL %d %lld %lld\n", blockIdx.x, t_acc - t_start, clock64() - t_start);
      if ((p.debug & 1) && et == 0 && (blockIdx.x % 37) == 0)
        printf("cta %d (BN=%d grid=%d): accumulator ready at %lld cycles, epilogue done at %lld\n", blockIdx.x, BN,
               gridDim.x, t_acc - t_start, clock64() - t_start);
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
    }
  }

  tcgen05_fence_before();
  __syncwarp();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
  if ((p.debug & 32768) && threadIdx.x == 0 && stamp_sm >= 0) {
    unsigned long long g_end;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g_end));
    g_assoc_stamps[stamp_sm][stamp_i][1] = g_end;
  }
  if ((p.debug & 1) && threadIdx.x == 0 && (blockIdx.x % 37) == 0) {
    unsigned long long g_end;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g_end));
    printf("cta %d: globaltimer start %llu end %llu (ns), lifetime %llu ns = %lld cycles\n", blockIdx.x, g_start, g_end,
           g_end - g_start, clock64() - t_start);
  }
}

// ------------------------------------------------------------------------------------------------
// CUDA-core kernel (IoU-only association, small feature sizes, feature sizes that are not a multiple
// of 64, and the exact cross-check of the tensor path in the tests): fp32 accumulation of fp32 or fp16
// operands -- with fp16 ingest the operands ARE the reference's values, so this path is exact.
// ------------------------------------------------------------------------------------------------
constexpr int ST = 64, SK = 16;

__device__ __forceinline__ float ld_feat(const float* p, size_t i) { return p[i]; }
__device__ __forceinline__ float ld_feat(const __half* p, size_t i) { return __half2float(p[i]); }

template <bool kDense, typename T>
__global__ void __launch_bounds__(256)
assoc_simt_kernel(const T* __restrict__ a, const T* __restrict__ b, const __grid_constant__ EpiParams p, int d) {
  __shared__ float sa[SK][ST + 1];
  __shared__ float sb[SK][ST + 1];
  const int k = blockIdx.z;
  const int pn = p.F->n[k], pm = p.F->m[k];
  const int row0 = blockIdx.y * ST, col0 = blockIdx.x * ST;
  if (row0 >= pn || col0 >= pm) return;
  const size_t ar0 = (size_t)p.F->a_row0[k], br0 = (size_t)p.F->b_row0[k];
  const size_t R0 = (size_t)p.F->row0[k], C0 = (size_t)p.F->col0[k];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < d; k0 += SK) {
    for (int e = threadIdx.x; e < ST * SK; e += 256) {
      const int r = e / SK, kk = e % SK;
      sa[kk][r] = (row0 + r < pn && k0 + kk < d) ? ld_feat(a, (ar0 + row0 + r) * d + k0 + kk) : 0.f;
      sb[kk][r] = (col0 + r < pm && k0 + kk < d) ? ld_feat(b, (br0 + col0 + r) * d + k0 + kk) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < SK; ++kk) {
      float av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { av[i] = sa[kk][ty * 4 + i]; bv[i] = sb[kk][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  // candidate mode: the same integer overlap screen as the tensor-core kernel decides which pairs can
  // have a cost below 1 at all (boxes overlap, or the appearance gate is open); only those reach the
  // float64 path
  uint2 cpk[4];
  int ckind[4];
  float cinv[4];
  if (!kDense) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = col0 + tx * 4 + j;
      ckind[j] = BT_COL_NONE;
      cpk[j] = make_uint2(0x7fff7fffu, 0x80008000u);
      cinv[j] = 1.0f;
      if (col < pm) {
        ckind[j] = p.col_kind[C0 + col];
        if (p.col_pk) cpk[j] = p.col_pk[C0 + col];
        else { const double* c = p.col_tlbr + (C0 + col) * 4; cpk[j] = pack16_box(c[0], c[1], c[2], c[3], true); }
        if (p.col_norm) { const float cn = p.col_norm[C0 + col]; cinv[j] = cn > 0.f ? 1.0f / cn : 0.f; }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = row0 + ty * 4 + i;
    if (row >= pn) continue;
    const int rkind = (!kDense) ? reinterpret_cast<const uint8_t*>(p.row_kind_base + p.F->kind_off[k])[row] : 0;
    uint2 rpk = make_uint2(0u, 0u);
    float rnorm = 1.0f;
    if (!kDense && rkind != BT_ROW_NONE) {
      if (p.row_tlbr_f32) {
        const float4 r = *reinterpret_cast<const float4*>(p.row_tlbr_f32 + (R0 + row) * 4);
        rpk = bt_pack16_f32(r.x, r.y, r.z, r.w, false);
      } else {
        const double* r = p.row_tlbr + (R0 + row) * 4;
        rpk = pack16_box(r[0], r[1], r[2], r[3], false);
      }
      if (p.row_norm) rnorm = p.row_norm[R0 + row];
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = col0 + tx * 4 + j;
      if (col >= pm) continue;
      if (kDense) {
        assoc_dense(p, row, col, acc[i][j]);
      } else {
        if (rkind == BT_ROW_NONE || ckind[j] == BT_COL_NONE) continue;
        float sim = acc[i][j];
        if (p.row_norm) sim = rnorm > 0.f ? sim / rnorm : 0.f;
        if (rkind == BT_ROW_UNCONFIRMED) sim *= cinv[j];
        const bool overlap = ((rpk.y - cpk[j].x) & (cpk[j].y - rpk.x) & 0x80008000u) == 0x80008000u;
        if (overlap || sim >= p.sim_gate || p.F->face_sim[k] != nullptr) assoc_exact(p, k, row, col, sim, rkind, ckind[j]);
      }
    }
  }
}

// Smallest float s with !((1.0f - s) > appearance): `1.0f - s` is monotone in s, so the reference's
// test `emb_dists > appearance_thresh` (demo:1545) on emb = 1 - sim is exactly `sim < s`.
static float sim_gate_for(float appearance) {
  float s = 1.0f - appearance;
  while ((1.0f - s) > appearance) s = nextafterf(s, 2.0f);
  while (!((1.0f - nextafterf(s, -2.0f)) > appearance)) s = nextafterf(s, -2.0f);
  return s;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);

}  // namespace

struct bt_gemm_ws {
  PFN_encodeTiled encode = nullptr;
  struct Entry { const void* base; int rows, d, box_rows; CUtensorMap map; };
  Entry cache[8] = {};
  int next = 0;
};

int32_t bt_gemm_ws_create(bt_ctx* ctx) {
  ctx->gemm = new bt_gemm_ws();
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  BT_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  BT_CHECK(fn != nullptr && qres == cudaDriverEntryPointSuccess, BT_ERR_CUDA,
           "cuTensorMapEncodeTiled not available from the driver");
  ctx->gemm->encode = reinterpret_cast<PFN_encodeTiled>(fn);
  return BT_OK;
}

void bt_gemm_ws_destroy(bt_ctx* ctx) {
  delete ctx->gemm;
  ctx->gemm = nullptr;
}

static int32_t make_tmap(bt_ctx* ctx, CUtensorMap* tm, const __half* base, int rows, int d, int box_rows) {
  bt_gemm_ws* ws = ctx->gemm;      // per ctx (a ctx belongs to one host thread): no shared state between threads
  for (const auto& e : ws->cache)
    if (e.base == base && e.rows == rows && e.d == d && e.box_rows == box_rows) { *tm = e.map; return BT_OK; }
  const cuuint64_t gdim[2] = {(cuuint64_t)d, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)d * sizeof(__half)};
  const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  const cuuint32_t estride[2] = {1, 1};
  CUresult r = ctx->gemm->encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(base), gdim,
                                 gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  BT_CHECK(r == CUDA_SUCCESS, BT_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  bt_gemm_ws::Entry& e = ws->cache[ws->next];
  ws->next = (ws->next + 1) % 8;
  e.base = base; e.rows = rows; e.d = d; e.box_rows = box_rows; e.map = *tm;
  return BT_OK;
}

template <int BN, bool kDense>
static int32_t launch_tc(bt_ctx* ctx, const bt_assoc_params& ap, const EpiParams& ep, int tiles) {
  CUtensorMap ta, tb;
  BT_TRY(make_tmap(ctx, &ta, ap.a16, ap.a_rows_alloc, ap.d, BM));
  BT_TRY(make_tmap(ctx, &tb, ap.b16, ap.b_rows_alloc, ap.d, BN));
  auto kern = assoc_tc_kernel<BN, kDense>;
  static std::once_flag attr_once[64];  // per instantiation and device ordinal; safe with several host threads
  cudaError_t attr_err = cudaSuccess;
  std::call_once(attr_once[ctx->device & 63], [&] {
    attr_err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TcSmem<BN>::kDyn);
  });
  BT_CUDA(attr_err);
  if (tiles <= 0) return BT_OK;
  // may start its ramp (and, with operands_early, its main loop) under the previous kernel (griddepcontrol.wait inside)
  BT_CUDA(bt_launch(ctx, true, kern, dim3(tiles < ctx->num_sms ? tiles : ctx->num_sms), dim3(kTcThreads), TcSmem<BN>::kDyn,
                    ta, tb, ep, ap.d));
  BT_LAUNCHED(ctx);
  return BT_OK;
}

int32_t btk_assoc_pick_bn(const bt_ctx* ctx, const int32_t* n, const int32_t* m, int32_t count) {
  int mx_m = 0;
  for (int k = 0; k < count; ++k) mx_m = m[k] > mx_m ? m[k] : mx_m;
  const char* e = getenv("BT_ASSOC_BN");
  if (e) {     // profiling override; a narrow tile that would not fit the 64 segments per row is ignored
    const int v = atoi(e);
    if ((v == 224 || v == 128 || v == 64) && (mx_m + v / 2 - 1) / (v / 2) <= BT_CAND_MAXSEG) return v;
    return 256;
  }
  // A tile's main loop costs the tensor pipe max(BN, 116) / 2 cycles per MMA (tools/mma_issue_probe.cu: N <= 112 sits on
  // the 58-cycle issue floor), so the launch costs waves x that.  Narrow tiles only pay for small problems -- a
  // 128 x 256 tile spends 8 us of MMA issue on a 64 x 64 frame -- and only there do they fit the candidate lists'
  // 64 segments per row (segment = half a tile).
  long best_cost = -1;
  int best = 256;
  for (int bn : {256, 224, 128, 64}) {
    if (bn < 224 && (mx_m + bn / 2 - 1) / (bn / 2) > BT_CAND_MAXSEG) continue;
    long tiles = 0;
    for (int k = 0; k < count; ++k) tiles += (long)((n[k] + BM - 1) / BM) * ((m[k] + bn - 1) / bn);
    const long waves = (tiles + ctx->num_sms - 1) / ctx->num_sms;
    const long cost = waves * (bn > 116 ? bn : 116);
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = bn; }
  }
  return best;
}

void btk_assoc_fill(const bt_assoc_params& ap, int32_t precision, bt_assoc_frame* f) {
  memset(f, 0, sizeof(*f));
  f->count = ap.count;
  const int bn = ap.bn ? ap.bn : 256;
  int tiles = 0;
  for (int k = 0; k < ap.count; ++k) {
    // an empty problem contributes no tiles: zero both sizes so that the tile arithmetic sees it
    const bool empty = ap.n[k] <= 0 || ap.m[k] <= 0;
    f->n[k] = empty ? 0 : ap.n[k]; f->m[k] = empty ? 0 : ap.m[k];
    f->a_row0[k] = ap.a_row0[k]; f->b_row0[k] = ap.b_row0[k];
    f->row0[k] = ap.row0[k]; f->col0[k] = ap.col0[k];
    f->kind_off[k] = ap.kind_off[k]; f->pos_off[k] = ap.pos_off[k];
    f->cand_sid[k] = ap.cand_sid[k];
    f->face_sim[k] = ap.face_sim[k];
    f->tile_start[k] = tiles;
    if (precision == 0) tiles += ((f->n[k] + BM - 1) / BM) * ((f->m[k] + bn - 1) / bn);
  }
  for (int k = ap.count; k <= BT_MAX_BATCH; ++k) f->tile_start[k] = tiles;
}

int32_t btk_assoc_launch(bt_ctx* ctx, const bt_assoc_params& ap, int32_t precision, const bt_assoc_frame& hf,
                         const bt_assoc_frame* df, int fixed, int max_rows, int max_cols) {
  BT_CHECK(ap.count >= 1 && ap.count <= BT_MAX_BATCH, BT_ERR_INVALID, "bad batch size %d", ap.count);
  int any = 0, mx_n = 0, mx_m = 0;
  for (int k = 0; k < ap.count; ++k) {
    if (hf.n[k] > 0 && hf.m[k] > 0) any = 1;
    mx_n = hf.n[k] > mx_n ? hf.n[k] : mx_n;
    mx_m = hf.m[k] > mx_m ? hf.m[k] : mx_m;
  }
  if (!any && !fixed) return BT_OK;
  if (fixed) { mx_n = max_rows; mx_m = max_cols; }
  EpiParams ep;
  memset(&ep, 0, sizeof(ep));
  ep.F = df;
  ep.row_tlbr = ap.row_tlbr; ep.row_tlbr_f32 = ap.row_tlbr_f32; ep.row_norm = ap.row_norm;
  ep.row_kind_base = ap.row_kind_base;
  ep.col_tlbr = ap.col_tlbr; ep.col_kind = ap.col_kind; ep.col_pk = ap.col_pk; ep.col_norm = ap.col_norm;
  ep.match_thresh = ap.match_thresh; ep.second_thresh = ap.second_thresh;
  ep.unconf_thresh = ap.unconf_thresh; ep.proximity = ap.proximity; ep.appearance = ap.appearance;
  ep.cand = ap.cand; ep.out_emb = ap.out_emb; ep.out_dists = ap.out_dists;
  ep.dense_stage = ap.dense_stage;
  ep.operands_early = ap.operands_early;
  ep.debug = getenv("BT_ASSOC_DEBUG") ? atoi(getenv("BT_ASSOC_DEBUG")) : 0;
  ep.sim_gate = sim_gate_for(ap.appearance);
  ep.gate_band = ap.gate_band;
  {
    double mx = ap.match_thresh > ap.second_thresh ? ap.match_thresh : ap.second_thresh;
    if (ap.unconf_thresh > mx) mx = ap.unconf_thresh;
    const double g = (1.0 - mx) * 0.999 - 1e-6;      // safely below the smallest IoU any stage can accept
    ep.iou_gate = g > 0.0 ? (float)g : 0.0f;
  }
  const bool dense = (ap.out_emb != nullptr) || (ap.out_dists != nullptr);
  BT_CHECK(!dense || ap.count == 1, BT_ERR_INVALID, "dense dumps take one problem");
  if (precision == 0) {
    BT_CHECK(ap.d % BK == 0 && ap.d >= BK, BT_ERR_INVALID,
             "tensor-core similarity needs feat_dim %% 64 == 0 (got %d)", ap.d);
    BT_CHECK(ap.a16 && ap.b16, BT_ERR_INVALID, "fp16 operands missing");
    const int bn = ap.bn ? ap.bn : 256;
    BT_CHECK(dense || ap.cand.seg * 2 == bn, BT_ERR_STATE,
             "candidate segment size %d does not match the tile width %d", ap.cand.seg, bn);
    const int tiles = fixed ? ap.count * ((max_rows + BM - 1) / BM) * ((max_cols + bn - 1) / bn) : hf.tile_start[ap.count];
    if (!fixed && ap.count == 1 && hf.n[0] > 0 && hf.m[0] > 0) {     // (a captured graph keeps its arguments: no hint)
      ep.hint_tiles_n = (hf.m[0] + bn - 1) / bn;
      ep.hint_num_tiles = hf.tile_start[1];
      ep.hint_a_row0 = hf.a_row0[0];
      ep.hint_b_row0 = hf.b_row0[0];
    }
    if (bn == 224) return dense ? launch_tc<224, true>(ctx, ap, ep, tiles) : launch_tc<224, false>(ctx, ap, ep, tiles);
    if (bn == 128 && !dense) return launch_tc<128, false>(ctx, ap, ep, tiles);
    if (bn == 64 && !dense) return launch_tc<64, false>(ctx, ap, ep, tiles);
    BT_CHECK(bn == 256, BT_ERR_INVALID, "unsupported association tile width %d", bn);
    return dense ? launch_tc<256, true>(ctx, ap, ep, tiles) : launch_tc<256, false>(ctx, ap, ep, tiles);
  }
  ep.gate_band = 0.f;
  dim3 grid((mx_m + ST - 1) / ST, (mx_n + ST - 1) / ST, ap.count);
  if (ap.a32 != nullptr || ap.d == 0) {
    BT_CHECK(ap.d == 0 || (ap.a32 && ap.b32), BT_ERR_INVALID, "fp32 operands missing");
    if (dense) BT_CUDA(bt_launch(ctx, false, assoc_simt_kernel<true, float>, grid, dim3(256), 0, ap.a32, ap.b32, ep, ap.d));
    else BT_CUDA(bt_launch(ctx, false, assoc_simt_kernel<false, float>, grid, dim3(256), 0, ap.a32, ap.b32, ep, ap.d));
  } else {
    BT_CHECK(ap.a16 && ap.b16, BT_ERR_INVALID, "operands missing");
    if (dense) BT_CUDA(bt_launch(ctx, false, assoc_simt_kernel<true, __half>, grid, dim3(256), 0, ap.a16, ap.b16, ep, ap.d));
    else BT_CUDA(bt_launch(ctx, false, assoc_simt_kernel<false, __half>, grid, dim3(256), 0, ap.a16, ap.b16, ep, ap.d));
  }
  BT_LAUNCHED(ctx);
  return BT_OK;
}

// stand-alone entry points: the frame description goes up through the ctx's descriptor scratch (a pageable
// source is staged by the runtime before cudaMemcpyAsync returns, so the local may die)
int32_t btk_assoc(bt_ctx* ctx, const bt_assoc_params& ap, int32_t precision) {
  bt_assoc_frame hf;
  btk_assoc_fill(ap, precision, &hf);
  bt_assoc_frame* df = reinterpret_cast<bt_assoc_frame*>(ctx->d_desc);
  BT_CUDA(cudaMemcpyAsync(df, &hf, sizeof(hf), cudaMemcpyHostToDevice, ctx->stream));
  return btk_assoc_launch(ctx, ap, precision, hf, df, 0, 0, 0);
}

// profiling aid (BT_ASSOC_DEBUG bit 32768): per-SM CTA lifetimes and hand-over gaps of the launches recorded since the last call
void btk_assoc_stamps_report(void) {
  static unsigned long long h[kStampSMs][kStampDepth][2];
  static int cnt[kStampSMs];
  if (cudaMemcpyFromSymbol(h, g_assoc_stamps, sizeof(h)) != cudaSuccess) return;
  if (cudaMemcpyFromSymbol(cnt, g_assoc_stamp_cnt, sizeof(cnt)) != cudaSuccess) return;
  double life = 0.0, gap = 0.0, gap_max = 0.0;
  long n_life = 0, n_gap = 0;
  for (int sm = 0; sm < kStampSMs; ++sm) {
    const int n = cnt[sm] < kStampDepth ? cnt[sm] : kStampDepth;
    if (cnt[sm] > kStampDepth) continue;          // wrapped: order lost
    for (int i = 0; i < n; ++i) {
      if (h[sm][i][1] > h[sm][i][0]) { life += (double)(h[sm][i][1] - h[sm][i][0]); ++n_life; }
      if (i + 1 < n && h[sm][i + 1][0] > h[sm][i][1]) {
        const double g = (double)(h[sm][i + 1][0] - h[sm][i][1]);
        gap += g; ++n_gap; if (g > gap_max) gap_max = g;
      }
    }
  }
  fprintf(stderr, "assoc stamps: %ld CTA lives, mean %.2f us; %ld per-SM hand-overs (exit of a launch's CTA -> entry of the next launch's CTA on that SM), mean %.2f us, max %.2f us\n",
          n_life, n_life ? life / n_life / 1e3 : 0.0, n_gap, n_gap ? gap / n_gap / 1e3 : 0.0, gap_max / 1e3);
  memset(cnt, 0, sizeof(cnt));
  cudaMemcpyToSymbol(g_assoc_stamp_cnt, cnt, sizeof(cnt));
}
