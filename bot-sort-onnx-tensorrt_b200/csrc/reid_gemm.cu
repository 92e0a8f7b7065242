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
//   warp 0   : TMA producer (cp.async.bulk.tensor, 128B swizzle, 4-stage mbarrier ring)
//   warp 1   : tcgen05.mma issuer (one elected lane), fp32 accumulators in TMEM, 2 stages
//   warps 2-5: epilogue -- tcgen05.ld the accumulator, apply the fusion rules above and emit
//              only the *candidate edges* (cost < stage threshold) into per-row lists.
// The N x M matrices never go to HBM on the tracker path (they are written only for the
// stand-alone entry points / parity dumps): a 2000 x 2000 frame reads 2*8.2 MB of fp16
// features and writes a few thousand edges.  Almost every pair is rejected by an fp32 test
// (no box overlap possible AND appearance gate closed), the exact float64 IoU runs only for
// the survivors.
#include "common.cuh"

#include <cuda.h>
#include <math.h>
#include <type_traits>
#include <stdlib.h>

namespace {

// ------------------------------------------------------------------------------------------------
// epilogue element logic shared by the tensor-core and the CUDA-core kernels
// ------------------------------------------------------------------------------------------------
struct EpiParams {
  const double* row_tlbr;
  const float* row_tlbr_f32;
  const uint8_t* row_kind;
  const double* col_tlbr;
  const uint8_t* col_kind;
  const uint2* col_pk;
  const float* face_sim;
  double match_thresh, second_thresh, unconf_thresh, proximity;
  float appearance;
  bt_cand cand;
  float* out_emb;
  double* out_dists;
  int dense_stage;
  int n, m;
  // BT_ASSOC_DEBUG bits (profiling / bisection aids of the tensor-core kernel, device printf):
  //   1 sampled CTA timestamps   2 similarity pass off   4 every CTA's timestamps   16 epilogue phase times
  //   32 no box pass (every pair evaluated by the similarity pass)   64 no early TMA prologue
  //   1024 globaltimer go/end of every epilogue warp
  int debug;
  float sim_gate;  // smallest similarity for which the appearance gate is open
  float iou_gate;  // a pair without open appearance gate needs IoU > 1 - max(stage thresholds); slightly lowered
};

__device__ __forceinline__ double iou_dist_f64(const double* __restrict__ a, const double* __restrict__ b) {
  const double ixmin = fmax(a[0], b[0]), iymin = fmax(a[1], b[1]);
  const double ixmax = fmin(a[2], b[2]), iymax = fmin(a[3], b[3]);
  if (ixmax <= ixmin || iymax <= iymin) return 1.0;
  const double inter = (ixmax - ixmin) * (iymax - iymin);
  const double area1 = (a[2] - a[0]) * (a[3] - a[1]);
  const double area2 = (b[2] - b[0]) * (b[3] - b[1]);
  return 1.0 - inter / (area1 + area2 - inter);
}

__device__ __forceinline__ double fuse_stage1(double iou_d, float sim, float face, float appearance) {
  float emb = 1.0f - sim;
  const float face_emb = 1.0f - face;
  if (fminf(emb, face_emb) > appearance) emb = 1.0f;
  return fmin(iou_d, (double)emb);
}

__device__ __forceinline__ double fuse_stage3(double iou_d, float sim, float appearance, double proximity) {
  float emb = 1.0f - fmaxf(0.0f, sim);
  if (emb > appearance) emb = 1.0f;
  if (iou_d > proximity) emb = 1.0f;
  return fmin(iou_d, (double)emb);
}

// degree bookkeeping for the LAP's row classification (no return value needed: RED atomics)
__device__ __forceinline__ void note_edge(const bt_cand& c, int list, int row, int col) {
  atomicAdd(&c.rowdeg[(size_t)list * c.rows_cap + row], 1);
  atomicAdd(&c.indeg[(size_t)list * c.cols_cap + col], 1);
  c.rowcol[(size_t)list * c.rows_cap + row] = col;
}
// atomic append into the (row, column-segment) sub-list (CUDA-core kernel: several threads share a row)
__device__ __forceinline__ void emit(const bt_cand& c, int list, int row, int col, double cost) {
  const int seg = col / c.seg;
  const int k = atomicAdd(c.cnt + ((size_t)list * c.rows_cap + row) * c.nseg + seg, 1);
  const size_t base = ((size_t)list * c.rows_cap + row) * (size_t)c.stride + (size_t)seg * c.seg + k;
  c.col[base] = col;
  c.cost[base] = cost;
  atomicAdd(&c.total[list], 1);
  if (k == 0) atomicOr(&c.segmask[(size_t)list * c.rows_cap + row], 1ull << seg);
  note_edge(c, list, row, col);
}
// tensor-core kernel: the owner of (row, segment) adds the row degree once, at the end of its tile
__device__ __forceinline__ void emit_owned_tc(const bt_cand& c, int list, int row, int seg, int k, int col, double cost) {
  const size_t base = ((size_t)list * c.rows_cap + row) * (size_t)c.stride + (size_t)seg * c.seg + k;
  c.col[base] = col;
  c.cost[base] = cost;
  atomicAdd(&c.indeg[(size_t)list * c.cols_cap + col], 1);
}

// exact path for one (row, col) pair that survived the cheap rejection test
__device__ __noinline__ void assoc_exact(const EpiParams& p, int row, int col, float sim, int rkind,
                                         int ckind) {
  const double iou_d = iou_dist_f64(p.row_tlbr + (size_t)row * 4, p.col_tlbr + (size_t)col * 4);
  const float face = p.face_sim ? p.face_sim[(size_t)row * p.m + col] : 0.0f;
  if (rkind == BT_ROW_UNCONFIRMED) {
    if (ckind == BT_COL_HIGH) {
      const double c3 = fuse_stage3(iou_d, sim, p.appearance, p.proximity);
      if (c3 < p.unconf_thresh) emit(p.cand, 2, row, col, c3);
    }
  } else {
    if (ckind == BT_COL_HIGH) {
      const double c1 = fuse_stage1(iou_d, sim, face, p.appearance);
      if (c1 < p.match_thresh) emit(p.cand, 0, row, col, c1);
    } else if (ckind == BT_COL_LOW && rkind == BT_ROW_POOL_TRACKED) {
      if (iou_d < p.second_thresh) emit(p.cand, 1, row, col, iou_d);
    }
  }
}

// Similarity pass of the tensor-core kernel for a pair that is not in the warp's shared-memory slots:
// evaluates it from scratch; if the box pass spilled edges of this row straight to the list (more than
// kPreSlots of them) the pair is looked up there first and re-costed in place.
// Returns 1 when a new edge was appended at position `k_next` of the caller's (row, seg) sub-list.
__device__ __noinline__ int open_pair_slow(const EpiParams& p, double r0, double r1, double r2, double r3,
                                           const double* __restrict__ cbox, int list, int row, int seg, int col,
                                           float sim, int k_next, int scan_from) {
  const double rb[4] = {r0, r1, r2, r3};
  const size_t eb = ((size_t)list * p.cand.rows_cap + row) * (size_t)p.cand.stride + (size_t)seg * p.cand.seg;
  int k_found = -1;
  if (scan_from >= 0)
    for (int j = scan_from; j < k_next; ++j)
      if (p.cand.col[eb + j] == col) k_found = j;
  const double iou_d = iou_dist_f64(rb, cbox);
  double cost = iou_d;
  if (list == 0) {
    const float face = p.face_sim ? p.face_sim[(size_t)row * p.m + col] : 0.0f;
    cost = fuse_stage1(iou_d, sim, face, p.appearance);
  } else if (list == 2) {
    cost = fuse_stage3(iou_d, sim, p.appearance, p.proximity);
  }
  if (k_found >= 0) {
    if (cost != iou_d) p.cand.cost[eb + k_found] = cost;
    return 0;
  }
  const double thr = list == 0 ? p.match_thresh : (list == 1 ? p.second_thresh : p.unconf_thresh);
  if (!(cost < thr)) return 0;
  emit_owned_tc(p.cand, list, row, seg, k_next, col, cost);
  return 1;
}

__device__ __forceinline__ void assoc_dense(const EpiParams& p, int row, int col, float sim) {
  if (p.out_emb) p.out_emb[(size_t)row * p.m + col] = 1.0f - fmaxf(0.0f, sim);
  if (p.out_dists) {
    const double iou_d = iou_dist_f64(p.row_tlbr + (size_t)row * 4, p.col_tlbr + (size_t)col * 4);
    const float face = p.face_sim ? p.face_sim[(size_t)row * p.m + col] : 0.0f;
    p.out_dists[(size_t)row * p.m + col] = (p.dense_stage == 3)
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
// multicast variant: the box lands at the same CTA-relative smem offset of every CTA in cta_mask and
// signals the mbarrier at the same offset in each of them
__device__ __forceinline__ void tma_load_2d_mc(void* smem_dst, const CUtensorMap* tmap, int x, int y,
                                               uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(x), "r"(y),
        "h"(cta_mask)
      : "memory");
}
__device__ __forceinline__ void tcgen05_commit_mc(uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(smem_u32(bar)), "h"(cta_mask)
      : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tcgen05_mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                                uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
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
  static constexpr int kColF64Off = kColKindOff + 256;                // double[BN][4] det box float64 (exact path)
  static constexpr int kPreIouOff = kColF64Off + BN * 32;             // double[kEpiWarps][kPreSlots][32] IoU distance of box-pass edges
  static constexpr int kPreKeyOff = kPreIouOff + 8 * 4 * 32 * 8;      // uint32[kEpiWarps][kPreSlots][32] their (column, position, list)
  static constexpr int kBarOff = kPreKeyOff + 8 * 4 * 32 * 4;         // barriers (8B aligned)
  static constexpr int kNumBars = 2 * kStages + 2 * kAccStages;
  static constexpr int kTmemPtrOff = kBarOff + kNumBars * 8;
  static constexpr int kTotal = kTmemPtrOff + 16;
  static constexpr int kDyn = kTotal + 1024;  // slack for manual 1024 B alignment
};

// Thread-block cluster of CM x CN CTAs (cluster rank = rm * CN + rn) covering CM x CN adjacent output
// tiles.  The CN CTAs of a cluster row need the same A tile and the CM CTAs of a cluster column the
// same B tile: every CTA loads a 1/CN slice of its A tile and a 1/CM slice of its B tile and TMA
// multicasts them to its row / column mates, cutting the L2->SM operand traffic per CTA from
// (BM + BN) to (BM/CN + BN/CM) rows per k-block (the v1 kernel was bound by exactly that traffic,
// profiles/README.md).
template <int BN, bool kDense, int CM, int CN>
__global__ void __launch_bounds__(kTcThreads, 1)
assoc_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                const __grid_constant__ EpiParams p, int d) {   // grid constant: &p (open_pair_slow) needs no local copy
  using L = TcSmem<BN>;
  constexpr int CS = CM * CN;
  constexpr int kASlice = BM / CN, kBSlice = BN / CM;  // rows each CTA loads itself
  static_assert(kASlice % 8 == 0 && kBSlice % 8 == 0, "slices must keep the 8-row swizzle atoms whole");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOff);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full = empty_bar + kStages;
  uint64_t* tmem_empty = tmem_full + kAccStages;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + L::kTmemPtrOff);
  uint2* s_colpk = reinterpret_cast<uint2*>(smem + L::kColPkOff);
  uint8_t* s_colkind = smem + L::kColKindOff;
  double* s_preiou = reinterpret_cast<double*>(smem + L::kPreIouOff);
  uint32_t* s_prekey = reinterpret_cast<uint32_t*>(smem + L::kPreKeyOff);
  double* s_col64 = reinterpret_cast<double*>(smem + L::kColF64Off);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long t_start = clock64();
  unsigned long long g_start = 0;
  if (p.debug) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g_start));
  const int tiles_m = (p.n + BM - 1) / BM, tiles_n = (p.m + BN - 1) / BN;
  const int num_kb = d / BK;
  // cluster geometry: cluster `cid` walks "cluster tiles" of CM x CN output tiles; tiles past the
  // matrix edge are still executed by their CTA (TMA zero-fills, the epilogue masks) so that every
  // CTA of a cluster runs the same number of pipeline steps.
  const int crank = (CS > 1) ? (int)cluster_ctarank() : 0;
  const int rm = crank / CN, rn = crank % CN;
  const int cid = blockIdx.x / CS, num_clusters = gridDim.x / CS;
  const int ctiles_n = (tiles_n + CN - 1) / CN;
  const int num_ctiles = ((tiles_m + CM - 1) / CM) * ctiles_n;
  uint16_t row_mask = 0, col_mask = 0;   // my cluster-row mates (share A), my cluster-column mates (share B)
#pragma unroll
  for (int j = 0; j < CN; ++j) row_mask |= (uint16_t)(1u << (rm * CN + j));
#pragma unroll
  for (int i = 0; i < CM; ++i) col_mask |= (uint16_t)(1u << (i * CN + rn));

  // Programmatic dependent launch: the next kernel of the stream may be scheduled as soon as SMs free
  // up (it blocks in its own griddepcontrol.wait until this grid has completed and flushed).
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  int pro_kb = 0;   // k-blocks of the first tile already requested by the prologue below
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_a)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_b)) : "memory");
    // a slot is free again when every CTA that receives my slices has consumed it
    for (int i = 0; i < kStages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], CM + CN - 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (CS == 1 && cid < num_ctiles && !(p.debug & 64)) {
      // the pipeline's first kStages loads do not wait for anybody: request them before the TMEM
      // allocation and the CTA-wide sync so that their latency overlaps the rest of the ramp
      asm volatile("griddepcontrol.wait;" ::: "memory");   // the operands come from the previous kernels
      const int m0 = (cid / ctiles_n) * BM, n0 = (cid % ctiles_n) * BN;
      pro_kb = num_kb < kStages ? num_kb : kStages;
      for (int kb = 0; kb < pro_kb; ++kb) {
        mbar_expect_tx(&full_bar[kb], L::kStageBytes);
        uint8_t* sa = smem + kb * L::kStageBytes;
        tma_load_2d(sa, &tmap_a, kb * BK, m0, &full_bar[kb]);
        tma_load_2d(sa + L::kABytes, &tmap_b, kb * BK, n0, &full_bar[kb]);
      }
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
  if (CS > 1) cluster_sync_all();   // peers' barriers are initialised before any multicast / remote arrive
  else __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int stage = pro_kb % kStages;
      uint32_t phase = pro_kb == kStages ? 1u : 0u;
      if (CS > 1) asm volatile("griddepcontrol.wait;" ::: "memory");
      for (int ct = cid; ct < num_ctiles; ct += num_clusters) {
        const int m0 = ((ct / ctiles_n) * CM + rm) * BM, n0 = ((ct % ctiles_n) * CN + rn) * BN;
        for (int kb = (ct == cid) ? pro_kb : 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_expect_tx(&full_bar[stage], L::kStageBytes);   // the whole stage lands here (own + mates' slices)
          uint8_t* sa = smem + stage * L::kStageBytes;
          uint8_t* sb = sa + L::kABytes;
          if (CN > 1) tma_load_2d_mc(sa + rn * kASlice * 128, &tmap_a, kb * BK, m0 + rn * kASlice, &full_bar[stage], row_mask);
          else tma_load_2d(sa, &tmap_a, kb * BK, m0, &full_bar[stage]);
          if (CM > 1) tma_load_2d_mc(sb + rm * kBSlice * 128, &tmap_b, kb * BK, n0 + rm * kBSlice, &full_bar[stage], col_mask);
          else tma_load_2d(sb, &tmap_b, kb * BK, n0, &full_bar[stage]);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      // instruction descriptor: D fp32 (1<<4), A/B fp16 K-major, N>>3 at [17,23), M>>4 at [24,29)
      const uint32_t idesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int ct = cid; ct < num_ctiles; ct += num_clusters) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          const uint32_t sa = smem_u32(smem + stage * L::kStageBytes);
          const uint64_t adesc = make_kmajor_sw128_desc(sa);
          const uint64_t bdesc = make_kmajor_sw128_desc(sa + L::kABytes);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // advance 16 fp16 = 32 B along K inside the swizzle atom: +2 in the (addr>>4) field
            tcgen05_mma_f16(tmem_d, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc,
                            (uint32_t)((kb | k) != 0));
          }
          // frees the smem slot (here and at every mate that multicasts into it) when these MMAs retire
          if (CS > 1) tcgen05_commit_mc(&empty_bar[stage], (uint16_t)(row_mask | col_mask));
          else tcgen05_commit(&empty_bar[stage]);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
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
    for (int ct = cid; ct < num_ctiles; ct += num_clusters) {
      const int m0 = ((ct / ctiles_n) * CM + rm) * BM, n0 = ((ct % ctiles_n) * CN + rn) * BN;
      const int row = m0 + quarter * 32 + lane;
      int rkind = BT_ROW_NONE;
      double rbox[4] = {0.0, 0.0, 0.0, 0.0};
      uint32_t r2h = 0, r1p = 0;   // packed conservative integer row box (see pack16 below)
      int rix1 = 0, riy1 = 0, rix2 = 0, riy2 = 0;
      float r_area_lb = 0.f;
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
        if (row < p.n) {
          const double2* s = reinterpret_cast<const double2*>(p.row_tlbr + (size_t)row * 4);
          rkind = p.row_kind[row];
          rlo = s[0]; rhi = s[1];
          if (p.row_tlbr_f32) r32 = *reinterpret_cast<const float4*>(p.row_tlbr_f32 + (size_t)row * 4);
        }
        if (c < BN && col < p.m) {
          const double2* s = reinterpret_cast<const double2*>(p.col_tlbr + (size_t)col * 4);
          ck = p.col_kind[col];
          clo = s[0]; chi = s[1];
          if (p.col_pk) cpk = p.col_pk[col];
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
      int last_a = 0, last_b = 0; // column of the latest edge per list (the row's only column if its degree stays 1)
      int spill_a = -1, spill_b = -1;   // 0 once the box pass wrote edges of a crowded row straight to the list
      int dbg_open = 0;
      const int la = (rkind == BT_ROW_UNCONFIRMED) ? 2 : 0;
      const double thr_a = (rkind == BT_ROW_UNCONFIRMED) ? p.unconf_thresh : p.match_thresh;
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
      const bool all_open = p.face_sim != nullptr || (p.debug & 32);   // no box pass: every pair goes through open_pair
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
            const int k = (list == 1) ? cnt_b++ : cnt_a++;
            if (list == 1) { last_b = n0 + lc; spill_b = 0; } else { last_a = n0 + lc; spill_a = 0; }
            emit_owned_tc(p.cand, list, row, seg, k, n0 + lc, iou_d);
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
      // appearance), sim_gate_for()) gets the fused cost: an edge of the box pass is overwritten in place,
      // otherwise the pair is evaluated from scratch.  Lane-local, no shuffles: the lane owns its row.
      auto open_pair = [&](const int lc, const float sim) {
        const int list = list_of(s_colkind[lc]);
        if (list < 0) return;
        const uint32_t want = (uint32_t)lc | ((uint32_t)(list == 1) << 16);
        bool found = false;
#pragma unroll
        for (int j = 0; j < kPreSlots; ++j) {
          if (j < npre && my_prekey[j * 32 + lane] == want) {
            // an edge of the box pass whose gate turned out open (slots exist only without a face term)
            const double iou_d = my_preiou[j * 32 + lane];
            if (list == 0) my_preiou[j * 32 + lane] = fuse_stage1(iou_d, sim, 0.0f, p.appearance);
            if (list == 2) my_preiou[j * 32 + lane] = fuse_stage3(iou_d, sim, p.appearance, p.proximity);
            found = true;
          }
        }
        if (found) return;
        const int cnt = (list == 1) ? cnt_b : cnt_a;
        const int added = open_pair_slow(p, rbox[0], rbox[1], rbox[2], rbox[3], s_col64 + lc * 4, list, row, seg, n0 + lc, sim,
                                         cnt, (list == 1) ? spill_b : spill_a);
        if (added) { if (list == 1) { ++cnt_b; last_b = n0 + lc; } else { ++cnt_a; last_a = n0 + lc; } }
      };
      long long t_l0 = 0, t_l1 = 0;
      if (kDense) {
#pragma unroll 1
        for (int ch = 0; ch < kChunks; ++ch) {
          issue(ch, va);
          tmem_ld_wait_dep(va);
          const int W = (ch < kFullChunks) ? 32 : kTailCols;
          if (row < p.n) {
#pragma unroll
            for (int c = 0; c < 32; ++c) {
              const int col = n0 + half * kHalfCols + ch * 32 + c;
              if (c < W && col < p.m) assoc_dense(p, row, col, __uint_as_float(va[c]));
            }
          }
        }
      } else {
        // phase A: chunk maxima, TMEM loads software-pipelined, no branches
        uint32_t gatebits = 0;
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
          if (mx >= p.sim_gate) gatebits |= 1u << ch;
        }
        if (all_open) gatebits = (1u << kChunks) - 1u;
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
        if ((p.debug & 2) || rkind == BT_ROW_NONE) gatebits = 0;
        // phase B: only chunks in which some row of the warp has an open gate are looked at again.
        // One rolled loop (one copy of the code: it runs once or twice per warp and would otherwise be
        // fetched cold every time); the tail chunk is loaded 32 wide and masked.
        uint32_t need = __reduce_or_sync(0xffffffffu, gatebits);
        while (need) {                          // warp-uniform
          const int ch = __ffs(need) - 1;
          need &= need - 1;
          tmem_ld_32x32b_x32(taddr + (uint32_t)(half * kHalfCols + ch * 32), va);
          tmem_ld_wait_dep(va);
          if ((gatebits >> ch) & 1u) {
            uint32_t hot = 0;
#pragma unroll
            for (int c = 0; c < 32; ++c)
              if (__uint_as_float(va[c]) >= p.sim_gate) hot |= 1u << c;
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
              const float sim = __uint_as_float((c & 1) ? s2[1] : s2[0]);
              open_pair(half * kHalfCols + ch * 32 + c, sim);
              ++dbg_open;
            }
          }
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
            if ((key >> 16) & 1u) {
              if (cost < p.second_thresh) { emit_owned_tc(p.cand, 1, row, seg, cnt_b++, col, cost); last_b = col; }
            } else if (cost < thr_a) {
              emit_owned_tc(p.cand, la, row, seg, cnt_a++, col, cost);
              last_a = col;
            }
          }
        }
        if (cnt_a) {
          const size_t ri = (size_t)la * p.cand.rows_cap + row;
          p.cand.cnt[ri * p.cand.nseg + seg] = cnt_a;
          atomicAdd(&p.cand.total[la], cnt_a);
          atomicOr(&p.cand.segmask[ri], 1ull << seg);
          atomicAdd(&p.cand.rowdeg[ri], cnt_a);
          if (cnt_a == 1) p.cand.rowcol[ri] = last_a;
        }
        if (cnt_b) {
          const size_t ri = (size_t)1 * p.cand.rows_cap + row;
          p.cand.cnt[ri * p.cand.nseg + seg] = cnt_b;
          atomicAdd(&p.cand.total[1], cnt_b);
          atomicOr(&p.cand.segmask[ri], 1ull << seg);
          atomicAdd(&p.cand.rowdeg[ri], cnt_b);
          if (cnt_b == 1) p.cand.rowcol[ri] = last_b;
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if ((p.debug & 16) && lane == 0 && (blockIdx.x % 37) == 0)
        printf("EPI cta %d warp %d: go %lld bar1 +%lld staged +%lld bar2 +%lld rowloads +%lld rows +%lld box pass done +%lld acc +%lld ld0 +%lld ld1 +%lld sim pass +%lld tail +%lld (lane 0: box edges %d, open pairs %d)\n", blockIdx.x,
               warp, t_go - t_start, t_s1 - t_go, t_s2 - t_go, t_s3 - t_go, t_rl - t_go, t_rows - t_go, t_shadow - t_go, t_acc - t_go, t_l0 - t_acc, t_l1 - t_l0, t_chunks - t_acc, clock64() - t_chunks, npre, dbg_open);
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
  if (CS > 1) cluster_sync_all();   // no CTA leaves while a mate may still arrive on its barriers
  else __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
  if (p.debug && threadIdx.x == 0 && (blockIdx.x % 37) == 0) {
    unsigned long long g_end;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g_end));
    printf("cta %d: globaltimer start %llu end %llu (ns), lifetime %llu ns = %lld cycles\n", blockIdx.x, g_start, g_end,
           g_end - g_start, clock64() - t_start);
  }
}

// ------------------------------------------------------------------------------------------------
// CUDA-core fp32 kernel (small problems, feature sizes that are not a multiple of 64, and the
// on-device cross-check of the tensor path in the tests)
// ------------------------------------------------------------------------------------------------
constexpr int ST = 64, SK = 16;

template <bool kDense>
__global__ void __launch_bounds__(256)
assoc_simt_kernel(const float* __restrict__ a, const float* __restrict__ b, EpiParams p, int d) {
  __shared__ float sa[SK][ST + 1];
  __shared__ float sb[SK][ST + 1];
  const int row0 = blockIdx.y * ST, col0 = blockIdx.x * ST;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < d; k0 += SK) {
    for (int e = threadIdx.x; e < ST * SK; e += 256) {
      const int r = e / SK, k = e % SK;
      sa[k][r] = (row0 + r < p.n && k0 + k < d) ? a[(size_t)(row0 + r) * d + k0 + k] : 0.f;
      sb[k][r] = (col0 + r < p.m && k0 + k < d) ? b[(size_t)(col0 + r) * d + k0 + k] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < SK; ++k) {
      float av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { av[i] = sa[k][ty * 4 + i]; bv[i] = sb[k][tx * 4 + i]; }
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
  if (!kDense) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = col0 + tx * 4 + j;
      ckind[j] = BT_COL_NONE;
      cpk[j] = make_uint2(0x7fff7fffu, 0x80008000u);
      if (col < p.m) {
        ckind[j] = p.col_kind[col];
        if (p.col_pk) cpk[j] = p.col_pk[col];
        else { const double* c = p.col_tlbr + (size_t)col * 4; cpk[j] = pack16_box(c[0], c[1], c[2], c[3], true); }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = row0 + ty * 4 + i;
    if (row >= p.n) continue;
    const int rkind = (!kDense) ? p.row_kind[row] : 0;
    uint2 rpk = make_uint2(0u, 0u);
    if (!kDense && rkind != BT_ROW_NONE) {
      if (p.row_tlbr_f32) {
        const float4 r = *reinterpret_cast<const float4*>(p.row_tlbr_f32 + (size_t)row * 4);
        rpk = bt_pack16_f32(r.x, r.y, r.z, r.w, false);
      } else {
        const double* r = p.row_tlbr + (size_t)row * 4;
        rpk = pack16_box(r[0], r[1], r[2], r[3], false);
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = col0 + tx * 4 + j;
      if (col >= p.m) continue;
      if (kDense) {
        assoc_dense(p, row, col, acc[i][j]);
      } else {
        if (rkind == BT_ROW_NONE || ckind[j] == BT_COL_NONE) continue;
        const bool overlap = ((rpk.y - cpk[j].x) & (cpk[j].y - rpk.x) & 0x80008000u) == 0x80008000u;
        if (overlap || acc[i][j] >= p.sim_gate || p.face_sim != nullptr) assoc_exact(p, row, col, acc[i][j], rkind, ckind[j]);
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
  bt_gemm_ws* ws = ctx->gemm;
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

template <int BN, bool kDense, int CM, int CN>
static int32_t launch_tc(bt_ctx* ctx, const bt_assoc_params& ap, const EpiParams& ep) {
  constexpr int CS = CM * CN;
  CUtensorMap ta, tb;
  BT_TRY(make_tmap(ctx, &ta, ap.a16, ap.a_rows_alloc > 0 ? ap.a_rows_alloc : ap.n, ap.d, BM / CN));  // box = one CTA's slice
  BT_TRY(make_tmap(ctx, &tb, ap.b16, ap.b_rows_alloc > 0 ? ap.b_rows_alloc : ap.m, ap.d, BN / CM));
  auto kern = assoc_tc_kernel<BN, kDense, CM, CN>;
  static int cached_clusters = -1;   // per instantiation: attribute + occupancy query only once
  if (cached_clusters < 0)
    BT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TcSmem<BN>::kDyn));
  const int tiles_m = (ap.n + BM - 1) / BM, tiles_n = (ap.m + BN - 1) / BN;
  const int ctiles = ((tiles_m + CM - 1) / CM) * ((tiles_n + CN - 1) / CN);
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(kTcThreads);
  cfg.dynamicSmemBytes = TcSmem<BN>::kDyn;
  cfg.stream = ctx->stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (cached_clusters < 0) {
    int q = ctx->num_sms / CS;
    if (CS > 1) {
      cfg.gridDim = dim3(CS);
      BT_CUDA(cudaOccupancyMaxActiveClusters(&q, kern, &cfg));
      BT_CHECK(q > 0, BT_ERR_CUDA, "cluster of %d CTAs cannot be scheduled", CS);
    }
    cached_clusters = q;
  }
  const int max_clusters = cached_clusters;
  const int nclusters = ctiles < max_clusters ? ctiles : max_clusters;
  cfg.gridDim = dim3(nclusters * CS);
  if (CS == 1) {
    // no cluster attribute; may start its ramp under the previous kernel's tail (griddepcontrol.wait inside)
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.numAttrs = ctx->pdl ? 1 : 0;
  }
  BT_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tb, ep, ap.d));
  BT_LAUNCHED(ctx);
  return BT_OK;
}

// Cluster shape of the association GEMM.  BT_ASSOC_CLUSTER=CMxCN overrides (profiling sweeps).
static void pick_cluster(int* cm, int* cn) {
  *cm = 1; *cn = 1;   // measured: multicast clusters do not pay (the kernel is MMA/epilogue bound, profiles/README.md)
  const char* e = getenv("BT_ASSOC_CLUSTER");
  if (e && e[0] >= '1' && e[0] <= '4' && e[1] == 'x' && e[2] >= '1' && e[2] <= '4') { *cm = e[0] - '0'; *cn = e[2] - '0'; }
}

int32_t btk_assoc_pick_bn(const bt_ctx* ctx, int32_t n, int32_t m) {
  const char* e = getenv("BT_ASSOC_BN");
  if (e) return atoi(e) == 224 ? 224 : 256;
  const int tiles_m = (n + BM - 1) / BM;
  long best_cost = -1;
  int best = 256;
  for (int bn : {256, 224}) {
    const long tiles = (long)tiles_m * ((m + bn - 1) / bn);
    const long waves = (tiles + ctx->num_sms - 1) / ctx->num_sms;
    const long cost = waves * bn;
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = bn; }
  }
  return best;
}

template <bool kDense>
static int32_t launch_tc_cluster(bt_ctx* ctx, const bt_assoc_params& ap, const EpiParams& ep) {
  int cm, cn;
  pick_cluster(&cm, &cn);
  if (ap.bn == 224) return launch_tc<224, kDense, 1, 1>(ctx, ap, ep);
  if (cm == 1 && cn == 1) return launch_tc<256, kDense, 1, 1>(ctx, ap, ep);
  if (cm == 1 && cn == 2) return launch_tc<256, kDense, 1, 2>(ctx, ap, ep);
  if (cm == 2 && cn == 1) return launch_tc<256, kDense, 2, 1>(ctx, ap, ep);
  if (cm == 2 && cn == 2) return launch_tc<256, kDense, 2, 2>(ctx, ap, ep);
  if (cm == 4 && cn == 2) return launch_tc<256, kDense, 4, 2>(ctx, ap, ep);
  if (cm == 2 && cn == 4) return launch_tc<256, kDense, 2, 4>(ctx, ap, ep);
  if (cm == 4 && cn == 1) return launch_tc<256, kDense, 4, 1>(ctx, ap, ep);
  if (cm == 1 && cn == 4) return launch_tc<256, kDense, 1, 4>(ctx, ap, ep);
  return bt_fail(ctx, BT_ERR_INVALID, "unsupported BT_ASSOC_CLUSTER %dx%d", cm, cn);
}

int32_t btk_assoc(bt_ctx* ctx, const bt_assoc_params& ap, int32_t precision) {
  if (ap.n <= 0 || ap.m <= 0) return BT_OK;
  EpiParams ep;
  ep.row_tlbr = ap.row_tlbr; ep.row_tlbr_f32 = ap.row_tlbr_f32; ep.row_kind = ap.row_kind;
  ep.col_tlbr = ap.col_tlbr; ep.col_kind = ap.col_kind; ep.col_pk = ap.col_pk; ep.face_sim = ap.face_sim;
  ep.match_thresh = ap.match_thresh; ep.second_thresh = ap.second_thresh;
  ep.unconf_thresh = ap.unconf_thresh; ep.proximity = ap.proximity; ep.appearance = ap.appearance;
  ep.cand = ap.cand; ep.out_emb = ap.out_emb; ep.out_dists = ap.out_dists;
  ep.dense_stage = ap.dense_stage; ep.n = ap.n; ep.m = ap.m;
  ep.debug = getenv("BT_ASSOC_DEBUG") ? atoi(getenv("BT_ASSOC_DEBUG")) : 0;
  ep.sim_gate = sim_gate_for(ap.appearance);
  {
    double mx = ap.match_thresh > ap.second_thresh ? ap.match_thresh : ap.second_thresh;
    if (ap.unconf_thresh > mx) mx = ap.unconf_thresh;
    const double g = (1.0 - mx) * 0.999 - 1e-6;      // safely below the smallest IoU any stage can accept
    ep.iou_gate = g > 0.0 ? (float)g : 0.0f;
  }
  const bool dense = (ap.out_emb != nullptr) || (ap.out_dists != nullptr);
  if (precision == 0) {
    BT_CHECK(ap.d % BK == 0 && ap.d >= BK, BT_ERR_INVALID,
             "tensor-core similarity needs feat_dim %% 64 == 0 (got %d)", ap.d);
    BT_CHECK(ap.a16 && ap.b16, BT_ERR_INVALID, "fp16 operands missing");
    BT_CHECK(dense || ap.cand.seg * 2 == (ap.bn ? ap.bn : 256), BT_ERR_STATE,
             "candidate segment size %d does not match the tile width %d", ap.cand.seg, ap.bn ? ap.bn : 256);
    if (dense) return launch_tc_cluster<true>(ctx, ap, ep);
    return launch_tc_cluster<false>(ctx, ap, ep);
  }
  BT_CHECK(ap.a32 && ap.b32, BT_ERR_INVALID, "fp32 operands missing");
  dim3 grid((ap.m + ST - 1) / ST, (ap.n + ST - 1) / ST);
  if (dense) assoc_simt_kernel<true><<<grid, 256, 0, ctx->stream>>>(ap.a32, ap.b32, ep, ap.d);
  else assoc_simt_kernel<false><<<grid, 256, 0, ctx->stream>>>(ap.a32, ap.b32, ep, ap.d);
  BT_LAUNCHED(ctx);
  return BT_OK;
}
