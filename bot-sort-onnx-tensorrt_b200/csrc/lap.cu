// Exact linear assignment on the GPU, replacing
//     lap.lapjv(cost_matrix, extend_cost=True, cost_limit=thresh)         (demo:1686)
// behind linear_assignment(cost_matrix, thresh) (demo:1682-1693;
// demo = /root/reference/demo_bottrack_onnx_tflite.py; lap==0.4.0 is a third-party package that
// is not vendored in the reference).
//
// lap's problem: the N x M cost is embedded in an (N+M)^2 matrix padded with
// thresh/2 and solved exactly, which is the same as minimising  sum_matched (c_ij - thresh):
// a max-weight (not perfect) bipartite matching over the edges with c_ij < thresh.  Instead of
// the dense 4000 x 4000 float64 matrix of the reference (128 MB at 2000 x 2000) this solver
// works on the *candidate edge lists* the association epilogue emits.
// ONE launch (`lap_stream_kernel`), ONE CTA of 1024 threads per video stream, solves the frame's three
// chained association stages (stage 2 masked by stage 1's matched rows, stage 3 by stage 1's matched
// columns); phases are separated by block barriers:
//   P1  every row is classified from the emitters' degree bookkeeping: isolated edges (row degree 1,
//       column in-degree 1) are matched on the spot -- the bulk of a tracking scene
//   S2  the remaining "complex" rows' edges are gathered into shared memory (warp per row)
//   S3  edges whose cost carries the tensor-core similarity are re-costed exactly (fp32 operands,
//       float64 accumulation; SURVEY hard part 2) -- so a near-tie between two tracks is decided by the
//       same numbers the reference sees, not by fp16 rounding
//   S4+ exact shortest-augmenting-path assignment (Jonker-Volgenant / Crouse formulation, float64 duals,
//       a private zero-cost dummy column per row = "stay unmatched"): a handful of rows by one warp, more
//       by labelling connected components (min-label propagation + pointer jumping) and solving them in
//       parallel (enumeration for <= 3 rows, one thread for <= 6, one warp above).
// Complex parts that do not fit on chip (> 512 rows, > 4096 edges, > 2560 detections: a crowded scene, a
// dense adversarial matrix) take the same algorithm over global scratch: slower, still exact.
// Ties between equal-cost optima are broken by lowest index, lap's own tie-breaking is not
// reproducible without its sources: "bit-exact" is defined on inputs with a unique optimum.
#include "common.cuh"

#include <float.h>
#include <stdlib.h>
#include <mutex>
#include <vector>

struct bt_lap_ws {
  bt_cand cand;            // ctx-wide candidate lists (3 lists per video stream)
  // scratch of the large-problem path, one slice per video stream (rows_stride / cols_stride entries apart)
  int32_t* label = nullptr;     // [rows]
  int32_t* collabel = nullptr;  // [cols]
  int32_t* clist = nullptr;     // [rows]  complex rows
  unsigned long long* rmaskg = nullptr;  // [rows] their segment masks
  int32_t* compidx = nullptr;   // [rows]  component index of a root row (indexed by row)
  int32_t* isroot = nullptr;    // [rows]  scan scratch (indexed by clist position)
  int32_t* rowcnt = nullptr;    // [rows+1] per component -> exclusive scan = row_start
  int32_t* colcnt = nullptr;    // [rows+1] per component -> exclusive scan = col_start
  int32_t* fill = nullptr;      // [rows+1]
  int32_t* sorted_rows = nullptr;  // [rows]
  int32_t* counters = nullptr;     // [8]: 1 ncomp, 2/3 changed flags
  double* u = nullptr;             // [rows]
  double* v = nullptr;             // [cols]
  double* dist = nullptr;          // [cols]
  int32_t* pathrow = nullptr;      // [cols]
  int32_t* seen = nullptr;         // [cols] stamp
  int32_t* insc = nullptr;         // [cols] stamp
  int32_t* touched = nullptr;      // [cols] scratch, sliced per component
  int32_t* treerows = nullptr;     // [rows] scratch, sliced per component
  int32_t* x = nullptr;            // [rows] own outputs for the dense API
  int32_t* y = nullptr;            // [cols]
  int rows = 0, cols = 0;
  int rows_stride = 0, cols_stride = 0;
  // a two-row dummy problem (never cleared) + its outputs: the kernel's instruction-cache warm-up run
  bt_cand warm = {};
  int32_t* warm_xy = nullptr;      // [streams][8]
  char* warm_blk = nullptr;
};

namespace {

constexpr int kLapThreads = 512;   // 128 registers per thread: the phases are latency chains, spills to local memory cost L2 round trips
constexpr int kInf = 0x7fffffff;

struct LapParams {
  int nstages;          // 1 (stand-alone: list `list0`) or 3 (the frame's chained stages: lists 0, 1, 2)
  int list0;
  double thresh[3];
  int clear_lists;      // tracker mode: leave segmask / rowdeg / indeg / total zeroed for the next frame
  int clear_cnt;        // ... and the segment counters too (only the CUDA-core emitter, which appends with atomics, needs
                        // them zeroed; the tensor-core epilogue overwrites the counters of the segments it flags)
  int debug;            // BT_LAP_DEBUG=1: phase timestamps (ns) by device printf
};

__device__ __forceinline__ bool edge_ok(const int32_t* __restrict__ col_block, int c) {
  return col_block == nullptr || col_block[c] < 0;
}

// exclusive scan of data[0..n) in place by ONE CTA, returns total
__device__ int block_exclusive_scan(int32_t* data, int n, int32_t* s_warp, int32_t* s_carry) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) *s_carry = 0;
  __syncthreads();
  for (int base = 0; base < n; base += kLapThreads) {
    const int i = base + tid;
    const int val = (i < n) ? data[i] : 0;
    int incl = val;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int w = (lane < kLapThreads / 32) ? s_warp[lane] : 0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += t;
      }
      s_warp[lane] = w;
    }
    __syncthreads();
    const int carry = *s_carry;
    const int warp_off = (warp == 0) ? 0 : s_warp[warp - 1];
    if (i < n) data[i] = carry + warp_off + incl - val;
    __syncthreads();
    if (tid == kLapThreads - 1) *s_carry = carry + s_warp[kLapThreads / 32 - 1];
    __syncthreads();
  }
  return *s_carry;
}

struct MinPair { double d; int c; int freecol; };

__device__ __forceinline__ bool better(const MinPair& a, const MinPair& b) {
  // smaller distance; ties: unassigned column first, then lower column index
  if (a.d != b.d) return a.d < b.d;
  if (a.freecol != b.freecol) return a.freecol > b.freecol;
  return a.c < b.c;
}

// Arrays the component solver works on: global scratch (large problems) or shared memory of CTA 0
// with local row / column ids (the usual case: a handful of complex rows).
struct SolveArrays {
  const int32_t* rowcnt; const int32_t* colcnt;   // component -> row-list start, scratch start
  int32_t* sorted_rows; int32_t* touched; int32_t* treerows;
  double* u; double* v; double* dist;
  int32_t* pathrow; int32_t* seen; int32_t* insc;
  const int32_t* rstart;   // CSR edge start per row, or nullptr: row * stride
  size_t stride;
};
__device__ __forceinline__ size_t ebase(const SolveArrays& a, int row) {
  return a.rstart ? (size_t)a.rstart[row] : (size_t)row * a.stride;
}

// one warp solves one connected component exactly
__device__ void solve_component(const SolveArrays& ws, int comp, double thresh,
                                const int32_t* col_block, const int32_t* cnt,
                                const int32_t* ecol, const double* ecost,
                                int32_t* x, int32_t* y, int lane) {
  {
    const int r0 = ws.rowcnt[comp], nr = ws.rowcnt[comp + 1] - r0;
    const int c0 = ws.colcnt[comp];
    int32_t* rows = ws.sorted_rows + r0;
    int32_t* touched = ws.touched + c0;
    int32_t* treerows = ws.treerows + r0;

    if (nr == 1) {
      // star component: the single row takes its cheapest valid column
      const int r = rows[0];
      const int deg = cnt[r];
      MinPair best{DBL_MAX, kInf, 0};
      for (int k = lane; k < deg; k += 32) {
        const int c = ecol[ebase(ws, r) + k];
        if (!edge_ok(col_block, c)) continue;
        MinPair cur{ecost[ebase(ws, r) + k], c, 0};
        if (better(cur, best)) best = cur;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        MinPair oth{__shfl_xor_sync(0xffffffffu, best.d, o), __shfl_xor_sync(0xffffffffu, best.c, o), 0};
        if (better(oth, best)) best = oth;
      }
      // (an edge that does not beat "stay unmatched" -- a column blocked by an earlier stage, a re-costed
      //  edge that turned out too expensive -- is no match)
      if (lane == 0 && best.c != kInf && best.d < thresh) { x[r] = best.c; y[best.c] = r; }
      return;
    }

    // deterministic processing order: ascending row index (rank sort of the scattered list)
    if (nr <= 32) {
      int mine = (lane < nr) ? rows[lane] : kInf;
      int rank = 0;
      for (int k = 0; k < nr; ++k) {
        const int o = __shfl_sync(0xffffffffu, mine, k);
        if (o < mine) ++rank;
      }
      __syncwarp();
      if (lane < nr) rows[rank] = mine;
    } else {
      // rank sort through the tree-row scratch
      for (int a = lane; a < nr; a += 32) {
        const int mine = rows[a];
        int rank = 0;
        for (int k = 0; k < nr; ++k) rank += (rows[k] < mine);
        treerows[rank] = mine;
      }
      __syncwarp();
      for (int a = lane; a < nr; a += 32) rows[a] = treerows[a];
    }
    __syncwarp();

    for (int ri = 0; ri < nr; ++ri) {
      const int i0 = rows[ri];
      const int sid = i0 + 1;  // unique search stamp
      int nT = 0, nTR = 0;
      int i = i0;
      double minVal = 0.0;
      double bestDummy = -ws.u[i0];
      int bestDummyRow = i0;
      int sink = -1;       // >= 0: real column; -2: dummy of bestDummyRow
      while (true) {
        // ---- relax the edges of row i ----
        const int deg = cnt[i];
        const double ui = ws.u[i];
        for (int k0 = 0; k0 < deg; k0 += 32) {
          const int k = k0 + lane;
          bool fresh = false;
          int c = -1;
          if (k < deg) {
            c = ecol[ebase(ws, i) + k];
            if (edge_ok(col_block, c) && ws.insc[c] != sid) {
              const double r = minVal + (ecost[ebase(ws, i) + k] - thresh) - ui - ws.v[c];
              if (ws.seen[c] != sid) {
                ws.seen[c] = sid;
                ws.dist[c] = r;
                ws.pathrow[c] = i;
                fresh = true;
              } else if (r < ws.dist[c]) {
                ws.dist[c] = r;
                ws.pathrow[c] = i;
              }
            }
          }
          const unsigned ball = __ballot_sync(0xffffffffu, fresh);
          if (fresh) touched[nT + __popc(ball & ((1u << lane) - 1))] = c;
          nT += __popc(ball);
        }
        __syncwarp();
        // ---- closest touched column outside the scanned set ----
        MinPair best{DBL_MAX, kInf, 0};
        for (int k = lane; k < nT; k += 32) {
          const int c = touched[k];
          if (ws.insc[c] == sid) continue;
          MinPair cur{ws.dist[c], c, (y[c] < 0) ? 1 : 0};
          if (better(cur, best)) best = cur;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          MinPair oth{__shfl_xor_sync(0xffffffffu, best.d, o), __shfl_xor_sync(0xffffffffu, best.c, o),
                      __shfl_xor_sync(0xffffffffu, best.freecol, o)};
          if (better(oth, best)) best = oth;
        }
        if (best.c == kInf || bestDummy <= best.d) {
          minVal = bestDummy;
          sink = -2;
          break;
        }
        minVal = best.d;
        const int j = best.c;
        __syncwarp();                       // every lane's scan reads of insc[] are behind us (racecheck: WAR)
        if (lane == 0) ws.insc[j] = sid;
        __syncwarp();
        if (best.freecol) { sink = j; break; }
        i = y[j];
        if (lane == 0) treerows[nTR] = i;
        ++nTR;
        const double cand_d = minVal - ws.u[i];
        if (cand_d < bestDummy) { bestDummy = cand_d; bestDummyRow = i; }
        __syncwarp();
      }
      // ---- dual updates (Crouse 2016, eq. step 4) ----
      __syncwarp();
      for (int k = lane; k < nTR; k += 32) {
        const int r = treerows[k];
        ws.u[r] += minVal - ws.dist[x[r]];
      }
      for (int k = lane; k < nT; k += 32) {
        const int c = touched[k];
        if (ws.insc[c] == sid) ws.v[c] -= minVal - ws.dist[c];
      }
      __syncwarp();
      if (lane == 0) {
        ws.u[i0] += minVal;
        // ---- augment ----
        int j;
        bool go = true;
        if (sink == -2) {
          if (bestDummyRow == i0) go = false;
          j = go ? x[bestDummyRow] : -1;
          if (go) x[bestDummyRow] = -1;
        } else {
          j = sink;
        }
        while (go) {
          const int pi = ws.pathrow[j];
          y[j] = pi;
          const int t = x[pi];
          x[pi] = j;
          j = t;
          if (pi == i0) break;
        }
      }
      __syncwarp();
    }
  }
}

// One THREAD solves one tiny component (a few rows): same shortest-augmenting-path algorithm as
// solve_component without the warp machinery -- for 2-3 row components the shuffles and warp
// barriers of the cooperative version cost far more than the arithmetic.
__device__ void solve_component_serial(const SolveArrays& ws, int comp, double thresh, const int32_t* cnt,
                                       const int32_t* ecol, const double* ecost, int32_t* x, int32_t* y) {
  const int r0 = ws.rowcnt[comp], nr = ws.rowcnt[comp + 1] - r0;
  const int c0 = ws.colcnt[comp];
  int32_t* rows = ws.sorted_rows + r0;
  int32_t* touched = ws.touched + c0;
  int32_t* treerows = ws.treerows + r0;
  for (int a = 1; a < nr; ++a) {            // insertion sort: ascending row index (deterministic order)
    const int key = rows[a];
    int b = a - 1;
    while (b >= 0 && rows[b] > key) { rows[b + 1] = rows[b]; --b; }
    rows[b + 1] = key;
  }
  for (int ri = 0; ri < nr; ++ri) {
    const int i0 = rows[ri];
    const int sid = i0 + 1;
    int nT = 0, nTR = 0, i = i0, sink = -1, bestDummyRow = i0;
    double minVal = 0.0, bestDummy = -ws.u[i0];
    while (true) {
      const int deg = cnt[i];
      const double ui = ws.u[i];
      const size_t eb = ebase(ws, i);
      for (int k = 0; k < deg; ++k) {
        const int c = ecol[eb + k];
        if (ws.insc[c] == sid) continue;
        const double r = minVal + (ecost[eb + k] - thresh) - ui - ws.v[c];
        if (ws.seen[c] != sid) {
          ws.seen[c] = sid; ws.dist[c] = r; ws.pathrow[c] = i; touched[nT++] = c;
        } else if (r < ws.dist[c]) {
          ws.dist[c] = r; ws.pathrow[c] = i;
        }
      }
      MinPair best{DBL_MAX, kInf, 0};
      for (int k = 0; k < nT; ++k) {
        const int c = touched[k];
        if (ws.insc[c] == sid) continue;
        const MinPair cur{ws.dist[c], c, (y[c] < 0) ? 1 : 0};
        if (better(cur, best)) best = cur;
      }
      if (best.c == kInf || bestDummy <= best.d) { minVal = bestDummy; sink = -2; break; }
      minVal = best.d;
      const int j = best.c;
      ws.insc[j] = sid;
      if (best.freecol) { sink = j; break; }
      i = y[j];
      treerows[nTR++] = i;
      const double cand_d = minVal - ws.u[i];
      if (cand_d < bestDummy) { bestDummy = cand_d; bestDummyRow = i; }
    }
    for (int k = 0; k < nTR; ++k) { const int r = treerows[k]; ws.u[r] += minVal - ws.dist[x[r]]; }
    for (int k = 0; k < nT; ++k) { const int c = touched[k]; if (ws.insc[c] == sid) ws.v[c] -= minVal - ws.dist[c]; }
    ws.u[i0] += minVal;
    int j;
    bool go = true;
    if (sink == -2) {
      if (bestDummyRow == i0) go = false;
      j = go ? x[bestDummyRow] : -1;
      if (go) x[bestDummyRow] = -1;
    } else {
      j = sink;
    }
    while (go) {
      const int pi = ws.pathrow[j];
      y[j] = pi;
      const int t = x[pi];
      x[pi] = j;
      j = t;
      if (pi == i0) break;
    }
  }
}

// Components of two or three rows (by far the most common complex case: two tracks competing for one
// or two detections) are solved by exhaustive enumeration in registers: every row takes one of its
// edges or stays unmatched, columns must be distinct, minimise sum (cost - thresh).
__device__ void solve_component_enum(const SolveArrays& ws, int comp, double thresh, const int32_t* cnt,
                                     const int32_t* ecol, const double* ecost, int32_t* x, int32_t* y) {
  const int r0 = ws.rowcnt[comp], nr = ws.rowcnt[comp + 1] - r0;
  int rows[3], deg[3];
  size_t eb[3];
  for (int a = 0; a < 3; ++a) {
    rows[a] = (a < nr) ? ws.sorted_rows[r0 + a] : -1;
    deg[a] = (a < nr) ? cnt[rows[a]] : 0;
    eb[a] = (a < nr) ? ebase(ws, rows[a]) : 0;
  }
  double best = 0.0;                 // everything unmatched
  int bk[3] = {-1, -1, -1};
  for (int k0 = -1; k0 < deg[0]; ++k0) {
    const int c0 = k0 < 0 ? -1 : ecol[eb[0] + k0];
    const double w0 = k0 < 0 ? 0.0 : ecost[eb[0] + k0] - thresh;
    for (int k1 = -1; k1 < deg[1]; ++k1) {
      const int c1 = k1 < 0 ? -1 : ecol[eb[1] + k1];
      if (c1 >= 0 && c1 == c0) continue;
      const double w1 = w0 + (k1 < 0 ? 0.0 : ecost[eb[1] + k1] - thresh);
      for (int k2 = -1; k2 < deg[2]; ++k2) {
        const int c2 = k2 < 0 ? -1 : ecol[eb[2] + k2];
        if (c2 >= 0 && (c2 == c0 || c2 == c1)) continue;
        const double w2 = w1 + (k2 < 0 ? 0.0 : ecost[eb[2] + k2] - thresh);
        if (w2 < best) { best = w2; bk[0] = k0; bk[1] = k1; bk[2] = k2; }
      }
    }
  }
  for (int a = 0; a < nr; ++a)
    if (bk[a] >= 0) {
      const int c = ecol[eb[a] + bk[a]];
      x[rows[a]] = c;
      y[c] = rows[a];
    }
}

// ---- exact re-costing of flagged edges (SURVEY hard part 2) ----------------------------------------------
// One warp evaluates one (row, column) pair from scratch: exact float64 IoU distance, the similarity as the
// reference sees it (fp32 values, accumulated in float64: within one ulp of the true dot product, where
// the reference's sgemm is within a few), the face term, and the stage's fusion rule.  The tensor-core
// similarity decided WHICH pairs exist; this decides what they cost whenever the cost can matter.
__device__ double refine_cost(const bt_refine& rf, const bt_lap_batch& B, int k, int list, int row, int col, int lane) {
  const size_t gs = (size_t)B.row0[k] + row, gd = (size_t)B.col0[k] + col, gi = (size_t)B.in0[k] + col;
  const int D = rf.d;
  // everything that does not depend on the dot product first: the loads below are all in flight together
  const float na = rf.f16 ? rf.a_norm[gs] : 1.0f;
  const float nb = (list == 2 && rf.b_norm) ? rf.b_norm[gd] : 1.0f;
  const double iou_d = bt_iou_dist_f64(rf.row_tlbr + gs * 4, rf.col_tlbr + gd * 4);
  float face = 0.f;
  if (list == 0 && B.face_sim[k]) {
    const int pr = reinterpret_cast<const int32_t*>(rf.ctrl + B.pos_off[k])[row];
    if (pr >= 0) face = B.face_sim[k][(size_t)pr * B.m[k] + col];
  }
  double acc = 0.0;
  if (rf.f16) {
    const __half* a = rf.a16 + gs * D;
    const __half* b = rf.b16 + gi * D;
    if ((D & 2047) == 0) {          // 2048-d Fast-ReID rows: 8 x 16 B per lane and operand, all issued at once
      for (int base = 0; base < D; base += 2048) {
        uint4 qa[8], qb[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          qa[j] = *reinterpret_cast<const uint4*>(a + base + j * 256 + lane * 8);
          qb[j] = *reinterpret_cast<const uint4*>(b + base + j * 256 + lane * 8);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const __half2* ha = reinterpret_cast<const __half2*>(&qa[j]);
          const __half2* hb = reinterpret_cast<const __half2*>(&qb[j]);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 fa = __half22float2(ha[e]), fb = __half22float2(hb[e]);
            acc += (double)(fa.x * fb.x) + (double)(fa.y * fb.y);     // fp16 x fp16 is exact in fp32
          }
        }
      }
    } else if ((D & 7) == 0) {
      for (int i = lane * 8; i < D; i += 256) {
        const uint4 qa = *reinterpret_cast<const uint4*>(a + i), qb = *reinterpret_cast<const uint4*>(b + i);
        const __half2* ha = reinterpret_cast<const __half2*>(&qa);
        const __half2* hb = reinterpret_cast<const __half2*>(&qb);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 fa = __half22float2(ha[e]), fb = __half22float2(hb[e]);
          acc += (double)(fa.x * fb.x) + (double)(fa.y * fb.y);
        }
      }
    } else {
      for (int i = lane; i < D; i += 32) acc += (double)(__half2float(a[i]) * __half2float(b[i]));
    }
  } else {
    const float* a = rf.a32 + gs * D;
    const float* b = rf.b32 + gi * D;
    if ((D & 1023) == 0) {
      for (int base = 0; base < D; base += 1024) {
        float4 fa[8], fb[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          fa[j] = *reinterpret_cast<const float4*>(a + base + j * 128 + lane * 4);
          fb[j] = *reinterpret_cast<const float4*>(b + base + j * 128 + lane * 4);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j)
          acc += (double)fa[j].x * fb[j].x + (double)fa[j].y * fb[j].y + (double)fa[j].z * fb[j].z + (double)fa[j].w * fb[j].w;
      }
    } else if ((D & 3) == 0) {
      for (int i = lane * 4; i < D; i += 128) {
        const float4 fa = *reinterpret_cast<const float4*>(a + i), fb = *reinterpret_cast<const float4*>(b + i);
        acc += (double)fa.x * fb.x + (double)fa.y * fb.y + (double)fa.z * fb.z + (double)fa.w * fb.w;
      }
    } else {
      for (int i = lane; i < D; i += 32) acc += (double)a[i] * b[i];
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  float sim;
  if (rf.f16) sim = na > 0.f ? (float)(acc / (double)na) : 0.f;
  else sim = (float)acc;
  if (list == 2) {      // stage 3 compares the NORMALISED detection feature (demo:1593-1599)
    if (rf.b_norm) sim = nb > 0.f ? sim / nb : 0.f;
    return bt_fuse_stage3(iou_d, sim, rf.appearance, rf.proximity);
  }
  return bt_fuse_stage1(iou_d, sim, face, rf.appearance);
}

// ---- on-chip path for the usual case: up to a few hundred complex rows -----------------------------------
// The CTA pulls the complex rows' edges into shared memory (CSR, local row ids in ascending row order,
// GLOBAL column ids: the per-column solver state of up to kSmallCols detections lives in shared memory
// too), re-costs the flagged ones, and solves: a handful of rows by one warp directly, more by labelling
// connected components and solving them in parallel.
constexpr int kSmallRows = 512;
constexpr int kSmallEdges = 4096;
constexpr int kSmallCols = 2560;
constexpr int kOneWarpRows = 4;
constexpr int kSmallPairs = 2048;  // non-empty (row, segment) pairs of the complex rows
constexpr double kTieGap = 1.0e-3;         // two costs closer than this may swap order within the tensor-core error
constexpr double kBlockedCost = 1.0e300;   // an edge whose column an earlier stage took: never beats staying unmatched

// a stage's arguments, in shared memory: by-reference arguments of a real function call would live in the
// caller's LOCAL memory (1024 threads x ~0.5 KB: every field access an L2 round trip)
struct StageArgs {
  bt_cand cand;
  bt_lap_ws W;
  int kb, list, n, m, clear, clear_cnt, need_y, debug;
  double thresh;
  int32_t* x; int32_t* y;
  const int32_t* row_block; const int32_t* col_block;
  unsigned long long* tq;
};

struct SmallSmem {
  StageArgs sa;
  unsigned long long tq[8];
  int32_t rg[kSmallRows], rdeg[kSmallRows], rstart[kSmallRows + 1], rlabel[kSmallRows], xl[kSmallRows];
  int32_t sorted_rows[kSmallRows], treerows[kSmallRows], compidx[kSmallRows], isroot[kSmallRows];
  int32_t rowcnt[kSmallRows + 1], colcnt[kSmallRows + 1], fill[kSmallRows + 1];
  double u[kSmallRows];
  int32_t ecol[kSmallEdges];
  double ecst[kSmallEdges];
  int32_t rlist[kSmallEdges];
  int32_t pairs[kSmallPairs];
  double v[kSmallCols], dist[kSmallCols];
  int32_t pathrow[kSmallCols], seen[kSmallCols], insc[kSmallCols], yl[kSmallCols], clabel[kSmallCols], touched[kSmallCols];
  int32_t s_warp[32];
  int32_t s_carry;
  double red_d[kLapThreads / 32];      // CTA-wide component solver: per-warp minima
  int32_t red_c[kLapThreads / 32], red_f[kLapThreads / 32];
  int nC, nE, nR, nP, changed, big;
};

static_assert(sizeof(SmallSmem) <= 227 * 1024, "the LAP's on-chip state must fit the SM's shared memory");

// Per-column solver state of the CTA-wide component solver: shared memory when the stream's detections fit
// (kSmallCols), the global scratch otherwise.  ymir mirrors y (column -> row) for the columns of the component.
struct ColState {
  double* v; double* dist;
  int32_t* pathrow; int32_t* seen; int32_t* insc; int32_t* ymir;
};

// The WHOLE CTA solves one large connected component exactly (a crowded scene, a dense cost matrix = one giant
// component): the same shortest-augmenting-path algorithm, tie-breaking and floating-point operation order as
// solve_component, with the relaxation of a row's edges and the minimum scan spread over all threads -- one warp
// walked a dense 512 x 512 component in 131 ms (profiles/r02_lap_dense.json).  The scan covers all m columns of the
// stream (stamps tell which are touched): no touched list, no compaction.
__device__ void solve_component_cta(const SolveArrays& ws, int comp, double thresh, int m,
                                    const int32_t* col_block, const int32_t* cnt,
                                    const int32_t* ecol, const double* ecost,
                                    int32_t* x, int32_t* y, const ColState& cs, SmallSmem& sm) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int GT = kLapThreads, NW = kLapThreads / 32;
  const int r0 = ws.rowcnt[comp], nr = ws.rowcnt[comp + 1] - r0;
  int32_t* rows = ws.sorted_rows + r0;
  int32_t* treerows = ws.treerows + r0;
  // deterministic processing order: ascending row index (rank sort through the tree-row scratch)
  for (int a = tid; a < nr; a += GT) {
    const int mine = rows[a];
    int rank = 0;
    for (int k = 0; k < nr; ++k) rank += (rows[k] < mine);
    treerows[rank] = mine;
  }
  __syncthreads();
  for (int a = tid; a < nr; a += GT) rows[a] = treerows[a];
  __syncthreads();

  for (int ri = 0; ri < nr; ++ri) {
    const int i0 = rows[ri];
    const int sid = i0 + 1;  // unique search stamp
    int nTR = 0;
    int i = i0;
    double ui = ws.u[i0];
    double minVal = 0.0;
    double bestDummy = -ui;
    int bestDummyRow = i0;
    int sink = -1;       // >= 0: real column; -2: dummy of bestDummyRow
    while (true) {
      // ---- relax the edges of row i (a column occurs once per row: no write conflicts) ----
      const int deg = cnt[i];
      const size_t eb = ebase(ws, i);
      for (int k = tid; k < deg; k += GT) {
        const int c = ecol[eb + k];
        if (edge_ok(col_block, c) && cs.insc[c] != sid) {
          const double r = minVal + (ecost[eb + k] - thresh) - ui - cs.v[c];
          if (cs.seen[c] != sid) { cs.seen[c] = sid; cs.dist[c] = r; cs.pathrow[c] = i; }
          else if (r < cs.dist[c]) { cs.dist[c] = r; cs.pathrow[c] = i; }
        }
      }
      __syncthreads();
      // ---- closest touched column outside the scanned set ----
      MinPair best{DBL_MAX, kInf, 0};
      for (int c = tid; c < m; c += GT) {
        if (cs.seen[c] != sid || cs.insc[c] == sid) continue;
        MinPair cur{cs.dist[c], c, (cs.ymir[c] < 0) ? 1 : 0};
        if (better(cur, best)) best = cur;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        MinPair oth{__shfl_xor_sync(0xffffffffu, best.d, o), __shfl_xor_sync(0xffffffffu, best.c, o),
                    __shfl_xor_sync(0xffffffffu, best.freecol, o)};
        if (better(oth, best)) best = oth;
      }
      if (lane == 0) { sm.red_d[warp] = best.d; sm.red_c[warp] = best.c; sm.red_f[warp] = best.freecol; }
      __syncthreads();
      best = MinPair{sm.red_d[0], sm.red_c[0], sm.red_f[0]};
#pragma unroll
      for (int w = 1; w < NW; ++w) {
        MinPair oth{sm.red_d[w], sm.red_c[w], sm.red_f[w]};
        if (better(oth, best)) best = oth;
      }
      if (best.c == kInf || bestDummy <= best.d) {
        minVal = bestDummy;
        sink = -2;
        break;
      }
      minVal = best.d;
      const int j = best.c;
      if (tid == 0) cs.insc[j] = sid;
      if (best.freecol) { sink = j; break; }
      i = cs.ymir[j];
      if (tid == 0) treerows[nTR] = i;
      ++nTR;
      ui = ws.u[i];
      const double cand_d = minVal - ui;
      if (cand_d < bestDummy) { bestDummy = cand_d; bestDummyRow = i; }
      __syncthreads();                    // insc[j] is set before the next relaxation reads it; the minima slots are free again
    }
    // ---- dual updates (Crouse 2016, eq. step 4) ----
    __syncthreads();
    for (int k = tid; k < nTR; k += GT) {
      const int r = treerows[k];
      ws.u[r] += minVal - cs.dist[x[r]];
    }
    for (int c = tid; c < m; c += GT)
      if (cs.insc[c] == sid) cs.v[c] -= minVal - cs.dist[c];
    __syncthreads();
    if (tid == 0) {
      ws.u[i0] += minVal;
      // ---- augment ----
      int j;
      bool go = true;
      if (sink == -2) {
        if (bestDummyRow == i0) go = false;
        j = go ? x[bestDummyRow] : -1;
        if (go) x[bestDummyRow] = -1;
      } else {
        j = sink;
      }
      while (go) {
        const int pi = cs.pathrow[j];
        y[j] = pi;
        cs.ymir[j] = pi;
        const int t = x[pi];
        x[pi] = j;
        j = t;
        if (pi == i0) break;
      }
    }
    __syncthreads();
  }
}

// One association stage of one video stream.  A real (noinline) function on purpose: the kernel calls it once on
// a two-row dummy problem BEFORE griddepcontrol.wait -- every phase of this latency-bound kernel runs exactly once
// per launch, so its time used to be dominated by cold instruction fetches; the dry run pulls the hot path's
// instructions into the SM's instruction cache while the association kernel is still running.
__device__ __noinline__ void lap_stage(const bt_lap_batch& B, const bt_refine& rf, SmallSmem& sm) {
  const StageArgs& a = sm.sa;
  const bt_cand& cand = a.cand;
  const bt_lap_ws& W = a.W;
  const int kb = a.kb, list = a.list, n = a.n, m = a.m;
  const double thresh = a.thresh;
  int32_t* x = a.x; int32_t* y = a.y;
  const int32_t* row_block = a.row_block; const int32_t* col_block = a.col_block;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int GT = kLapThreads, NW = kLapThreads / 32;
#undef LAP_T
#define LAP_T(i) do { if (a.debug && tid == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(sm.tq[i])); } while (0)
    int32_t* segcnt_all = cand.cnt + (size_t)list * cand.rows_cap * cand.nseg;
  int32_t* ecol = cand.col + (size_t)list * cand.rows_cap * cand.stride;
  double* ecost = cand.cost + (size_t)list * cand.rows_cap * cand.stride;
  int32_t* indeg = cand.indeg + (size_t)list * cand.cols_cap;
  // ---- P1: classify every row from the emitters' degree bookkeeping (no edge traversal for the bulk):
  //      isolated edges (row degree 1, column in-degree 1, gate decision not in doubt) are final; rows
  //      with several candidates, a contested column or an ambiguous gate are "complex".  The phase is a chain
  //      of dependent global round trips (degree -> the row's column -> that column's in-degree), so every
  //      thread keeps four rows in flight and issues each level's loads for all of them together. ----
  // One SM classifies the whole stream, and the loop is bound by instruction issue: it only does what every row
  // needs (two loads, one dependent load, a few coalesced stores); complex rows are merely noted here and set
  // up by the pass below.
  constexpr int kP1 = 4;
  LAP_T(6);
  const size_t rbase = (size_t)list * cand.rows_cap;
  for (int r0 = tid; r0 < n; r0 += kP1 * GT) {
    int degs[kP1], rcs[kP1], inds[kP1], rbl[kP1];
    unsigned long long masks[kP1];      // the row's non-empty segments: fetched with the first level (no dependency)
#pragma unroll
    for (int q = 0; q < kP1; ++q) {
      const int r = r0 + q * GT;
      degs[q] = 0; rcs[q] = 0; rbl[q] = -1; masks[q] = 0ull;
      if (r < n) {
        degs[q] = cand.rowdeg[rbase + r]; rcs[q] = cand.rowcol[rbase + r];
        if (a.clear_cnt) masks[q] = cand.segmask[rbase + r];
        if (row_block) rbl[q] = row_block[r];
      }
    }
    if (a.debug && tid == 0) { asm volatile("" :: "r"(degs[0]), "r"(rcs[0]) : "memory"); asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(sm.tq[7])); }
#pragma unroll
    for (int q = 0; q < kP1; ++q) {
      const int c1 = rcs[q] & BT_EDGE_COLMASK;
      inds[q] = (degs[q] == 1) ? indeg[c1] : 0;
      if (degs[q] == 1 && col_block && col_block[c1] >= 0) inds[q] = -1;     // its only column is taken already
    }
#pragma unroll
    for (int q = 0; q < kP1; ++q) {
      const int r = r0 + q * GT, deg = degs[q], rc = rcs[q];
      if (deg == 0) continue;                      // nothing was emitted for this row
      if (a.clear) cand.rowdeg[rbase + r] = 0;
      const bool dead = rbl[q] >= 0 || (deg == 1 && inds[q] < 0);             // row matched by stage 1 / its only column taken
      const bool simple = deg == 1 && inds[q] == 1 && !(rc & BT_EDGE_AMBIG);   // isolated edge
      if (dead || simple) {
        if (simple && !dead) {
          const int col1 = rc & BT_EDGE_COLMASK;
          x[r] = col1;
          if (a.need_y) y[col1] = r;               // (a scattered store per row: only when somebody reads y)
        }
        if (a.clear_cnt) {                         // leave the segment counters zeroed
          unsigned long long mm = masks[q];
          while (mm) { const int g = __ffsll((long long)mm) - 1; mm &= mm - 1; segcnt_all[(size_t)r * cand.nseg + g] = 0; }
        }
        if (a.clear) cand.segmask[rbase + r] = 0ull;
      } else {
        const int kc = atomicAdd(&sm.nC, 1);
        W.clist[kc] = r;
        W.isroot[kc] = deg;
      }
    }
  }
  __syncthreads();
  // ---- P1b: set up the complex rows: their segment masks, their slice of the on-chip edge arrays, one
  //      (row, segment) work item per non-empty segment for the gather ----
  {
    const int nNoted = sm.nC;
    for (int kc = tid; kc < nNoted; kc += GT) {
      const int r = W.clist[kc], deg = W.isroot[kc];
      const unsigned long long mask = cand.segmask[rbase + r];
      if (a.clear) cand.segmask[rbase + r] = 0ull;
      W.rmaskg[r] = mask;                       // (the large-problem path walks the row's segments from here)
      const int e0 = atomicAdd(&sm.nE, deg);
      if (kc < kSmallRows && e0 + deg <= kSmallEdges && r < 4096) {
        sm.rg[kc] = r; sm.rdeg[kc] = deg; sm.rstart[kc] = e0;
        sm.fill[kc] = 0; sm.compidx[kc] = 0;
        unsigned long long mm = mask;
        while (mm) {
          const int g = __ffsll((long long)mm) - 1; mm &= mm - 1;
          const int pi = atomicAdd(&sm.nP, 1);
          if (pi < kSmallPairs) sm.pairs[pi] = (kc << 6) | g; else sm.big = 1;
        }
      } else {
        sm.big = 1;
      }
    }
  }
  if (a.debug && tid == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(sm.tq[0]));
  __syncthreads();
  LAP_T(2);
  const int nC = sm.nC;
  if (a.clear)                              // every read of the in-degrees is behind us
    for (int c = tid; c < cand.cols_cap; c += GT) indeg[c] = 0;
  const bool big = sm.big != 0 || m > kSmallCols;
  int dbg_ncomp = 0;
  (void)dbg_ncomp;

  if (nC > 0 && !big) {
    // per-column solver state (global column ids)
    for (int c = tid; c < m; c += GT) { sm.v[c] = 0.0; sm.seen[c] = 0; sm.insc[c] = 0; sm.yl[c] = -1; sm.clabel[c] = kInf; }
    // ---- S2: gather: one thread per non-empty (row, segment) pair -- two dependent round trips for the whole
    //      complex part (segment count -> its edges), whatever the number of rows ----
    const int nP = sm.nP;
    for (int q = tid; q < nP; q += GT) {
      const int i = sm.pairs[q] >> 6, g = sm.pairs[q] & 63;
      const int r = sm.rg[i];
      const size_t ri = (size_t)list * cand.rows_cap + r;
      int32_t* segcnt = cand.cnt + ri * cand.nseg;
      const int kk = segcnt[g];
      if (a.clear_cnt) segcnt[g] = 0;
      const int32_t* rc = ecol + (size_t)r * cand.stride + (size_t)g * cand.seg;
      const double* rv = ecost + (size_t)r * cand.stride + (size_t)g * cand.seg;
      const int dst0 = sm.rstart[i] + atomicAdd(&sm.fill[i], kk);
      for (int j = 0; j < kk; ++j) {
        const int colf = rc[j];
        const int col = colf & BT_EDGE_COLMASK;
        const bool ok = edge_ok(col_block, col);
        sm.ecol[dst0 + j] = col;
        sm.ecst[dst0 + j] = ok ? rv[j] : kBlockedCost;
        if (ok && (colf & BT_EDGE_AMBIG)) sm.compidx[i] = 1;       // the row's gate decision is in doubt
        if (ok && rf.enabled && (colf & (BT_EDGE_SIM | BT_EDGE_AMBIG))) sm.rlist[atomicAdd(&sm.nR, 1)] = (r << 12) | (dst0 + j);
      }
    }
    if (tid == 0) sm.changed = 0;
    __syncthreads();
    // the row's cheapest and second cheapest usable edge (provisional costs)
    for (int i = tid; i < nC; i += GT) {
      double b1 = kBlockedCost, b2 = kBlockedCost;
      int bc = kInf;
      const int e0 = sm.rstart[i], deg = sm.rdeg[i];
      for (int j = 0; j < deg; ++j) {
        const double cst = sm.ecst[e0 + j];
        const int col = sm.ecol[e0 + j];
        if (cst < b1 || (cst == b1 && col < bc)) { b2 = b1; b1 = cst; bc = col; }
        else if (cst < b2) b2 = cst;
      }
      // decided without looking at anybody else: no gate in doubt, and the row's choice cannot change within
      // the error of a tensor-core similarity (its runner-up is out of reach or clearly worse)
      const bool decided = sm.compidx[i] == 0 && (b1 >= thresh || b2 >= thresh || b2 - b1 > kTieGap);
      const int c = (b1 < thresh) ? bc : -1;
      sm.xl[i] = c;
      // ---- S2b: every complex row takes its cheapest edge -- if those choices are decided and pairwise
      //      distinct they are the optimum (each row at its own lower bound), which is the usual frame ----
      bool clash = !decided;
      if (c >= 0 && atomicCAS(&sm.yl[c], -1, i) != -1) clash = true;
      if (clash) sm.changed = 1;
    }
    __syncthreads();
    LAP_T(5);
    const bool greedy_ok = sm.changed == 0;
    __syncthreads();
    if (greedy_ok) {
      for (int i = tid; i < nC; i += GT) {
        const int c = sm.xl[i];
        if (c >= 0) { x[sm.rg[i]] = c; if (a.need_y) y[c] = sm.rg[i]; }
      }
      dbg_ncomp = -1;
      LAP_T(3);
    } else {
    for (int i = tid; i < nC; i += GT) {
      const int c = sm.xl[i];
      if (c >= 0) sm.yl[c] = -1;
    }
    __syncthreads();
    // ---- S1: ascending row order (deterministic tie-breaking of the general solver), by rank ----
    {
      int my_r = 0, my_deg = 0, my_start = 0, rank = 0;
      if (tid < nC) {
        my_r = sm.rg[tid]; my_deg = sm.rdeg[tid]; my_start = sm.rstart[tid];
        for (int j = 0; j < nC; ++j) rank += (sm.rg[j] < my_r) ? 1 : 0;
      }
      __syncthreads();
      if (tid < nC) { sm.rg[rank] = my_r; sm.rdeg[rank] = my_deg; sm.rstart[rank] = my_start; }
      __syncthreads();
    }
    // ---- S3: exact re-costing of the flagged edges, one warp each ----
    {
      const int nR = sm.nR;
      for (int q = warp; q < nR; q += NW) {
        const int r = sm.rlist[q] >> 12, e = sm.rlist[q] & 4095;
        const double c = refine_cost(rf, B, kb, list, r, sm.ecol[e], lane);
        if (lane == 0) sm.ecst[e] = c;
      }
    }
    for (int i = tid; i < nC; i += GT) { sm.xl[i] = -1; sm.u[i] = 0.0; sm.rlabel[i] = i; }
    __syncthreads();
    LAP_T(3);
    const SolveArrays A{sm.rowcnt, sm.colcnt, sm.sorted_rows, sm.touched, sm.treerows, sm.u, sm.v, sm.dist,
                        sm.pathrow, sm.seen, sm.insc, sm.rstart, 0};
    if (nC <= kOneWarpRows) {
      // ---- a handful of rows: one warp runs the shortest-augmenting-path solver over all of them ----
      if (warp == 0) {
        if (lane < nC) sm.sorted_rows[lane] = lane;
        if (lane == 0) { sm.rowcnt[0] = 0; sm.rowcnt[1] = nC; sm.colcnt[0] = 0; sm.colcnt[1] = 0; }
        __syncwarp();
        solve_component(A, 0, thresh, nullptr, sm.rdeg, sm.ecol, sm.ecst, sm.xl, sm.yl, lane);
      }
      dbg_ncomp = 1;
    } else {
      // ---- S4: components by min-label propagation + pointer jumping (shared memory) ----
      while (true) {
        if (tid == 0) sm.changed = 0;
        __syncthreads();
        bool changed = false;
        for (int i = tid; i < nC; i += GT) {
          int lr = sm.rlabel[i];
          const int l0 = lr;
          for (int q = sm.rstart[i]; q < sm.rstart[i] + sm.rdeg[i]; ++q) {
            const int c = sm.ecol[q];
            const int lc = sm.clabel[c];
            if (lc < lr) lr = lc;
            else if (lc > lr) { atomicMin(&sm.clabel[c], lr); changed = true; }
          }
          if (lr < l0) { atomicMin(&sm.rlabel[i], lr); changed = true; }
        }
        __syncthreads();
        for (int i = tid; i < nC; i += GT) {
          const int l = sm.rlabel[i];
          const int ll = sm.rlabel[l];
          if (ll < l) { atomicMin(&sm.rlabel[i], ll); changed = true; }
        }
        if (changed) sm.changed = 1;
        __syncthreads();
        const int any = sm.changed;
        __syncthreads();
        if (!any) break;
      }
      // ---- S5: grouping ----
      for (int i = tid; i < nC; i += GT) sm.isroot[i] = (sm.rlabel[i] == i) ? 1 : 0;
      for (int i = tid; i <= nC; i += GT) { sm.rowcnt[i] = 0; sm.colcnt[i] = 0; sm.fill[i] = 0; }
      __syncthreads();
      const int ncomp = block_exclusive_scan(sm.isroot, nC, sm.s_warp, &sm.s_carry);
      for (int i = tid; i < nC; i += GT)
        if (sm.rlabel[i] == i) sm.compidx[i] = sm.isroot[i];
      __syncthreads();
      for (int i = tid; i < nC; i += GT) atomicAdd(&sm.rowcnt[sm.compidx[sm.rlabel[i]]], 1);
      for (int c = tid; c < m; c += GT)
        if (sm.clabel[c] != kInf) atomicAdd(&sm.colcnt[sm.compidx[sm.clabel[c]]], 1);
      __syncthreads();
      block_exclusive_scan(sm.rowcnt, ncomp + 1, sm.s_warp, &sm.s_carry);
      block_exclusive_scan(sm.colcnt, ncomp + 1, sm.s_warp, &sm.s_carry);
      for (int i = tid; i < nC; i += GT) {
        const int q = sm.compidx[sm.rlabel[i]];
        sm.sorted_rows[sm.rowcnt[q] + atomicAdd(&sm.fill[q], 1)] = i;
      }
      __syncthreads();
      // ---- S6: tiny components one thread each, the others one warp each ----
      constexpr int kSerialRows = 6;
      for (int comp = tid; comp < ncomp; comp += GT) {
        const int nr = sm.rowcnt[comp + 1] - sm.rowcnt[comp];
        if (nr <= 3) solve_component_enum(A, comp, thresh, sm.rdeg, sm.ecol, sm.ecst, sm.xl, sm.yl);
        else if (nr <= kSerialRows) solve_component_serial(A, comp, thresh, sm.rdeg, sm.ecol, sm.ecst, sm.xl, sm.yl);
      }
      __syncwarp();
      for (int comp = warp; comp < ncomp; comp += NW)
        if (sm.rowcnt[comp + 1] - sm.rowcnt[comp] > kSerialRows)
          solve_component(A, comp, thresh, nullptr, sm.rdeg, sm.ecol, sm.ecst, sm.xl, sm.yl, lane);
      dbg_ncomp = ncomp;
    }
    __syncthreads();
    // ---- S7: write back ----
    for (int i = tid; i < nC; i += GT) {
      const int c = sm.xl[i];
      if (c >= 0) { x[sm.rg[i]] = c; if (a.need_y) y[c] = sm.rg[i]; }
    }
    }   // general (non-greedy) path
  } else if (nC > 0) {
    // ---- large complex part (a crowded scene, a dense cost matrix, more detections than the on-chip
    //      arrays hold): same algorithm over global scratch ----
    int32_t* cnt = cand.deg + (size_t)list * cand.rows_cap;
    for (int c = tid; c < m; c += GT) { W.collabel[c] = kInf; W.v[c] = 0.0; W.seen[c] = 0; W.insc[c] = 0; }
    for (int r = tid; r < n; r += GT) { W.u[r] = 0.0; W.label[r] = kInf; }
    if (tid < 8) W.counters[tid] = 0;
    __syncthreads();
    // compact the complex rows' segments in place (ascending copy), one thread per row
    for (int i = tid; i < nC; i += GT) {
      const int r = W.clist[i];
      const size_t ri = (size_t)list * cand.rows_cap + r;
      const unsigned long long mask = W.rmaskg[r];
      int32_t* segcnt = cand.cnt + ri * cand.nseg;
      int32_t* rc = ecol + (size_t)r * cand.stride;
      double* rv = ecost + (size_t)r * cand.stride;
      int total = 0;
      for (int g = 0; g < cand.nseg; ++g) {
        if (!((mask >> g) & 1ull)) continue;
        const int kk = segcnt[g];
        if (kk == 0) continue;
        if (a.clear_cnt) segcnt[g] = 0;
        const int src = g * cand.seg;
        for (int e = 0; e < kk; ++e) {
          if (src + e != total) { rc[total] = rc[src + e]; rv[total] = rv[src + e]; }
          ++total;
        }
      }
      cnt[r] = total;
      W.label[r] = r;
    }
    __syncthreads();
    // exact re-costing of the flagged edges: one warp per row, flagged edges one after the other
    for (int i = warp; i < nC; i += NW) {
      const int r = W.clist[i];
      const int deg = cnt[r];
      int32_t* rc = ecol + (size_t)r * cand.stride;
      double* rv = ecost + (size_t)r * cand.stride;
      for (int q0 = 0; q0 < deg; q0 += 32) {
        const int q = q0 + lane;
        const int colf = q < deg ? rc[q] : 0;
        const bool ok = q < deg && edge_ok(col_block, colf & BT_EDGE_COLMASK);
        unsigned todo = __ballot_sync(0xffffffffu, ok && rf.enabled && (colf & (BT_EDGE_SIM | BT_EDGE_AMBIG)));
        while (todo) {
          const int bsrc = __ffs(todo) - 1;
          todo &= todo - 1;
          const int col = __shfl_sync(0xffffffffu, colf, bsrc) & BT_EDGE_COLMASK;
          const double c = refine_cost(rf, B, kb, list, r, col, lane);
          if (lane == bsrc) rv[q] = c;
        }
        if (q < deg) rc[q] = colf & BT_EDGE_COLMASK;
      }
    }
    __syncthreads();
    // components of the complex part
    for (int iter = 0;; ++iter) {
      int32_t* flag = &W.counters[2 + (iter & 1)];
      bool changed = false;
      for (int i = tid; i < nC; i += GT) {
        const int r = W.clist[i];
        int lr = W.label[r];
        const int l0 = lr;
        const int deg = cnt[r];
        const int32_t* e = ecol + (size_t)r * cand.stride;
        for (int q = 0; q < deg; ++q) {
          const int c = e[q];
          if (!edge_ok(col_block, c)) continue;
          const int lc = W.collabel[c];
          if (lc < lr) lr = lc;
          else if (lc > lr) { atomicMin(&W.collabel[c], lr); changed = true; }
        }
        if (lr < l0) { atomicMin(&W.label[r], lr); changed = true; }
      }
      __syncthreads();
      for (int i = tid; i < nC; i += GT) {
        const int r = W.clist[i];
        const int l = W.label[r];
        const int ll = W.label[l];
        if (ll < l) { atomicMin(&W.label[r], ll); changed = true; }
      }
      if (changed) *flag = 1;
      if (tid == 0) W.counters[2 + ((iter + 1) & 1)] = 0;   // the other flag, for the next round
      __syncthreads();
      if (*flag == 0) break;
    }
    // component bookkeeping
    for (int i = tid; i < nC; i += GT) W.isroot[i] = (W.label[W.clist[i]] == W.clist[i]) ? 1 : 0;
    for (int i = tid; i <= nC; i += GT) { W.rowcnt[i] = 0; W.colcnt[i] = 0; W.fill[i] = 0; }
    __syncthreads();
    const int ncomp = block_exclusive_scan(W.isroot, nC, sm.s_warp, &sm.s_carry);
    for (int i = tid; i < nC; i += GT) {
      const int r = W.clist[i];
      if (W.label[r] == r) W.compidx[r] = W.isroot[i];
    }
    __syncthreads();
    for (int i = tid; i < nC; i += GT) atomicAdd(&W.rowcnt[W.compidx[W.label[W.clist[i]]]], 1);
    for (int c = tid; c < m; c += GT) {
      const int l = W.collabel[c];
      if (l != kInf) atomicAdd(&W.colcnt[W.compidx[l]], 1);
    }
    __syncthreads();
    block_exclusive_scan(W.rowcnt, ncomp + 1, sm.s_warp, &sm.s_carry);
    block_exclusive_scan(W.colcnt, ncomp + 1, sm.s_warp, &sm.s_carry);
    for (int i = tid; i < nC; i += GT) {
      const int r = W.clist[i];
      const int q = W.compidx[W.label[r]];
      W.sorted_rows[W.rowcnt[q] + atomicAdd(&W.fill[q], 1)] = r;
    }
    __syncthreads();
    const SolveArrays GA{W.rowcnt, W.colcnt, W.sorted_rows, W.touched, W.treerows, W.u, W.v, W.dist,
                         W.pathrow, W.seen, W.insc, nullptr, (size_t)cand.stride};
    // large components: the whole CTA, one after the other (per-column state in shared memory when it fits) ...
    constexpr int kCtaRows = 48;
    bool any_cta = false;
    for (int comp = 0; comp < ncomp; ++comp) any_cta = any_cta || (W.rowcnt[comp + 1] - W.rowcnt[comp] > kCtaRows);
    if (any_cta) {                                  // CTA-uniform
      ColState cs{W.v, W.dist, W.pathrow, W.seen, W.insc, y};
      if (m <= kSmallCols) {
        cs = ColState{sm.v, sm.dist, sm.pathrow, sm.seen, sm.insc, sm.yl};
        for (int c = tid; c < m; c += GT) { sm.v[c] = 0.0; sm.seen[c] = 0; sm.insc[c] = 0; sm.yl[c] = -1; }
      }
      __syncthreads();
      for (int comp = 0; comp < ncomp; ++comp)
        if (W.rowcnt[comp + 1] - W.rowcnt[comp] > kCtaRows)
          solve_component_cta(GA, comp, thresh, m, col_block, cnt, ecol, ecost, x, y, cs, sm);
      __syncthreads();
    }
    // ... the others one warp each
    for (int comp = warp; comp < ncomp; comp += NW)
      if (W.rowcnt[comp + 1] - W.rowcnt[comp] <= kCtaRows)
        solve_component(GA, comp, thresh, col_block, cnt, ecol, ecost, x, y, lane);
    dbg_ncomp = ncomp;
  }
  (void)kb;
}

// One CTA per video stream solves the (up to three chained) association stages of its frame.
__global__ void __launch_bounds__(kLapThreads, 1)
lap_stream_kernel(bt_cand cand_base, bt_lap_ws ws_base, const bt_lap_batch* __restrict__ Bp, LapParams P,
                  const __grid_constant__ bt_refine rf) {
  const bt_lap_batch& B = *Bp;     // the frame's batch description, in device memory (same kernel arguments every frame)
  extern __shared__ __align__(16) unsigned char lap_dyn_smem[];
  SmallSmem& sm = *reinterpret_cast<SmallSmem*>(lap_dyn_smem);
  const int kb = blockIdx.x;
  const int sid = B.sid[kb];
  const int n = B.n[kb], m = B.m[kb];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int GT = kLapThreads, NW = kLapThreads / 32;
  const bt_cand cand = bt_cand_of(cand_base, sid);
  bt_lap_ws W = ws_base;
  {
    const size_t ro = (size_t)sid * ws_base.rows_stride, co = (size_t)sid * ws_base.cols_stride;
    W.label += ro; W.clist += ro; W.rmaskg += ro; W.compidx += ro; W.isroot += ro; W.rowcnt += ro + sid; W.colcnt += ro + sid;
    W.fill += ro + sid; W.sorted_rows += ro; W.u += ro; W.treerows += ro;
    W.collabel += co; W.v += co; W.dist += co; W.pathrow += co; W.seen += co; W.insc += co; W.touched += co;
    W.counters += (size_t)sid * 8;
  }
  unsigned long long* tq = sm.tq;
#undef LAP_T
#define LAP_T(i) do { if (P.debug && tid == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tq[i])); } while (0)
  LAP_T(0);

  // ---- P0: outputs (this kernel's own) ----
  for (int s = 0; s < P.nstages; ++s) {
    int32_t* x = B.x[kb] + (size_t)s * B.x_stride[kb];
    int32_t* y = B.y[kb] + (size_t)s * B.y_stride;
    for (int r = tid; r < n; r += GT) x[r] = -1;
    for (int c = tid; c < m; c += GT) y[c] = -1;
  }
  if (tid == 0 && B.zero_word[kb]) *B.zero_word[kb] = 0;
  // Everything above touches only this kernel's own outputs, so under programmatic dependent launch it
  // runs while the emitting kernel drains; the candidate lists are read below.
  // ---- dry run on the ctx's two-row dummy lists (see lap_stage): instruction-cache warm-up under the wait ----
  if (P.nstages == 3) {
    StageArgs& wa = sm.sa;
    if (tid == 0) {
    wa.cand = ws_base.warm; wa.W = W; wa.kb = kb; wa.list = 0; wa.n = 2; wa.m = 3; wa.clear = 0; wa.clear_cnt = 0; wa.need_y = 1; wa.debug = 0;
    wa.thresh = 0.8; wa.x = ws_base.warm_xy + (size_t)sid * 8; wa.y = wa.x + 4; wa.row_block = nullptr; wa.col_block = nullptr; wa.tq = nullptr;
    sm.nC = 0; sm.nE = 0; sm.nR = 0; sm.nP = 0; sm.big = 0;
    }
    __syncthreads();
    lap_stage(B, rf, sm);
    __syncthreads();
  }
  // a stage's arguments (a few hundred bytes written by one thread); need_y is patched once the totals are known
  auto setup_stage = [&](const int stage) {
    StageArgs& sa = sm.sa;
    const int list = P.nstages == 1 ? P.list0 : stage;
    sa.cand = cand; sa.W = W; sa.kb = kb; sa.list = list; sa.n = n; sa.m = m; sa.clear = P.clear_lists; sa.debug = P.debug;
    sa.clear_cnt = P.clear_lists && P.clear_cnt;
    sa.need_y = 1;
    sa.thresh = P.thresh[stage];
    sa.x = B.x[kb] + (size_t)stage * B.x_stride[kb];
    sa.y = B.y[kb] + (size_t)stage * B.y_stride;
    // stage 2 (demo:1568-1571): rows unmatched in stage 1; stage 3 (demo:1588-1604): columns unmatched in stage 1
    sa.row_block = (P.nstages == 3 && stage == 1) ? B.x[kb] : nullptr;
    sa.col_block = (P.nstages == 3 && stage == 2) ? B.y[kb] : nullptr;
    sa.tq = nullptr;
    sm.nC = 0; sm.nE = 0; sm.nR = 0; sm.nP = 0; sm.big = 0;
  };
  if (tid == 0) setup_stage(0);      // the first stage's arguments are ready before the candidate lists are
  asm volatile("griddepcontrol.wait;" ::: "memory");
  __syncthreads();
  LAP_T(1);

  // which stages have any candidates at all: one round trip for the three counters
  const int tot0 = cand.total[0], tot1 = cand.total[1], tot2 = cand.total[2];
  for (int stage = 0; stage < P.nstages; ++stage) {
    const int list = P.nstages == 1 ? P.list0 : stage;
    if ((list == 0 ? tot0 : (list == 1 ? tot1 : tot2)) == 0) continue;   // nothing was emitted for this stage (CTA-uniform)
    if (tid == 0) {
      if (stage != 0) setup_stage(stage);
      // y (column -> row) of a stage is read by the caller of the stand-alone solver and, for the frame's first
      // stage, by stage 3 (columns stage 1 took are blocked) -- when stage 3 has any candidates at all
      sm.sa.need_y = (P.nstages == 1) || (stage == 0 && tot2 != 0);
    }
    __syncthreads();
    lap_stage(B, rf, sm);
    __syncthreads();                                  // x / y of this stage gate the next one
    if (P.clear_lists && tid == 0) cand.total[list] = 0;
    LAP_T(4);
    if (P.debug && tid == 0)
      printf("lap stream %d stage %d: n=%d m=%d complex=%d edges=%d flagged=%d | classify %llu [entry +%llu, level-1 loads +%llu, thread 0 done +%llu] gather %llu (to recost/greedy end %llu) solve %llu ns\n",
             sid, stage, n, m, sm.nC, sm.nE, sm.nR, tq[2] - tq[1], tq[6] - tq[1], tq[7] - tq[1], tq[0] - tq[1], tq[5] - tq[2], tq[3] - tq[2], tq[4] - tq[3]);
    LAP_T(1);
  }
  // ---- direct results: the assignment vectors go to the host by plain stores over PCIe, then the flag ----
  if (B.hx[kb]) {
    for (int s = 0; s < P.nstages; ++s) {
      const int32_t* x = B.x[kb] + (size_t)s * B.x_stride[kb];
      int32_t* hx = B.hx[kb] + (size_t)s * B.x_stride[kb];
      for (int r = tid; r < n; r += GT) hx[r] = x[r];
    }
    __threadfence_system();
    __syncthreads();
    if (tid == 0) *reinterpret_cast<volatile uint32_t*>(B.hflag[kb]) = 1u;
  }
}

// dense float64 cost -> candidate list (ordered by column): one warp per row
__global__ void __launch_bounds__(256)
lap_compact_dense_kernel(const double* __restrict__ cost, int n, int m, double thresh, bt_cand cand,
                         int list) {
  const int lane = threadIdx.x & 31;
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= n) return;
  int32_t* ecol = cand.col + ((size_t)list * cand.rows_cap + row) * cand.stride;
  double* ecost = cand.cost + ((size_t)list * cand.rows_cap + row) * cand.stride;
  int32_t* segcnt = cand.cnt + ((size_t)list * cand.rows_cap + row) * cand.nseg;
  int count = 0, rowtotal = 0;
  for (int c0 = 0; c0 < m; c0 += 32) {
    if (c0 > 0 && (c0 % cand.seg) == 0) {     // next segment (seg is a multiple of 32 here)
      if (lane == 0) segcnt[c0 / cand.seg - 1] = count;
      count = 0;
    }
    const int c = c0 + lane;
    const double v = (c < m) ? cost[(size_t)row * m + c] : DBL_MAX;
    const bool keep = (c < m) && (v < thresh);
    const unsigned ball = __ballot_sync(0xffffffffu, keep);
    if (keep) {
      const int pos = (c0 / cand.seg) * cand.seg + count + __popc(ball & ((1u << lane) - 1));
      ecol[pos] = c;
      ecost[pos] = v;
      atomicAdd(&cand.indeg[(size_t)list * cand.cols_cap + c], 1);
      cand.rowcol[(size_t)list * cand.rows_cap + row] = c;
    }
    count += __popc(ball);
    rowtotal += __popc(ball);
  }
  if (lane == 0) cand.rowdeg[(size_t)list * cand.rows_cap + row] = rowtotal;
  if (lane == 0) segcnt[(m - 1) / cand.seg] = count;
  if (lane == 0) {
    atomicAdd(&cand.total[list], 1);   // any non-zero value means "stage not empty"
    const int nseg_used = (m - 1) / cand.seg + 1;
    cand.segmask[(size_t)list * cand.rows_cap + row] = (nseg_used >= 64) ? ~0ull : ((1ull << nseg_used) - 1);
  }
}


}  // namespace

int32_t bt_lap_ws_create(bt_ctx* ctx) {
  auto* ws = new bt_lap_ws();
  ctx->lap = ws;
  const int rows = ctx->max_tracks, cols = ctx->max_dets, S = ctx->n_streams;
  ws->rows = rows;
  ws->cols = cols;
  ws->rows_stride = rows;
  ws->cols_stride = cols;
  const int stride = (cols + 127) / 128 * 128 + 128;
  bt_cand& c = ws->cand;
  c.rows_cap = rows;
  c.stride = stride;
  c.nseg = BT_CAND_MAXSEG;
  c.seg = 128;
  c.cols_cap = cols;
  BT_CHECK(cols <= BT_CAND_MAXSEG * 112, BT_ERR_CAPACITY, "max_dets %d exceeds %d", cols, BT_CAND_MAXSEG * 112);
  // one "clear block" per video stream: cnt | total | segmask | rowdeg | indeg (one memset clears them all)
  const size_t cnt_ints = (3 * (size_t)rows * c.nseg + 4 + 1) & ~size_t(1);   // keeps segmask 8 B aligned
  c.clear_bytes = sizeof(int32_t) * cnt_ints + sizeof(unsigned long long) * 3 * rows +
                  sizeof(int32_t) * 3 * ((size_t)rows + cols);
  c.s_cnt = (c.clear_bytes + 255) & ~size_t(255);
  c.s_rowcol = 3 * (size_t)rows;
  c.s_deg = 3 * (size_t)rows;
  c.s_edges = 3 * (size_t)rows * stride;
  char* blk = nullptr;
  BT_CUDA(cudaMalloc(&blk, c.s_cnt * S));
  BT_CUDA(cudaMemset(blk, 0, c.s_cnt * S));
  c.cnt = reinterpret_cast<int32_t*>(blk);
  c.total = c.cnt + 3 * (size_t)rows * c.nseg;
  c.segmask = reinterpret_cast<unsigned long long*>(c.cnt + cnt_ints);
  c.rowdeg = reinterpret_cast<int32_t*>(c.segmask + 3 * (size_t)rows);
  c.indeg = c.rowdeg + 3 * (size_t)rows;
  BT_CUDA(cudaMalloc(&c.rowcol, sizeof(int32_t) * c.s_rowcol * S));
  BT_CUDA(cudaMalloc(&c.deg, sizeof(int32_t) * c.s_deg * S));
  BT_CUDA(cudaMalloc(&c.col, sizeof(int32_t) * c.s_edges * S));
  BT_CUDA(cudaMalloc(&c.cost, sizeof(double) * c.s_edges * S));
#define BT_LAP_ALLOC(field, type, count) BT_CUDA(cudaMalloc(&ws->field, sizeof(type) * (size_t)(count) * S))
  BT_LAP_ALLOC(label, int32_t, rows);
  BT_LAP_ALLOC(collabel, int32_t, cols);
  BT_LAP_ALLOC(clist, int32_t, rows);
  BT_LAP_ALLOC(rmaskg, unsigned long long, rows);
  BT_LAP_ALLOC(compidx, int32_t, rows);
  BT_LAP_ALLOC(isroot, int32_t, rows);
  BT_LAP_ALLOC(rowcnt, int32_t, rows + 1);
  BT_LAP_ALLOC(colcnt, int32_t, rows + 1);
  BT_LAP_ALLOC(fill, int32_t, rows + 1);
  BT_LAP_ALLOC(sorted_rows, int32_t, rows);
  BT_LAP_ALLOC(counters, int32_t, 8);
  BT_LAP_ALLOC(u, double, rows);
  BT_LAP_ALLOC(v, double, cols);
  BT_LAP_ALLOC(dist, double, cols);
  BT_LAP_ALLOC(pathrow, int32_t, cols);
  BT_LAP_ALLOC(seen, int32_t, cols);
  BT_LAP_ALLOC(insc, int32_t, cols);
  BT_LAP_ALLOC(touched, int32_t, cols);
  BT_LAP_ALLOC(treerows, int32_t, rows);
#undef BT_LAP_ALLOC
  BT_CUDA(cudaMalloc(&ws->x, sizeof(int32_t) * rows));
  BT_CUDA(cudaMalloc(&ws->y, sizeof(int32_t) * cols));
  {
    // dummy lists: rows 0 and 1 with two candidates each (row 0: columns 0, 1; row 1: columns 1, 2), so both are
    // "complex" rows whose cheapest edges are distinct: classification, gather and the greedy check all run
    bt_cand& w = ws->warm;
    w.rows_cap = 4; w.cols_cap = 4; w.nseg = BT_CAND_MAXSEG; w.seg = 128; w.stride = 256;
    const size_t n_cnt = 3 * 4 * BT_CAND_MAXSEG, n_edges = 3 * 4 * 256;
    const size_t bytes = sizeof(int32_t) * (n_cnt + 4 + 12 + 12 + 12 + 12) + sizeof(unsigned long long) * 12 +
                         sizeof(int32_t) * n_edges + sizeof(double) * n_edges + 64;
    std::vector<char> h(bytes, 0);
    BT_CUDA(cudaMalloc(&ws->warm_blk, bytes));
    size_t off = 0;
    auto carve = [&](size_t nbytes, size_t align) { off = (off + align - 1) & ~(align - 1); const size_t o = off; off += nbytes; return o; };
    const size_t o_cost = carve(sizeof(double) * n_edges, 8), o_mask = carve(sizeof(unsigned long long) * 12, 8);
    const size_t o_cnt = carve(sizeof(int32_t) * n_cnt, 4), o_total = carve(16, 4), o_rowdeg = carve(48, 4), o_indeg = carve(48, 4),
                 o_rowcol = carve(48, 4), o_deg = carve(48, 4), o_col = carve(sizeof(int32_t) * n_edges, 4);
    double* hc = reinterpret_cast<double*>(h.data() + o_cost);
    int32_t* hcol = reinterpret_cast<int32_t*>(h.data() + o_col);
    int32_t* hcnt = reinterpret_cast<int32_t*>(h.data() + o_cnt);
    hcol[0 * 256 + 0] = 0; hc[0 * 256 + 0] = 0.1; hcol[0 * 256 + 1] = 1; hc[0 * 256 + 1] = 0.5;
    hcol[1 * 256 + 0] = 1; hc[1 * 256 + 0] = 0.1; hcol[1 * 256 + 1] = 2; hc[1 * 256 + 1] = 0.6;
    hcnt[0 * BT_CAND_MAXSEG] = 2; hcnt[1 * BT_CAND_MAXSEG] = 2;
    reinterpret_cast<unsigned long long*>(h.data() + o_mask)[0] = 1ull;
    reinterpret_cast<unsigned long long*>(h.data() + o_mask)[1] = 1ull;
    reinterpret_cast<int32_t*>(h.data() + o_total)[0] = 4;
    reinterpret_cast<int32_t*>(h.data() + o_rowdeg)[0] = 2; reinterpret_cast<int32_t*>(h.data() + o_rowdeg)[1] = 2;
    int32_t* hin = reinterpret_cast<int32_t*>(h.data() + o_indeg);
    hin[0] = 1; hin[1] = 2; hin[2] = 1;
    BT_CUDA(cudaMemcpy(ws->warm_blk, h.data(), bytes, cudaMemcpyHostToDevice));
    char* d = ws->warm_blk;
    w.cost = reinterpret_cast<double*>(d + o_cost);
    w.segmask = reinterpret_cast<unsigned long long*>(d + o_mask);
    w.cnt = reinterpret_cast<int32_t*>(d + o_cnt);
    w.total = reinterpret_cast<int32_t*>(d + o_total);
    w.rowdeg = reinterpret_cast<int32_t*>(d + o_rowdeg);
    w.indeg = reinterpret_cast<int32_t*>(d + o_indeg);
    w.rowcol = reinterpret_cast<int32_t*>(d + o_rowcol);
    w.deg = reinterpret_cast<int32_t*>(d + o_deg);
    w.col = reinterpret_cast<int32_t*>(d + o_col);
    BT_CUDA(cudaMalloc(&ws->warm_xy, sizeof(int32_t) * 8 * (size_t)S));
  }
  return BT_OK;
}

void bt_lap_ws_destroy(bt_ctx* ctx) {
  bt_lap_ws* ws = ctx->lap;
  if (!ws) return;
  void* ptrs[] = {ws->cand.cnt, ws->cand.rowcol, ws->cand.deg, ws->cand.col, ws->cand.cost, ws->label, ws->collabel,
                  ws->clist, ws->rmaskg, ws->compidx, ws->isroot, ws->rowcnt, ws->colcnt, ws->fill,
                  ws->sorted_rows, ws->counters, ws->u, ws->v, ws->dist, ws->pathrow, ws->seen, ws->insc,
                  ws->touched, ws->treerows, ws->x, ws->y, ws->warm_blk, ws->warm_xy};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  delete ws;
  ctx->lap = nullptr;
}

const bt_cand* bt_lap_own_cand(bt_ctx* ctx) { return &ctx->lap->cand; }

int32_t btk_lap_compact_dense(bt_ctx* ctx, const double* cost, int32_t n, int32_t m, double thresh,
                              const bt_cand& cand, int32_t list) {
  if (n <= 0 || m <= 0) return BT_OK;
  lap_compact_dense_kernel<<<(n * 32 + 255) / 256, 256, 0, ctx->stream>>>(cost, n, m, thresh, cand, list);
  BT_LAUNCHED(ctx);
  return BT_OK;
}

static int32_t launch_lap(bt_ctx* ctx, const bt_cand& cand, const bt_lap_batch& B, const bt_lap_batch* dB,
                          const LapParams& P, const bt_refine& rf) {
  for (int k = 0; k < B.count; ++k)
    BT_CHECK(B.n[k] <= ctx->lap->rows && B.m[k] <= ctx->lap->cols, BT_ERR_CAPACITY,
             "linear assignment %d x %d exceeds ctx capacity %d x %d", B.n[k], B.m[k], ctx->lap->rows, ctx->lap->cols);
  if (B.count <= 0) return BT_OK;
  static std::once_flag attr_once[64];     // function attributes are per device: one flag per device ordinal
  cudaError_t attr_err = cudaSuccess;
  std::call_once(attr_once[ctx->device & 63], [&] {
    attr_err = cudaFuncSetAttribute(lap_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmallSmem));
  });
  BT_CUDA(attr_err);
  BT_CUDA(bt_launch(ctx, true, lap_stream_kernel, dim3(B.count), dim3(kLapThreads), sizeof(SmallSmem), cand, *ctx->lap, dB, P, rf));
  BT_LAUNCHED(ctx);
  return BT_OK;
}

int32_t btk_lap_solve(bt_ctx* ctx, const bt_cand& cand, int32_t list, int32_t n, int32_t m,
                      double thresh, int32_t* x, int32_t* y) {
  if (n <= 0 && m <= 0) return BT_OK;
  LapParams P = {};
  P.nstages = 1;
  P.list0 = list;
  P.thresh[0] = thresh;
  bt_lap_batch B = {};
  B.count = 1;
  B.n[0] = n; B.m[0] = m;
  B.x[0] = x; B.x_stride[0] = n; B.y[0] = y; B.y_stride = m;
  bt_refine rf = {};
  // stand-alone: the batch description goes up through the ctx's descriptor scratch (second half)
  bt_lap_batch* dB = reinterpret_cast<bt_lap_batch*>(ctx->d_desc + 8192);
  BT_CUDA(cudaMemcpyAsync(dB, &B, sizeof(B), cudaMemcpyHostToDevice, ctx->stream));
  return launch_lap(ctx, cand, B, dB, P, rf);
}

int32_t btk_lap_solve3(bt_ctx* ctx, const bt_cand& cand, const bt_lap_batch& B, const bt_lap_batch* dB,
                       const double thresh[3], const bt_refine& rf, int atomic_emitter) {
  LapParams P = {};
  P.clear_lists = 1;
  P.clear_cnt = atomic_emitter;
  P.nstages = 3;
  for (int s = 0; s < 3; ++s) P.thresh[s] = thresh[s];
  P.debug = getenv("BT_LAP_DEBUG") != nullptr;
  return launch_lap(ctx, cand, B, dB, P, rf);
}
