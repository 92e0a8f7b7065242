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
// works on the *candidate edge lists* the association epilogue emits:
// ONE kernel launch (`lap_cluster_kernel`) solves up to three chained association stages.  It runs
// as a single thread-block cluster (up to 8 CTAs x 1024 threads) so that the phases below can be
// separated by hardware cluster barriers instead of kernel boundaries:
//   P1  per row: compact the (row, column-segment) sub-lists, count valid edges, column in-degrees
//   P2  isolated edges (row degree 1, column in-degree 1) are matched on the spot -- the bulk of a
//       tracking scene; the remaining "complex" rows are collected
//   P3  connected components of the complex part: min-label propagation + pointer jumping
//   P4  component row lists / scratch slices by block scans (CTA 0)
//   P5  one warp per component: exact shortest-augmenting-path assignment (Jonker-Volgenant /
//       Crouse formulation, float64 duals) where every row owns a private zero-cost dummy column
//       ("stay unmatched"); lanes parallelise the edge relaxations and the minimum scans.
// Stage 2 is masked by stage 1's matched rows and stage 3 by stage 1's matched columns, so the
// three solves of a frame chain inside the launch.  A single giant component (dense adversarial
// cost matrix) is still solved exactly, just slower.
// Ties between equal-cost optima are broken by lowest index, lap's own tie-breaking is not
// reproducible without its sources: "bit-exact" is defined on inputs with a unique optimum.
#include "common.cuh"

#include <float.h>
#include <stdlib.h>

struct bt_lap_ws {
  bt_cand cand;            // ctx-wide candidate lists (3 lists)
  int32_t* label = nullptr;     // [rows]
  int32_t* collabel = nullptr;  // [cols]
  int32_t* nvalid = nullptr;    // [rows]  valid out-degree
  int32_t* onlycol = nullptr;   // [rows]  a valid column of the row (the only one when nvalid == 1)
  int32_t* clist = nullptr;     // [rows]  complex rows
  int32_t* compidx = nullptr;   // [rows]  component index of a root row (indexed by row)
  int32_t* isroot = nullptr;    // [rows]  scan scratch (indexed by clist position)
  int32_t* rowcnt = nullptr;    // [rows+1] per component -> exclusive scan = row_start
  int32_t* colcnt = nullptr;    // [rows+1] per component -> exclusive scan = col_start
  int32_t* fill = nullptr;      // [rows]
  int32_t* sorted_rows = nullptr;  // [rows]
  int32_t* counters = nullptr;     // [8]: 0 ncomplex, 1 ncomp, 2/3 changed flags
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
};

namespace {

constexpr int kLapThreads = 1024;
constexpr int kLapMaxCtas = 8;
constexpr int kInf = 0x7fffffff;

struct LapStage {
  int list;
  double thresh;
  const int32_t* row_block;  // edge valid only if row_block[r] < 0 (nullptr: all rows)
  const int32_t* col_block;  // edge valid only if col_block[c] < 0
  int32_t* x;
  int32_t* y;
};
struct LapParams {
  LapStage st[3];
  int nstages;
  int n, m;
  int clear_lists;      // tracker mode: leave cnt / segmask / total zeroed for the next frame
  int32_t* zero_word;   // optional device word to clear (the frame's duplicate-pair counter)
  int debug;   // BT_LAP_DEBUG=1: phase timestamps (ns) by device printf
};

__device__ __forceinline__ bool edge_ok(const int32_t* __restrict__ col_block, int c) {
  return col_block == nullptr || col_block[c] < 0;
}

__device__ __forceinline__ void cluster_barrier() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
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
      int w = s_warp[lane];
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
    if (tid == kLapThreads - 1) *s_carry = carry + s_warp[31];
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
      if (lane == 0 && best.c != kInf) { x[r] = best.c; y[best.c] = r; }
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

// ---- on-chip path for the usual case: a handful of complex rows -------------------------------
// CTA 0 pulls the complex rows' valid edges into shared memory (CSR, local row ids; a column's
// local id is the smallest local edge index that touches it), labels components, groups them and
// lets its 32 warps solve them -- all with shared-memory latencies instead of L2 round trips.
constexpr int kSmallRows = 512;
constexpr int kSmallEdges = 2048;

struct SmallSmem {
  int32_t rg[kSmallRows], rdeg[kSmallRows], rstart[kSmallRows + 1], rlabel[kSmallRows], xl[kSmallRows];
  int32_t sorted_rows[kSmallRows], treerows[kSmallRows], compidx[kSmallRows], isroot[kSmallRows];
  int32_t rowcnt[kSmallRows + 1], colcnt[kSmallRows + 1], fill[kSmallRows + 1];
  double u[kSmallRows];
  int32_t lcl[kSmallEdges];
  double ecst[kSmallEdges];
  int32_t cglob[kSmallEdges], clabel[kSmallEdges], pathrow[kSmallEdges], seen[kSmallEdges], insc[kSmallEdges];
  int32_t yl[kSmallEdges], touched[kSmallEdges];
  double v[kSmallEdges], dist[kSmallEdges];
  int changed;
  int scan_total;
};

constexpr int kSmallThreads = 128;   // the on-chip path runs on the first four warps of CTA 0
#define SMALL_SYNC() asm volatile("bar.sync 2, 128;" ::: "memory")

// exclusive scan of a (shared-memory) array of up to a few thousand ints by warp 0; all kSmallThreads call it
__device__ int small_scan(int32_t* data, int n, int* total_slot) {
  SMALL_SYNC();
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    int carry = 0;
    for (int base = 0; base < n; base += 32) {
      const int i = base + lane;
      const int val = (i < n) ? data[i] : 0;
      int incl = val;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      if (i < n) data[i] = carry + incl - val;
      carry += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) *total_slot = carry;
  }
  SMALL_SYNC();
  return *total_slot;
}

__device__ void small_complex_solve(const bt_cand& cand, const bt_lap_ws& W, const LapStage& S, int nC,
                                    const int32_t* cnt, const int32_t* ecol, const double* ecost,
                                    int32_t* x, int32_t* y, SmallSmem& sm, int debug) {
  unsigned long long ts[8] = {0,0,0,0,0,0,0,0};
#define ST(i) do { if (debug && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ts[i])); } while (0)
  ST(0);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // S1: valid degree per complex row -> CSR offsets
  for (int i = tid; i < nC; i += kSmallThreads) {
    const int r = W.clist[i];
    sm.rg[i] = r;
    const int deg = cnt[r];
    const int32_t* e = ecol + (size_t)r * cand.stride;
    int v = 0;
    for (int k = 0; k < deg; ++k) v += edge_ok(S.col_block, e[k]) ? 1 : 0;
    sm.rdeg[i] = v;
    sm.rstart[i] = v;
    sm.rlabel[i] = i;
    sm.xl[i] = -1;
    sm.u[i] = 0.0;
  }
  if (tid == 0) sm.rstart[nC] = 0;
  SMALL_SYNC();
  const int E = small_scan(sm.rstart, nC + 1, &sm.scan_total);
  ST(1);
  // S2: edges; column representative = smallest local edge index touching the column
  for (int i = tid; i < nC; i += kSmallThreads) {
    const int r = sm.rg[i];
    const int deg = cnt[r];
    const int32_t* e = ecol + (size_t)r * cand.stride;
    const double* w = ecost + (size_t)r * cand.stride;
    int pos = sm.rstart[i];
    for (int k = 0; k < deg; ++k) {
      const int c = e[k];
      if (!edge_ok(S.col_block, c)) continue;
      sm.lcl[pos] = c;               // global column for now
      sm.ecst[pos] = w[k];
      atomicMin(&W.collabel[c], pos);
      ++pos;
    }
  }
  SMALL_SYNC();
  for (int e = tid; e < E; e += kSmallThreads) {
    const int c = sm.lcl[e];
    const int rep = __ldcg(&W.collabel[c]);
    sm.lcl[e] = rep;
    if (rep == e) sm.cglob[e] = c;
    sm.clabel[e] = kInf;
    sm.v[e] = 0.0;
    sm.seen[e] = 0;
    sm.insc[e] = 0;
    sm.yl[e] = -1;
  }
  SMALL_SYNC();
  ST(2);
  // S4: components by min-label propagation + pointer jumping (shared memory)
  while (true) {
    if (tid == 0) sm.changed = 0;
    SMALL_SYNC();
    bool changed = false;
    for (int i = tid; i < nC; i += kSmallThreads) {
      int lr = sm.rlabel[i];
      const int l0 = lr;
      for (int k = sm.rstart[i]; k < sm.rstart[i] + sm.rdeg[i]; ++k) {
        const int c = sm.lcl[k];
        const int lc = sm.clabel[c];
        if (lc < lr) lr = lc;
        else if (lc > lr) { atomicMin(&sm.clabel[c], lr); changed = true; }
      }
      if (lr < l0) { atomicMin(&sm.rlabel[i], lr); changed = true; }
    }
    SMALL_SYNC();
    for (int i = tid; i < nC; i += kSmallThreads) {
      const int l = sm.rlabel[i];
      const int ll = sm.rlabel[l];
      if (ll < l) { atomicMin(&sm.rlabel[i], ll); changed = true; }
    }
    if (changed) sm.changed = 1;
    SMALL_SYNC();
    const int any = sm.changed;
    SMALL_SYNC();
    if (!any) break;
  }
  ST(3);
  // S5: grouping
  for (int i = tid; i < nC; i += kSmallThreads) sm.isroot[i] = (sm.rlabel[i] == i) ? 1 : 0;
  for (int i = tid; i <= nC; i += kSmallThreads) { sm.rowcnt[i] = 0; sm.colcnt[i] = 0; sm.fill[i] = 0; }
  SMALL_SYNC();
  const int ncomp = small_scan(sm.isroot, nC, &sm.scan_total);
  for (int i = tid; i < nC; i += kSmallThreads)
    if (sm.rlabel[i] == i) sm.compidx[i] = sm.isroot[i];
  SMALL_SYNC();
  for (int i = tid; i < nC; i += kSmallThreads) atomicAdd(&sm.rowcnt[sm.compidx[sm.rlabel[i]]], 1);
  for (int e = tid; e < E; e += kSmallThreads)
    if (sm.lcl[e] == e) atomicAdd(&sm.colcnt[sm.compidx[sm.clabel[e]]], 1);
  SMALL_SYNC();
  small_scan(sm.rowcnt, ncomp + 1, &sm.scan_total);
  small_scan(sm.colcnt, ncomp + 1, &sm.scan_total);
  for (int i = tid; i < nC; i += kSmallThreads) {
    const int k = sm.compidx[sm.rlabel[i]];
    sm.sorted_rows[sm.rowcnt[k] + atomicAdd(&sm.fill[k], 1)] = i;
  }
  SMALL_SYNC();
  ST(4);
  // S6: one warp per component, everything in shared memory
  const SolveArrays A{sm.rowcnt, sm.colcnt, sm.sorted_rows, sm.touched, sm.treerows, sm.u, sm.v, sm.dist,
                      sm.pathrow, sm.seen, sm.insc, sm.rstart, 0};
  constexpr int kSerialRows = 6;     // components up to this many rows: one thread each
  for (int comp = tid; comp < ncomp; comp += kSmallThreads) {
    const int nr = sm.rowcnt[comp + 1] - sm.rowcnt[comp];
    if (nr <= 3) solve_component_enum(A, comp, S.thresh, sm.rdeg, sm.lcl, sm.ecst, sm.xl, sm.yl);
    else if (nr <= kSerialRows) solve_component_serial(A, comp, S.thresh, sm.rdeg, sm.lcl, sm.ecst, sm.xl, sm.yl);
  }
  __syncwarp();
  for (int comp = warp; comp < ncomp; comp += kSmallThreads / 32)
    if (sm.rowcnt[comp + 1] - sm.rowcnt[comp] > kSerialRows)
      solve_component(A, comp, S.thresh, nullptr, sm.rdeg, sm.lcl, sm.ecst, sm.xl, sm.yl, lane);
  SMALL_SYNC();
  ST(5);
  // S7: write back with global ids
  for (int i = tid; i < nC; i += kSmallThreads) {
    const int c = sm.xl[i];
    if (c >= 0) {
      x[sm.rg[i]] = sm.cglob[c];
      y[sm.cglob[c]] = sm.rg[i];
    }
  }
  ST(6);
  if (debug && threadIdx.x == 0)
    printf("small path nC=%d E=%d ncomp=%d: S1 %llu S2+3 %llu label %llu group %llu solve %llu write %llu ns\n", nC, E, ncomp,
           ts[1]-ts[0], ts[2]-ts[1], ts[3]-ts[2], ts[4]-ts[3], ts[5]-ts[4], ts[6]-ts[5]);
}

__global__ void __launch_bounds__(kLapThreads, 1)
lap_cluster_kernel(bt_cand cand, bt_lap_ws ws, LapParams P) {
  __shared__ int32_t s_warp[32];
  __shared__ int32_t s_carry;
  extern __shared__ __align__(16) unsigned char lap_dyn_smem[];
  SmallSmem& sm = *reinterpret_cast<SmallSmem*>(lap_dyn_smem);
  uint32_t nctas, crank;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(nctas));
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
  const int tid = threadIdx.x, lane = tid & 31;
  const int gtid = (int)crank * kLapThreads + tid;
  const int GT = (int)nctas * kLapThreads;
  const int gwarp = gtid >> 5, nwarps = GT >> 5;
  const int n = P.n, m = P.m;
  unsigned long long tq[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#define LAP_T(i) do { if (P.debug && gtid == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tq[i])); } while (0)
  LAP_T(0);

  // ---- P0 (all stages at once): outputs and per-stage scratch ----
  for (int stage = 0; stage < P.nstages; ++stage) {
    const LapStage S = P.st[stage];
    const size_t co = (size_t)stage * ws.cols, ro = (size_t)stage * ws.rows;
    for (int c = gtid; c < m; c += GT) {
      S.y[c] = -1;
      ws.collabel[co + c] = kInf;
      ws.v[co + c] = 0.0;
      ws.seen[co + c] = 0;
      ws.insc[co + c] = 0;
    }
    for (int r = gtid; r < n; r += GT) {
      S.x[r] = -1;
      ws.u[ro + r] = 0.0;
    }
  }
  if (gtid < 24) ws.counters[gtid] = 0;
  if (gtid == 0 && P.zero_word) *P.zero_word = 0;
  // Everything above touches only this kernel's own outputs and scratch, so under programmatic
  // dependent launch it runs while the emitting kernel drains; the candidate lists are read below.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  cluster_barrier();
  LAP_T(1);

  for (int stage = 0; stage < P.nstages; ++stage) {
    const LapStage S = P.st[stage];
    if (cand.total[S.list] == 0) continue;   // nothing was emitted for this stage (cluster-uniform)
    bt_lap_ws W = ws;                          // this stage's slices of the scratch arrays
    W.collabel += (size_t)stage * ws.cols; W.v += (size_t)stage * ws.cols;
    W.seen += (size_t)stage * ws.cols; W.insc += (size_t)stage * ws.cols; W.u += (size_t)stage * ws.rows;
    W.counters += stage * 8;
    int32_t* cnt = cand.deg + (size_t)S.list * cand.rows_cap;
    int32_t* ecol = cand.col + (size_t)S.list * cand.rows_cap * cand.stride;
    double* ecost = cand.cost + (size_t)S.list * cand.rows_cap * cand.stride;
    int32_t* x = S.x;
    int32_t* y = S.y;

    // ---- P1: classify every row from the emitters' degree bookkeeping (no edge traversal for the bulk):
    //      isolated edges (row degree 1, column in-degree 1) are final; rows with several candidates or a
    //      contested column are "complex": only those have their segments compacted ----
    const int32_t* indeg = cand.indeg + (size_t)S.list * cand.cols_cap;
    for (int r = gtid; r < n; r += GT) {
      W.label[r] = kInf;
      const size_t ri = (size_t)S.list * cand.rows_cap + r;
      const int deg = cand.rowdeg[ri];
      if (deg == 0) continue;                      // nothing was emitted for this row
      const int col1 = cand.rowcol[ri];
      const bool row_on = S.row_block == nullptr || S.row_block[r] < 0;
      if (P.clear_lists) cand.rowdeg[ri] = 0;
      const bool col1_ok = edge_ok(S.col_block, col1);
      const bool contested1 = deg == 1 && row_on && col1_ok && indeg[col1] != 1;
      const bool need_edges = row_on && (deg > 1 || contested1);
      int total = 0, valid = 0;
      unsigned long long mask = cand.segmask[ri];
      if (P.clear_lists) cand.segmask[ri] = 0ull;
      int32_t* segcnt = cand.cnt + ri * cand.nseg;
      int32_t* rc = ecol + (size_t)r * cand.stride;
      double* rv = ecost + (size_t)r * cand.stride;
      while (mask) {                               // only the non-empty segments, in ascending order
        const int g = __ffsll((long long)mask) - 1;
        mask &= mask - 1;
        const int k = need_edges ? segcnt[g] : 0;  // complex rows: compact in place (ascending copy)
        if (P.clear_lists) segcnt[g] = 0;
        const int src = g * cand.seg;
        for (int e = 0; e < k; ++e) {
          const int c = rc[src + e];
          if (src + e != total) { rc[total] = c; rv[total] = rv[src + e]; }
          ++total;
          if (edge_ok(S.col_block, c)) ++valid;
        }
      }
      cnt[r] = total;
      if (!row_on) continue;
      if (deg == 1 && !contested1) {               // isolated edge, or its only column is taken already
        if (col1_ok) { x[r] = col1; y[col1] = r; }
        continue;
      }
      if (valid == 0) continue;
      const int pos = atomicAdd(&W.counters[0], 1);
      atomicAdd(&W.counters[4], valid);            // valid edges of the complex part
      W.clist[pos] = r;
      W.label[r] = r;
    }
    cluster_barrier();
    LAP_T(2);
    const int nC = W.counters[0];
    if (P.clear_lists)                              // every read of the in-degrees is behind us
      for (int c = gtid; c < cand.cols_cap; c += GT) cand.indeg[(size_t)S.list * cand.cols_cap + c] = 0;
    LAP_T(3);

    const bool small = nC > 0 && nC <= kSmallRows && W.counters[4] <= kSmallEdges;
    if (small) {
      if (crank == 0 && tid < kSmallThreads) small_complex_solve(cand, W, S, nC, cnt, ecol, ecost, x, y, sm, P.debug);
      LAP_T(4); tq[5] = tq[4];
    } else if (nC > 0) {
      // ---- P3: connected components of the complex part ----
      for (int iter = 0;; ++iter) {
        int32_t* flag = &W.counters[2 + (iter & 1)];
        bool changed = false;
        for (int i = gtid; i < nC; i += GT) {
          const int r = W.clist[i];
          int lr = W.label[r];
          const int l0 = lr;
          const int deg = cnt[r];
          const int32_t* e = ecol + (size_t)r * cand.stride;
          for (int k = 0; k < deg; ++k) {
            const int c = e[k];
            if (!edge_ok(S.col_block, c)) continue;
            const int lc = W.collabel[c];
            if (lc < lr) lr = lc;
            else if (lc > lr) { atomicMin(&W.collabel[c], lr); changed = true; }
          }
          if (lr < l0) { atomicMin(&W.label[r], lr); changed = true; }
        }
        cluster_barrier();
        for (int i = gtid; i < nC; i += GT) {
          const int r = W.clist[i];
          const int l = W.label[r];
          const int ll = W.label[l];
          if (ll < l) { atomicMin(&W.label[r], ll); changed = true; }
        }
        if (changed) *flag = 1;
        if (gtid == 0) W.counters[2 + ((iter + 1) & 1)] = 0;   // the other flag, for the next round
        cluster_barrier();
        if (*flag == 0) break;
      }

      LAP_T(4);
      // ---- P4: component bookkeeping (CTA 0) ----
      if (crank == 0) {
        for (int i = tid; i < nC; i += kLapThreads) {
          const int r = W.clist[i];
          W.isroot[i] = (W.label[r] == r) ? 1 : 0;
        }
        for (int i = tid; i <= nC; i += kLapThreads) { W.rowcnt[i] = 0; W.colcnt[i] = 0; W.fill[i] = 0; }
        __syncthreads();
        const int ncomp = block_exclusive_scan(W.isroot, nC, s_warp, &s_carry);
        for (int i = tid; i < nC; i += kLapThreads) {
          const int r = W.clist[i];
          if (W.label[r] == r) W.compidx[r] = W.isroot[i];
        }
        __syncthreads();
        for (int i = tid; i < nC; i += kLapThreads) atomicAdd(&W.rowcnt[W.compidx[W.label[W.clist[i]]]], 1);
        for (int c = tid; c < m; c += kLapThreads) {
          const int l = W.collabel[c];
          if (l != kInf) atomicAdd(&W.colcnt[W.compidx[l]], 1);
        }
        if (tid == 0) W.counters[1] = ncomp;
        __syncthreads();
        block_exclusive_scan(W.rowcnt, ncomp + 1, s_warp, &s_carry);
        block_exclusive_scan(W.colcnt, ncomp + 1, s_warp, &s_carry);
        for (int i = tid; i < nC; i += kLapThreads) {
          const int r = W.clist[i];
          const int k = W.compidx[W.label[r]];
          W.sorted_rows[W.rowcnt[k] + atomicAdd(&W.fill[k], 1)] = r;
        }
      }
      cluster_barrier();
      LAP_T(5);

      // ---- P5: one warp per component ----
      const int ncomp = W.counters[1];
      const SolveArrays GA{W.rowcnt, W.colcnt, W.sorted_rows, W.touched, W.treerows, W.u, W.v, W.dist,
                           W.pathrow, W.seen, W.insc, nullptr, (size_t)cand.stride};
      for (int comp = gwarp; comp < ncomp; comp += nwarps)
        solve_component(GA, comp, S.thresh, S.col_block, cnt, ecol, ecost, x, y, lane);
    }
    if (stage + 1 < P.nstages) cluster_barrier();   // x / y of this stage gate the next one
    if (P.clear_lists && gtid == 0) cand.total[S.list] = 0;   // every CTA has read it (barriers above / kernel end)
    LAP_T(6);
    if (P.debug && gtid == 0)
      printf("lap stage %d: n=%d m=%d complex=%d comps=%d | init %llu classify %llu clear %llu complex-part %llu tail %llu ns\n", stage, n, m,
             nC, W.counters[1], tq[1] - tq[0], tq[2] - tq[1], tq[3] - tq[2], tq[5] - tq[3], tq[6] - tq[5]);
    LAP_T(1);
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
  const int rows = ctx->max_tracks, cols = ctx->max_dets;
  ws->rows = rows;
  ws->cols = cols;
  const int stride = (cols + 127) / 128 * 128 + 128;
  ws->cand.rows_cap = rows;
  ws->cand.stride = stride;
  ws->cand.nseg = BT_CAND_MAXSEG;
  ws->cand.seg = 128;
  const size_t cnt_ints = (3 * (size_t)rows * ws->cand.nseg + 4 + 1) & ~size_t(1);   // keeps segmask 8 B aligned
  ws->cand.cols_cap = cols;
  ws->cand.clear_bytes = sizeof(int32_t) * cnt_ints + sizeof(unsigned long long) * 3 * rows +
                         sizeof(int32_t) * 3 * ((size_t)rows + cols);
  BT_CUDA(cudaMalloc(&ws->cand.cnt, ws->cand.clear_bytes));
  ws->cand.total = ws->cand.cnt + 3 * (size_t)rows * ws->cand.nseg;   // cleared by the same memset as cnt
  ws->cand.segmask = reinterpret_cast<unsigned long long*>(ws->cand.cnt + cnt_ints);
  ws->cand.rowdeg = reinterpret_cast<int32_t*>(ws->cand.segmask + 3 * (size_t)rows);
  ws->cand.indeg = ws->cand.rowdeg + 3 * (size_t)rows;
  BT_CUDA(cudaMalloc(&ws->cand.rowcol, sizeof(int32_t) * 3 * (size_t)rows));
  BT_CUDA(cudaMalloc(&ws->cand.deg, sizeof(int32_t) * 3 * rows));
  BT_CUDA(cudaMalloc(&ws->cand.col, sizeof(int32_t) * 3 * (size_t)rows * stride));
  BT_CUDA(cudaMalloc(&ws->cand.cost, sizeof(double) * 3 * (size_t)rows * stride));
  BT_CUDA(cudaMemset(ws->cand.cnt, 0, ws->cand.clear_bytes));
  BT_CHECK(cols <= BT_CAND_MAXSEG * 112, BT_ERR_CAPACITY, "max_dets %d exceeds %d", cols, BT_CAND_MAXSEG * 112);
#define BT_LAP_ALLOC(field, type, count) BT_CUDA(cudaMalloc(&ws->field, sizeof(type) * (size_t)(count)))
  BT_LAP_ALLOC(label, int32_t, rows);
  BT_LAP_ALLOC(collabel, int32_t, 3 * (size_t)cols);
  BT_LAP_ALLOC(nvalid, int32_t, rows);
  BT_LAP_ALLOC(onlycol, int32_t, rows);
  BT_LAP_ALLOC(clist, int32_t, rows);
  BT_LAP_ALLOC(compidx, int32_t, rows);
  BT_LAP_ALLOC(isroot, int32_t, rows);
  BT_LAP_ALLOC(rowcnt, int32_t, rows + 1);
  BT_LAP_ALLOC(colcnt, int32_t, rows + 1);
  BT_LAP_ALLOC(fill, int32_t, rows + 1);
  BT_LAP_ALLOC(sorted_rows, int32_t, rows);
  BT_LAP_ALLOC(counters, int32_t, 24);
  BT_LAP_ALLOC(u, double, 3 * (size_t)rows);
  BT_LAP_ALLOC(v, double, 3 * (size_t)cols);
  BT_LAP_ALLOC(dist, double, cols);
  BT_LAP_ALLOC(pathrow, int32_t, cols);
  BT_LAP_ALLOC(seen, int32_t, 3 * (size_t)cols);
  BT_LAP_ALLOC(insc, int32_t, 3 * (size_t)cols);
  BT_LAP_ALLOC(touched, int32_t, cols);
  BT_LAP_ALLOC(treerows, int32_t, rows);
  BT_LAP_ALLOC(x, int32_t, rows);
  BT_LAP_ALLOC(y, int32_t, cols);
#undef BT_LAP_ALLOC
  return BT_OK;
}

void bt_lap_ws_destroy(bt_ctx* ctx) {
  bt_lap_ws* ws = ctx->lap;
  if (!ws) return;
  void* ptrs[] = {ws->cand.cnt, ws->cand.rowcol, ws->cand.deg, ws->cand.col, ws->cand.cost, ws->label, ws->collabel,
                  ws->nvalid, ws->onlycol, ws->clist, ws->compidx, ws->isroot, ws->rowcnt, ws->colcnt, ws->fill,
                  ws->sorted_rows, ws->counters, ws->u, ws->v, ws->dist, ws->pathrow, ws->seen, ws->insc,
                  ws->touched, ws->treerows, ws->x, ws->y};
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

static int32_t launch_lap(bt_ctx* ctx, const bt_cand& cand, const LapParams& P) {
  BT_CHECK(P.n <= ctx->lap->rows && P.m <= ctx->lap->cols, BT_ERR_CAPACITY,
           "linear assignment %d x %d exceeds ctx capacity %d x %d", P.n, P.m, ctx->lap->rows, ctx->lap->cols);
  if (P.n <= 0 && P.m <= 0) return BT_OK;
  const int work = P.n > P.m ? P.n : P.m;
  int nctas = (work + 255) / 256;          // ~256 rows per CTA keeps every thread busy in the row phases
  if (nctas < 1) nctas = 1;
  if (nctas > kLapMaxCtas) nctas = kLapMaxCtas;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(nctas);
  cfg.blockDim = dim3(kLapThreads);
  cfg.dynamicSmemBytes = sizeof(SmallSmem);
  cfg.stream = ctx->stream;
  static bool attr_done = false;
  if (!attr_done) {
    BT_CUDA(cudaFuncSetAttribute(lap_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)sizeof(SmallSmem)));
    attr_done = true;
  }
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = nctas;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = ctx->pdl ? 2 : 1;
  BT_CUDA(cudaLaunchKernelEx(&cfg, lap_cluster_kernel, cand, *ctx->lap, P));
  BT_LAUNCHED(ctx);
  return BT_OK;
}

int32_t btk_lap_solve(bt_ctx* ctx, const bt_cand& cand, int32_t list, int32_t n, int32_t m,
                      double thresh, const int32_t* row_block, const int32_t* col_block, int32_t* x,
                      int32_t* y) {
  LapParams P = {};
  P.nstages = 1;
  P.n = n;
  P.m = m;
  P.st[0] = LapStage{list, thresh, row_block, col_block, x, y};
  return launch_lap(ctx, cand, P);
}

int32_t btk_lap_solve3(bt_ctx* ctx, const bt_cand& cand, int32_t n, int32_t m, const double thresh[3],
                       int32_t* const x[3], int32_t* const y[3], int32_t* zero_word) {
  LapParams P = {};
  P.clear_lists = 1;
  P.zero_word = zero_word;
  P.nstages = 3;
  P.n = n;
  P.m = m;
  P.st[0] = LapStage{0, thresh[0], nullptr, nullptr, x[0], y[0]};
  P.st[1] = LapStage{1, thresh[1], x[0], nullptr, x[1], y[1]};   // stage 2: rows unmatched in stage 1
  P.st[2] = LapStage{2, thresh[2], nullptr, y[0], x[2], y[2]};   // stage 3: columns unmatched in stage 1
  P.debug = getenv("BT_LAP_DEBUG") != nullptr;
  return launch_lap(ctx, cand, P);
}
