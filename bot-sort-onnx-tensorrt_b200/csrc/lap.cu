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
//   1. lap_setup_kernel  (one CTA): connected components of the candidate graph by min-label
//      propagation with pointer jumping; component row lists / scratch slices by block scans.
//   2. lap_solve_kernel  (one warp per component): exact shortest-augmenting-path assignment
//      (Jonker-Volgenant / Crouse formulation, float64 duals) where every row owns a private
//      zero-cost dummy column ("stay unmatched"); lanes parallelise the edge relaxations and the
//      minimum scans.  Components of a tracking scene are tiny (1-5 rows), so thousands of
//      warps run independently; a single giant component is still solved exactly, just slower.
// Ties between equal-cost optima are broken by lowest index, lap's own tie-breaking is not
// reproducible without its sources: "bit-exact" is defined on inputs with a unique optimum.
#include "common.cuh"

#include <float.h>

struct bt_lap_ws {
  bt_cand cand;            // ctx-wide candidate lists (3 lists)
  int32_t* label = nullptr;     // [rows]
  int32_t* collabel = nullptr;  // [cols]
  int32_t* compidx = nullptr;   // [rows]  component index of a root row
  int32_t* rowcnt = nullptr;    // [rows+1] per component -> exclusive scan = row_start
  int32_t* colcnt = nullptr;    // [rows+1] per component -> exclusive scan = col_start
  int32_t* fill = nullptr;      // [rows]
  int32_t* sorted_rows = nullptr;  // [rows]
  int32_t* comp_root = nullptr;    // [rows]
  int32_t* ncomp = nullptr;        // [1]
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

constexpr int kSetupThreads = 1024;
constexpr int kInf = 0x7fffffff;

__device__ __forceinline__ bool edge_ok(const int32_t* __restrict__ col_block, int c) {
  return col_block == nullptr || col_block[c] < 0;
}

// exclusive scan of data[0..n) in place, returns total; all threads of the (single) CTA call it
__device__ int block_exclusive_scan(int32_t* data, int n, int32_t* s_warp, int32_t* s_carry) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) *s_carry = 0;
  __syncthreads();
  for (int base = 0; base < n; base += kSetupThreads) {
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
      s_warp[lane] = w;  // inclusive over warps
    }
    __syncthreads();
    const int carry = *s_carry;
    const int warp_off = (warp == 0) ? 0 : s_warp[warp - 1];
    if (i < n) data[i] = carry + warp_off + incl - val;
    __syncthreads();
    if (tid == kSetupThreads - 1) *s_carry = carry + s_warp[31];
    __syncthreads();
  }
  return *s_carry;
}

__global__ void __launch_bounds__(kSetupThreads)
lap_setup_kernel(bt_cand cand, int list, int n, int m, const int32_t* __restrict__ row_block,
                 const int32_t* __restrict__ col_block, bt_lap_ws ws, int32_t* __restrict__ x,
                 int32_t* __restrict__ y) {
  __shared__ int32_t s_warp[32];
  __shared__ int32_t s_carry;
  __shared__ int s_changed;
  const int tid = threadIdx.x;
  const int32_t* cnt = cand.cnt + (size_t)list * cand.rows_cap;
  const int32_t* ecol = cand.col + (size_t)list * cand.rows_cap * cand.stride;

  for (int c = tid; c < m; c += kSetupThreads) {
    y[c] = -1;
    ws.collabel[c] = kInf;
    ws.v[c] = 0.0;
    ws.seen[c] = 0;
    ws.insc[c] = 0;
  }
  for (int r = tid; r <= n; r += kSetupThreads) {
    ws.rowcnt[r] = 0;
    ws.colcnt[r] = 0;
  }
  for (int r = tid; r < n; r += kSetupThreads) {
    x[r] = -1;
    ws.u[r] = 0.0;
    ws.fill[r] = 0;
    int lab = kInf;
    if (row_block == nullptr || row_block[r] < 0) {
      const int deg = cnt[r];
      const int32_t* e = ecol + (size_t)r * cand.stride;
      for (int k = 0; k < deg; ++k)
        if (edge_ok(col_block, e[k])) { lab = r; break; }
    }
    ws.label[r] = lab;
  }
  __syncthreads();

  // ---- connected components: min-label propagation over edges + pointer jumping ----
  while (true) {
    if (tid == 0) s_changed = 0;
    __syncthreads();
    bool changed = false;
    for (int r = tid; r < n; r += kSetupThreads) {
      int lr = ws.label[r];
      if (lr == kInf) continue;
      const int deg = cnt[r];
      const int32_t* e = ecol + (size_t)r * cand.stride;
      const int l0 = lr;
      for (int k = 0; k < deg; ++k) {
        const int c = e[k];
        if (!edge_ok(col_block, c)) continue;
        const int lc = ws.collabel[c];
        if (lc < lr) lr = lc;
        else if (lc > lr) { atomicMin(&ws.collabel[c], lr); changed = true; }
      }
      if (lr < l0) { atomicMin(&ws.label[r], lr); changed = true; }
    }
    __syncthreads();
    for (int r = tid; r < n; r += kSetupThreads) {
      const int l = ws.label[r];
      if (l == kInf) continue;
      const int ll = ws.label[l];
      if (ll < l) { atomicMin(&ws.label[r], ll); changed = true; }
    }
    if (changed) s_changed = 1;
    __syncthreads();
    const int any = s_changed;
    __syncthreads();
    if (!any) break;
  }

  // ---- component bookkeeping ----
  for (int r = tid; r < n; r += kSetupThreads) ws.compidx[r] = (ws.label[r] == r) ? 1 : 0;
  __syncthreads();
  const int ncomp = block_exclusive_scan(ws.compidx, n, s_warp, &s_carry);
  for (int r = tid; r < n; r += kSetupThreads) {
    const int l = ws.label[r];
    if (l == kInf) continue;
    if (l == r) ws.comp_root[ws.compidx[r]] = r;
    atomicAdd(&ws.rowcnt[ws.compidx[l]], 1);
  }
  for (int c = tid; c < m; c += kSetupThreads) {
    const int l = ws.collabel[c];
    if (l != kInf) atomicAdd(&ws.colcnt[ws.compidx[l]], 1);
  }
  if (tid == 0) { *ws.ncomp = ncomp; }
  __syncthreads();
  block_exclusive_scan(ws.rowcnt, ncomp + 1, s_warp, &s_carry);
  block_exclusive_scan(ws.colcnt, ncomp + 1, s_warp, &s_carry);
  for (int r = tid; r < n; r += kSetupThreads) {
    const int l = ws.label[r];
    if (l == kInf) continue;
    const int k = ws.compidx[l];
    const int pos = ws.rowcnt[k] + atomicAdd(&ws.fill[k], 1);
    ws.sorted_rows[pos] = r;
  }
}

struct MinPair { double d; int c; int freecol; };

__device__ __forceinline__ bool better(const MinPair& a, const MinPair& b) {
  // smaller distance; ties: unassigned column first, then lower column index
  if (a.d != b.d) return a.d < b.d;
  if (a.freecol != b.freecol) return a.freecol > b.freecol;
  return a.c < b.c;
}

__global__ void __launch_bounds__(256)
lap_solve_kernel(bt_cand cand, int list, double thresh, const int32_t* __restrict__ col_block,
                 bt_lap_ws ws, int32_t* __restrict__ x, int32_t* __restrict__ y) {
  const int lane = threadIdx.x & 31;
  const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int ncomp = *ws.ncomp;
  const int32_t* cnt = cand.cnt + (size_t)list * cand.rows_cap;
  const int32_t* ecol = cand.col + (size_t)list * cand.rows_cap * cand.stride;
  const double* ecost = cand.cost + (size_t)list * cand.rows_cap * cand.stride;

  for (int comp = warp_global; comp < ncomp; comp += nwarps) {
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
        const int c = ecol[(size_t)r * cand.stride + k];
        if (!edge_ok(col_block, c)) continue;
        MinPair cur{ecost[(size_t)r * cand.stride + k], c, 0};
        if (better(cur, best)) best = cur;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        MinPair oth{__shfl_xor_sync(0xffffffffu, best.d, o), __shfl_xor_sync(0xffffffffu, best.c, o), 0};
        if (better(oth, best)) best = oth;
      }
      if (lane == 0 && best.c != kInf) { x[r] = best.c; y[best.c] = r; }
      continue;
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
            c = ecol[(size_t)i * cand.stride + k];
            if (edge_ok(col_block, c) && ws.insc[c] != sid) {
              const double r = minVal + (ecost[(size_t)i * cand.stride + k] - thresh) - ui - ws.v[c];
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

// dense float64 cost -> candidate list (ordered by column): one warp per row
__global__ void __launch_bounds__(256)
lap_compact_dense_kernel(const double* __restrict__ cost, int n, int m, double thresh, bt_cand cand,
                         int list) {
  const int lane = threadIdx.x & 31;
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= n) return;
  int32_t* ecol = cand.col + ((size_t)list * cand.rows_cap + row) * cand.stride;
  double* ecost = cand.cost + ((size_t)list * cand.rows_cap + row) * cand.stride;
  int count = 0;
  for (int c0 = 0; c0 < m; c0 += 32) {
    const int c = c0 + lane;
    const double v = (c < m) ? cost[(size_t)row * m + c] : DBL_MAX;
    const bool keep = (c < m) && (v < thresh);
    const unsigned ball = __ballot_sync(0xffffffffu, keep);
    if (keep) {
      const int pos = count + __popc(ball & ((1u << lane) - 1));
      ecol[pos] = c;
      ecost[pos] = v;
    }
    count += __popc(ball);
  }
  if (lane == 0) cand.cnt[(size_t)list * cand.rows_cap + row] = count;
}

}  // namespace

int32_t bt_lap_ws_create(bt_ctx* ctx) {
  auto* ws = new bt_lap_ws();
  ctx->lap = ws;
  const int rows = ctx->max_tracks, cols = ctx->max_dets;
  ws->rows = rows;
  ws->cols = cols;
  ws->cand.rows_cap = rows;
  ws->cand.stride = cols;
  BT_CUDA(cudaMalloc(&ws->cand.cnt, sizeof(int32_t) * 3 * rows));
  BT_CUDA(cudaMalloc(&ws->cand.col, sizeof(int32_t) * 3 * (size_t)rows * cols));
  BT_CUDA(cudaMalloc(&ws->cand.cost, sizeof(double) * 3 * (size_t)rows * cols));
  BT_CUDA(cudaMemset(ws->cand.cnt, 0, sizeof(int32_t) * 3 * rows));
#define BT_LAP_ALLOC(field, type, count) BT_CUDA(cudaMalloc(&ws->field, sizeof(type) * (size_t)(count)))
  BT_LAP_ALLOC(label, int32_t, rows);
  BT_LAP_ALLOC(collabel, int32_t, cols);
  BT_LAP_ALLOC(compidx, int32_t, rows);
  BT_LAP_ALLOC(rowcnt, int32_t, rows + 1);
  BT_LAP_ALLOC(colcnt, int32_t, rows + 1);
  BT_LAP_ALLOC(fill, int32_t, rows);
  BT_LAP_ALLOC(sorted_rows, int32_t, rows);
  BT_LAP_ALLOC(comp_root, int32_t, rows);
  BT_LAP_ALLOC(ncomp, int32_t, 1);
  BT_LAP_ALLOC(u, double, rows);
  BT_LAP_ALLOC(v, double, cols);
  BT_LAP_ALLOC(dist, double, cols);
  BT_LAP_ALLOC(pathrow, int32_t, cols);
  BT_LAP_ALLOC(seen, int32_t, cols);
  BT_LAP_ALLOC(insc, int32_t, cols);
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
  void* ptrs[] = {ws->cand.cnt, ws->cand.col, ws->cand.cost, ws->label, ws->collabel, ws->compidx,
                  ws->rowcnt, ws->colcnt, ws->fill, ws->sorted_rows, ws->comp_root, ws->ncomp, ws->u,
                  ws->v, ws->dist, ws->pathrow, ws->seen, ws->insc, ws->touched, ws->treerows, ws->x, ws->y};
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

int32_t btk_lap_solve(bt_ctx* ctx, const bt_cand& cand, int32_t list, int32_t n, int32_t m,
                      double thresh, const int32_t* row_block, const int32_t* col_block, int32_t* x,
                      int32_t* y) {
  BT_CHECK(n <= ctx->lap->rows && m <= ctx->lap->cols, BT_ERR_CAPACITY,
           "linear assignment %d x %d exceeds ctx capacity %d x %d", n, m, ctx->lap->rows, ctx->lap->cols);
  if (n <= 0 && m <= 0) return BT_OK;
  lap_setup_kernel<<<1, kSetupThreads, 0, ctx->stream>>>(cand, list, n, m, row_block, col_block, *ctx->lap, x, y);
  BT_LAUNCHED(ctx);
  if (n > 0 && m > 0) {
    const int warps = n;
    int blocks = (warps * 32 + 255) / 256;
    if (blocks > 4 * ctx->num_sms) blocks = 4 * ctx->num_sms;
    lap_solve_kernel<<<blocks, 256, 0, ctx->stream>>>(cand, list, thresh, col_block, *ctx->lap, x, y);
    BT_LAUNCHED(ctx);
  }
  return BT_OK;
}
