// N x M IoU-distance kernels, replacing bbox_iou / bbox_ious / iou_distance of the reference
// (demo:1695-1761; demo = /root/reference/demo_bottrack_onnx_tflite.py).  float64 like the
// reference (a float64 track box against an integer-valued detection box), strict `<=`
// empty-intersection rule and no +1 pixel convention (demo:1702).
//
// The dense kernel is HBM-write-bound: algorithmic bytes 32*(N+M) + 8*N*M.  Each CTA stages
// 64 row boxes + 64 column boxes in shared memory and every thread produces a 4x4 patch
// whose rows are written as 2 x double2 (32 B sectors fully used).
#include "common.cuh"

namespace {

struct Box { double x1, y1, x2, y2; };

// bbox_iou, demo:1695-1713
__device__ __forceinline__ double iou_of(const Box& a, const Box& b) {
  const double ixmin = fmax(a.x1, b.x1), iymin = fmax(a.y1, b.y1);
  const double ixmax = fmin(a.x2, b.x2), iymax = fmin(a.y2, b.y2);
  if (ixmax <= ixmin || iymax <= iymin) return 0.0;
  const double inter = (ixmax - ixmin) * (iymax - iymin);
  const double area1 = (a.x2 - a.x1) * (a.y2 - a.y1);
  const double area2 = (b.x2 - b.x1) * (b.y2 - b.y1);
  return inter / (area1 + area2 - inter);
}

constexpr int kTile = 64;

__global__ void __launch_bounds__(256)
iou_distance_kernel(const double* __restrict__ a, int n, const double* __restrict__ b, int m,
                    double* __restrict__ out) {
  __shared__ Box sa[kTile];
  __shared__ Box sb[kTile];
  const int row0 = blockIdx.y * kTile, col0 = blockIdx.x * kTile;
  const int tid = threadIdx.x;
  if (tid < kTile) {
    const int r = row0 + tid;
    if (r < n) {
      const double2* p = reinterpret_cast<const double2*>(a + (size_t)r * 4);
      const double2 lo = p[0], hi = p[1];
      sa[tid] = Box{lo.x, lo.y, hi.x, hi.y};
    }
  } else if (tid < 2 * kTile) {
    const int c = col0 + tid - kTile;
    if (c < m) {
      const double2* p = reinterpret_cast<const double2*>(b + (size_t)c * 4);
      const double2 lo = p[0], hi = p[1];
      sb[tid - kTile] = Box{lo.x, lo.y, hi.x, hi.y};
    }
  }
  __syncthreads();
  const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads, 4 x 4 patch each
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int lr = ty * 4 + i, r = row0 + lr;
    if (r >= n) continue;
    double v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int lc = tx * 4 + j;
      v[j] = (col0 + lc < m) ? 1.0 - iou_of(sa[lr], sb[lc]) : 0.0;
    }
    double* dst = out + (size_t)r * m + col0 + tx * 4;
    const int c = col0 + tx * 4;
    if (c + 3 < m && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
      reinterpret_cast<double2*>(dst)[0] = make_double2(v[0], v[1]);
      reinterpret_cast<double2*>(dst)[1] = make_double2(v[2], v[3]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (c + j < m) dst[j] = v[j];
    }
  }
}

__global__ void fuse_score_kernel(const double* __restrict__ d, const double* __restrict__ s, int n, int m,
                                  double* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)n * m) return;
  const int c = (int)(i % m);
  const double sim = 1 - d[i];
  out[i] = 1 - sim * s[c];
}

// remove_duplicate_stracks' test (demo:1665-1668): pairs (p, q) with 1 - IoU < limit between two
// lists of track slots.  Sparse output (the matrix itself is never needed).
__global__ void __launch_bounds__(256)
iou_pairs_below_kernel(const double* __restrict__ tlbr, const int32_t* __restrict__ a_idx, int n,
                       const int32_t* __restrict__ b_idx, int m, double limit,
                       int32_t* __restrict__ pairs, int32_t* __restrict__ pair_count, int pair_cap) {
  __shared__ Box sb[256];
  const int p = blockIdx.x * 256 + threadIdx.x;
  Box a{0, 0, 0, 0};
  if (p < n) {
    const double* s = tlbr + (size_t)a_idx[p] * 4;
    a = Box{s[0], s[1], s[2], s[3]};
  }
  for (int c0 = 0; c0 < m; c0 += 256) {
    __syncthreads();
    if (c0 + threadIdx.x < m) {
      const double* s = tlbr + (size_t)b_idx[c0 + threadIdx.x] * 4;
      sb[threadIdx.x] = Box{s[0], s[1], s[2], s[3]};
    }
    __syncthreads();
    if (p < n) {
      const int lim = min(256, m - c0);
      for (int j = 0; j < lim; ++j) {
        const double dist = 1.0 - iou_of(a, sb[j]);
        if (dist < limit) {
          const int slot = atomicAdd(pair_count, 1);
          if (slot < pair_cap) {
            pairs[2 * slot] = p;
            pairs[2 * slot + 1] = c0 + j;
          }
        }
      }
    }
  }
}

}  // namespace

int32_t btk_iou_distance(bt_ctx* ctx, const double* a, int32_t n, const double* b, int32_t m,
                         double* out) {
  if (n <= 0 || m <= 0) return BT_OK;
  dim3 grid((m + kTile - 1) / kTile, (n + kTile - 1) / kTile);
  iou_distance_kernel<<<grid, 256, 0, ctx->stream>>>(a, n, b, m, out);
  BT_LAUNCHED(ctx);
  return BT_OK;
}

int32_t btk_fuse_score(bt_ctx* ctx, const double* d, const double* s, int32_t n, int32_t m, double* out) {
  if (n <= 0 || m <= 0) return BT_OK;
  const size_t total = (size_t)n * m;
  fuse_score_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(d, s, n, m, out);
  BT_LAUNCHED(ctx);
  return BT_OK;
}

int32_t btk_iou_pairs_below(bt_ctx* ctx, const double* tlbr, const int32_t* a_idx, int32_t n,
                            const int32_t* b_idx, int32_t m, double limit, int32_t* pairs,
                            int32_t* pair_count, int32_t pair_cap) {
  if (n <= 0 || m <= 0) return BT_OK;
  iou_pairs_below_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(tlbr, a_idx, n, b_idx, m, limit,
                                                                    pairs, pair_count, pair_cap);
  BT_LAUNCHED(ctx);
  return BT_OK;
}
