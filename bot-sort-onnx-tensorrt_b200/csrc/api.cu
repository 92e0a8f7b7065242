// C-ABI entry points of libbotsort_b200.so (include/botsort_b200.h): ctx lifetime, staging of
// host/device buffers and the stand-alone (stateless) kernels.  The stateful tracker entry
// points live in track_step.cu.
#include "common.cuh"

#include <stdarg.h>
#include <string.h>

thread_local std::string g_bt_create_error;

int32_t bt_fail(bt_ctx* ctx, int32_t code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (ctx) ctx->err = buf;
  else g_bt_create_error = buf;
  return code;
}

// ---- arena ------------------------------------------------------------------------------------
int32_t bt_arena_reset(bt_ctx* ctx) {
  ctx->arena_off = 0;
  return BT_OK;
}

int32_t bt_arena_alloc(bt_ctx* ctx, size_t bytes, void** out) {
  const size_t aligned = (bytes + 255) & ~size_t(255);
  if (ctx->arena_off + aligned > ctx->arena_cap) {
    // grow: everything enqueued so far must finish before the old arena is released; live
    // allocations of the current call are not preserved, so growth is only legal at offset 0.
    // Calls therefore reserve their total first (bt_arena_reserve).
    return bt_fail(ctx, BT_ERR_STATE, "arena exhausted (%zu + %zu > %zu)", ctx->arena_off, aligned,
                   ctx->arena_cap);
  }
  *out = ctx->arena + ctx->arena_off;
  ctx->arena_off += aligned;
  return BT_OK;
}

int32_t bt_arena_reserve(bt_ctx* ctx, size_t total) {
  ctx->arena_off = 0;
  total += 64 * 256;  // alignment slack for up to 64 allocations
  if (total <= ctx->arena_cap) return BT_OK;
  BT_CUDA(cudaStreamSynchronize(ctx->stream));
  if (ctx->arena) BT_CUDA(cudaFree(ctx->arena));
  ctx->arena = nullptr;
  ctx->arena_cap = 0;
  size_t cap = total + total / 2;
  BT_CUDA(cudaMalloc(&ctx->arena, cap));
  ctx->arena_cap = cap;
  return BT_OK;
}

int32_t bt_stage_in(bt_ctx* ctx, const void* src, size_t bytes, int32_t loc, const void** dev) {
  if (loc == BT_DEVICE || bytes == 0) {
    *dev = src;
    return BT_OK;
  }
  void* p = nullptr;
  BT_TRY(bt_arena_alloc(ctx, bytes, &p));
  BT_CUDA(cudaMemcpyAsync(p, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
  *dev = p;
  return BT_OK;
}

int32_t bt_stage_out(bt_ctx* ctx, void* dst, size_t bytes, int32_t loc, void** dev) {
  if (loc == BT_DEVICE || bytes == 0) {
    *dev = dst;
    return BT_OK;
  }
  return bt_arena_alloc(ctx, bytes, dev);
}

int32_t bt_unstage_out(bt_ctx* ctx, void* dst, const void* dev, size_t bytes, int32_t loc) {
  if (loc == BT_DEVICE || bytes == 0) return BT_OK;
  BT_CUDA(cudaMemcpyAsync(dst, dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  return BT_OK;
}

int32_t bt_finish(bt_ctx* ctx, int32_t loc) {
  if (loc == BT_HOST) BT_CUDA(cudaStreamSynchronize(ctx->stream));
  return BT_OK;
}

#define BT_ENTER(ctx)                                                                              \
  do {                                                                                             \
    if (!(ctx)) return BT_ERR_INVALID;                                                             \
    if (cudaSetDevice((ctx)->device) != cudaSuccess)                                               \
      return bt_fail(ctx, BT_ERR_CUDA, "cudaSetDevice(%d) failed", (ctx)->device);                 \
  } while (0)

#define BT_LOC_OK(loc) BT_CHECK((loc) == BT_HOST || (loc) == BT_DEVICE, BT_ERR_INVALID, "bad loc %d", (int)(loc))

extern "C" {

int32_t bt_version(void) { return BT_VERSION; }

const char* bt_last_error(const bt_ctx* ctx) { return ctx ? ctx->err.c_str() : g_bt_create_error.c_str(); }

void bt_default_config(bt_config* cfg) {
  cfg->track_high_thresh = 0.40;
  cfg->track_low_thresh = 0.1;
  cfg->new_track_thresh = 0.9;
  cfg->match_thresh = 0.8;
  cfg->second_thresh = 0.5;
  cfg->unconfirmed_thresh = 0.7;
  cfg->proximity_thresh = 0.5;
  cfg->appearance_thresh = 0.25;
  cfg->duplicate_iou_dist = 0.15;
  cfg->ema_alpha = 0.9;
  cfg->track_buffer = 300;
  cfg->frame_rate = 30;
  cfg->with_reid = 1;
  cfg->reserved = 0;
}

void bt_default_yolox_config(bt_yolox_config* cfg) {
  cfg->in_h = 480;
  cfg->in_w = 640;
  cfg->img_h = 480;
  cfg->img_w = 640;
  cfg->num_classes = 4;
  cfg->nms_score_thresh = 0.15f;
  cfg->nms_iou_thresh = 0.80f;
  cfg->max_per_class = 50;
  cfg->post_score_thresh = 0.35f;
}

int32_t bt_create(int32_t device, int32_t max_tracks, int32_t max_dets, int32_t feat_dim, uint32_t flags,
                  bt_ctx** out) {
  return bt_create_streams(device, 1, max_tracks, max_dets, feat_dim, flags, out);
}

int32_t bt_num_streams(const bt_ctx* ctx) { return ctx ? ctx->n_streams : 0; }

int32_t bt_create_streams(int32_t device, int32_t n_streams, int32_t max_tracks, int32_t max_dets, int32_t feat_dim,
                          uint32_t flags, bt_ctx** out) {
  if (!out) return bt_fail(nullptr, BT_ERR_INVALID, "out is NULL");
  *out = nullptr;
  if (max_tracks <= 0 || max_dets <= 0 || feat_dim <= 0)
    return bt_fail(nullptr, BT_ERR_INVALID, "max_tracks, max_dets, feat_dim must be positive");
  if (n_streams < 1 || n_streams > 1024)
    return bt_fail(nullptr, BT_ERR_INVALID, "n_streams must be 1..1024");
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count <= 0)
    return bt_fail(nullptr, BT_ERR_CUDA, "no CUDA device: %s (this library has no CPU path)",
                   cudaGetErrorString(e));
  if (device < 0 || device >= count) return bt_fail(nullptr, BT_ERR_INVALID, "device %d out of range", device);
  if (cudaSetDevice(device) != cudaSuccess) return bt_fail(nullptr, BT_ERR_CUDA, "cudaSetDevice failed");
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess)
    return bt_fail(nullptr, BT_ERR_CUDA, "cudaGetDeviceProperties failed");
  if (prop.major != 10)
    return bt_fail(nullptr, BT_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", device,
                   prop.major, prop.minor);
  bt_ctx* c = new bt_ctx();
  c->device = device;
  c->max_tracks = (max_tracks + 127) / 128 * 128;
  c->max_dets = (max_dets + 3) / 4 * 4;   // keeps every stream's detection rows 16-byte aligned
  c->feat_dim = feat_dim;
  c->n_streams = n_streams;
  c->flags = flags;
  c->num_sms = prop.multiProcessorCount;
  c->pdl = getenv("BT_NO_PDL") ? 0 : 1;
  int32_t s = BT_OK;
  auto fail = [&](int32_t code) {
    g_bt_create_error = c->err;
    bt_destroy(c);
    return code;
  };
  if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&c->side_stream, cudaStreamNonBlocking) != cudaSuccess) {
    bt_fail(c, BT_ERR_CUDA, "cudaStreamCreate failed");
    return fail(BT_ERR_CUDA);
  }
  if (cudaMalloc(&c->d_desc, kBtCropLutOffset + 3 * 256 * sizeof(float)) != cudaSuccess) {
    bt_fail(c, BT_ERR_CUDA, "cudaMalloc failed");
    return fail(BT_ERR_CUDA);
  }
  {
    // FastReID._preprocess (demo:1101-1142): (pixel / 255 - mean) / std in float64, then float32 -- one value per
    // (RGB plane, 8-bit pixel value), tabulated once (IEEE float64 division: the host's results are the device's)
    static const float mean[3] = {0.485f, 0.456f, 0.406f}, stdv[3] = {0.229f, 0.224f, 0.225f};
    float lut[3][256];
    for (int ch = 0; ch < 3; ++ch)
      for (int v = 0; v < 256; ++v) {
        volatile double q = (double)v / 255.0;
        volatile double d = q - (double)mean[ch];
        lut[ch][v] = (float)(d / (double)stdv[ch]);
      }
    if (cudaMemcpy(c->d_desc + kBtCropLutOffset, lut, sizeof(lut), cudaMemcpyHostToDevice) != cudaSuccess) {
      bt_fail(c, BT_ERR_CUDA, "cudaMemcpy failed");
      return fail(BT_ERR_CUDA);
    }
  }
  if ((s = bt_arena_reserve(c, 1 << 20)) != BT_OK) return fail(s);
  if ((s = bt_lap_ws_create(c)) != BT_OK) return fail(s);
  if ((s = bt_gemm_ws_create(c)) != BT_OK) return fail(s);
  if ((s = bt_tracker_create(c)) != BT_OK) return fail(s);
  *out = c;
  return BT_OK;
}

int32_t bt_destroy(bt_ctx* ctx) {
  if (!ctx) return BT_OK;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
  if (ctx->side_stream) cudaStreamSynchronize(ctx->side_stream);
  bt_tracker_destroy(ctx);
  bt_gemm_ws_destroy(ctx);
  bt_lap_ws_destroy(ctx);
  if (ctx->arena) cudaFree(ctx->arena);
  if (ctx->d_desc) cudaFree(ctx->d_desc);
  if (ctx->pinned) cudaFreeHost(ctx->pinned);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->side_stream) cudaStreamDestroy(ctx->side_stream);
  delete ctx;
  return BT_OK;
}

int32_t bt_sync(bt_ctx* ctx) {
  BT_ENTER(ctx);
  BT_CUDA(cudaStreamSynchronize(ctx->stream));
  return BT_OK;
}

void* bt_stream(bt_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

int64_t bt_launch_count(const bt_ctx* ctx) { return ctx ? ctx->launches : 0; }

// ---- Kalman --------------------------------------------------------------------------------------
int32_t bt_kalman_initiate(bt_ctx* ctx, const float* xywh, double* mean, double* cov, int32_t k, int32_t loc) {
  BT_ENTER(ctx);
  BT_LOC_OK(loc);
  BT_CHECK(k >= 0, BT_ERR_INVALID, "k < 0");
  if (k == 0) return BT_OK;
  BT_CHECK(xywh && mean && cov, BT_ERR_INVALID, "NULL buffer");
  BT_TRY(bt_arena_reserve(ctx, loc == BT_HOST ? (size_t)k * (16 + 64 + 512) : 0));
  const float* d_z; double *d_m, *d_c;
  BT_TRY(bt_in(ctx, xywh, (size_t)k * 4, loc, &d_z));
  BT_TRY(bt_out(ctx, mean, (size_t)k * 8, loc, &d_m));
  BT_TRY(bt_out(ctx, cov, (size_t)k * 64, loc, &d_c));
  BT_TRY(btk_kalman_initiate(ctx, d_z, nullptr, d_m, d_c, nullptr, nullptr, nullptr, k));
  BT_TRY(bt_unstage_out(ctx, mean, d_m, sizeof(double) * k * 8, loc));
  BT_TRY(bt_unstage_out(ctx, cov, d_c, sizeof(double) * k * 64, loc));
  return bt_finish(ctx, loc);
}

int32_t bt_kalman_multi_predict(bt_ctx* ctx, double* mean, double* cov, const int32_t* state, int32_t n,
                                int32_t noise_f32, int32_t loc) {
  BT_ENTER(ctx);
  BT_LOC_OK(loc);
  BT_CHECK(n >= 0, BT_ERR_INVALID, "n < 0");
  if (n == 0) return BT_OK;  // STrack.multi_predict is a no-op on an empty list (demo:526)
  BT_CHECK(mean && cov, BT_ERR_INVALID, "NULL buffer");
  BT_TRY(bt_arena_reserve(ctx, loc == BT_HOST ? (size_t)n * (64 + 512 + 4) : 0));
  const double *d_m_in, *d_c_in; const int32_t* d_s = nullptr;
  BT_TRY(bt_in(ctx, (const double*)mean, (size_t)n * 8, loc, &d_m_in));
  BT_TRY(bt_in(ctx, (const double*)cov, (size_t)n * 64, loc, &d_c_in));
  if (state) BT_TRY(bt_in(ctx, state, (size_t)n, loc, &d_s));
  double* d_m = const_cast<double*>(d_m_in);
  double* d_c = const_cast<double*>(d_c_in);
  BT_TRY(btk_kalman_predict(ctx, d_m, d_c, nullptr, nullptr, d_s, nullptr, n, noise_f32));
  BT_TRY(bt_unstage_out(ctx, mean, d_m, sizeof(double) * n * 8, loc));
  BT_TRY(bt_unstage_out(ctx, cov, d_c, sizeof(double) * n * 64, loc));
  return bt_finish(ctx, loc);
}

int32_t bt_kalman_update(bt_ctx* ctx, double* mean, double* cov, const double* meas, const int32_t* track_idx,
                         const int32_t* meas_idx, const uint8_t* noise_f32, int32_t k, int32_t loc) {
  BT_ENTER(ctx);
  BT_LOC_OK(loc);
  BT_CHECK(k >= 0, BT_ERR_INVALID, "k < 0");
  if (k == 0) return BT_OK;
  BT_CHECK(mean && cov && meas, BT_ERR_INVALID, "NULL buffer");
  BT_CHECK(loc == BT_DEVICE || (track_idx == nullptr && meas_idx == nullptr), BT_ERR_INVALID,
           "index lists are only supported for device buffers (host callers gather first)");
  BT_TRY(bt_arena_reserve(ctx, loc == BT_HOST ? (size_t)k * (64 + 512 + 32 + 1) : 0));
  const double *d_m_in, *d_c_in, *d_z; const uint8_t* d_f = nullptr;
  BT_TRY(bt_in(ctx, (const double*)mean, (size_t)k * 8, loc, &d_m_in));
  BT_TRY(bt_in(ctx, (const double*)cov, (size_t)k * 64, loc, &d_c_in));
  BT_TRY(bt_in(ctx, meas, (size_t)k * 4, loc, &d_z));
  if (noise_f32) BT_TRY(bt_in(ctx, noise_f32, (size_t)k, loc, &d_f));
  double* d_m = const_cast<double*>(d_m_in);
  double* d_c = const_cast<double*>(d_c_in);
  BT_TRY(btk_kalman_update(ctx, d_m, d_c, nullptr, nullptr, d_z, track_idx, meas_idx, d_f, k));
  BT_TRY(bt_unstage_out(ctx, mean, d_m, sizeof(double) * k * 8, loc));
  BT_TRY(bt_unstage_out(ctx, cov, d_c, sizeof(double) * k * 64, loc));
  return bt_finish(ctx, loc);
}

int32_t bt_kalman_project(bt_ctx* ctx, const double* mean, const double* cov, double* pmean, double* pcov,
                          int32_t n, int32_t loc) {
  BT_ENTER(ctx);
  BT_LOC_OK(loc);
  BT_CHECK(n >= 0, BT_ERR_INVALID, "n < 0");
  if (n == 0) return BT_OK;
  BT_CHECK(mean && cov && pmean && pcov, BT_ERR_INVALID, "NULL buffer");
  BT_TRY(bt_arena_reserve(ctx, loc == BT_HOST ? (size_t)n * (64 + 512 + 32 + 128) : 0));
  const double *d_m, *d_c; double *d_pm, *d_pc;
  BT_TRY(bt_in(ctx, mean, (size_t)n * 8, loc, &d_m));
  BT_TRY(bt_in(ctx, cov, (size_t)n * 64, loc, &d_c));
  BT_TRY(bt_out(ctx, pmean, (size_t)n * 4, loc, &d_pm));
  BT_TRY(bt_out(ctx, pcov, (size_t)n * 16, loc, &d_pc));
  BT_TRY(btk_kalman_project(ctx, d_m, d_c, d_pm, d_pc, n));
  BT_TRY(bt_unstage_out(ctx, pmean, d_pm, sizeof(double) * n * 4, loc));
  BT_TRY(bt_unstage_out(ctx, pcov, d_pc, sizeof(double) * n * 16, loc));
  return bt_finish(ctx, loc);
}

// ---- matching --------------------------------------------------------------------------------------
int32_t bt_iou_distance(bt_ctx* ctx, const double* a_tlbr, int32_t n, const double* b_tlbr, int32_t m,
                        double* out, int32_t loc) {
  BT_ENTER(ctx);
  BT_LOC_OK(loc);
  BT_CHECK(n >= 0 && m >= 0, BT_ERR_INVALID, "negative size");
  if (n == 0 || m == 0) return BT_OK;  // empty matrix (demo:1737-1739)
  BT_CHECK(a_tlbr && b_tlbr && out, BT_ERR_INVALID, "NULL buffer");
  BT_TRY(bt_arena_reserve(ctx, loc == BT_HOST ? 32 * ((size_t)n + m) + 8 * (size_t)n * m : 0));
  const double *d_a, *d_b; double* d_o;
  BT_TRY(bt_in(ctx, a_tlbr, (size_t)n * 4, loc, &d_a));
  BT_TRY(bt_in(ctx, b_tlbr, (size_t)m * 4, loc, &d_b));
  BT_TRY(bt_out(ctx, out, (size_t)n * m, loc, &d_o));
  BT_TRY(btk_iou_distance(ctx, d_a, n, d_b, m, d_o));
  BT_TRY(bt_unstage_out(ctx, out, d_o, sizeof(double) * n * m, loc));
  return bt_finish(ctx, loc);
}

int32_t bt_fuse_score(bt_ctx* ctx, const double* iou_dists, const double* det_scores, int32_t n, int32_t m,
                      double* out, int32_t loc) {
  BT_ENTER(ctx);
  BT_LOC_OK(loc);
  BT_CHECK(n >= 0 && m >= 0, BT_ERR_INVALID, "negative size");
  if (n == 0 || m == 0) return BT_OK;
  BT_CHECK(iou_dists && det_scores && out, BT_ERR_INVALID, "NULL buffer");
  BT_TRY(bt_arena_reserve(ctx, loc == BT_HOST ? 16 * (size_t)n * m + 8 * (size_t)m : 0));
  const double *d_d, *d_s; double* d_o;
  BT_TRY(bt_in(ctx, iou_dists, (size_t)n * m, loc, &d_d));
  BT_TRY(bt_in(ctx, det_scores, (size_t)m, loc, &d_s));
  BT_TRY(bt_out(ctx, out, (size_t)n * m, loc, &d_o));
  BT_TRY(btk_fuse_score(ctx, d_d, d_s, n, m, d_o));
  BT_TRY(bt_unstage_out(ctx, out, d_o, sizeof(double) * n * m, loc));
  return bt_finish(ctx, loc);
}

static int32_t assoc_dense_common(bt_ctx* ctx, const double* trk_tlbr, int32_t n, const double* det_tlbr,
                                  int32_t m, const float* a, const float* b, int32_t d, const float* face_sim,
                                  int32_t stage, float* out_emb, double* out_dists, int32_t precision,
                                  int32_t loc) {
  const size_t fa = (size_t)n * d, fb = (size_t)m * d, nm = (size_t)n * m;
  size_t need = 2 * (fa + fb) + 64;                       // fp16 copies
  if (loc == BT_HOST) need += 4 * (fa + fb) + 32 * ((size_t)n + m) + 4 * nm + 8 * nm + 4 * nm;
  BT_TRY(bt_arena_reserve(ctx, need));
  const float *d_a, *d_b, *d_face = nullptr; const double *d_rt = nullptr, *d_ct = nullptr;
  BT_TRY(bt_in(ctx, a, fa, loc, &d_a));
  BT_TRY(bt_in(ctx, b, fb, loc, &d_b));
  if (trk_tlbr) BT_TRY(bt_in(ctx, trk_tlbr, (size_t)n * 4, loc, &d_rt));
  if (det_tlbr) BT_TRY(bt_in(ctx, det_tlbr, (size_t)m * 4, loc, &d_ct));
  if (face_sim) BT_TRY(bt_in(ctx, face_sim, nm, loc, &d_face));
  float* d_emb = nullptr; double* d_dists = nullptr;
  if (out_emb) BT_TRY(bt_out(ctx, out_emb, nm, loc, &d_emb));
  if (out_dists) BT_TRY(bt_out(ctx, out_dists, nm, loc, &d_dists));
  bt_assoc_params p;
  memset(&p, 0, sizeof(p));
  p.count = 1;
  p.n[0] = n; p.m[0] = m; p.d = d;
  p.a_rows_alloc = n; p.b_rows_alloc = m;
  p.face_sim[0] = d_face;
  // fp16 operands round unit-norm rows of few components coarsely (d = 64: up to 5e-4 on the similarity):
  // below 512 components the CUDA-core fp32 kernel is both exact and fast enough
  if (precision == 0 && (d < 512 || d % 64 != 0)) precision = 1;
  if (precision == 1) { p.a32 = d_a; p.b32 = d_b; }
  if (precision == 0) {
    __half *a16, *b16;
    BT_TRY(bt_arena(ctx, fa, &a16));
    BT_TRY(bt_arena(ctx, fb, &b16));
    BT_TRY(btk_feature_prep(ctx, d_a, n, d, nullptr, a16, 0));
    BT_TRY(btk_feature_prep(ctx, d_b, m, d, nullptr, b16, 0));
    p.a16 = a16; p.b16 = b16;
  }
  p.row_tlbr = d_rt; p.col_tlbr = d_ct;
  bt_config cfg;
  bt_default_config(&cfg);
  p.match_thresh = cfg.match_thresh; p.second_thresh = cfg.second_thresh;
  p.unconf_thresh = cfg.unconfirmed_thresh; p.proximity = cfg.proximity_thresh;
  p.appearance = (float)cfg.appearance_thresh;
  p.cand.cnt = nullptr;
  p.out_emb = d_emb; p.out_dists = d_dists; p.dense_stage = stage;
  BT_TRY(btk_assoc(ctx, p, precision));
  if (out_emb) BT_TRY(bt_unstage_out(ctx, out_emb, d_emb, sizeof(float) * nm, loc));
  if (out_dists) BT_TRY(bt_unstage_out(ctx, out_dists, d_dists, sizeof(double) * nm, loc));
  return bt_finish(ctx, loc);
}

int32_t bt_embedding_distance(bt_ctx* ctx, const float* a, int32_t n, const float* b, int32_t m, int32_t d,
                              float* out, int32_t precision, int32_t loc) {
  BT_ENTER(ctx);
  BT_LOC_OK(loc);
  BT_CHECK(n >= 0 && m >= 0 && d > 0, BT_ERR_INVALID, "bad size");
  BT_CHECK(precision == 0 || precision == 1, BT_ERR_INVALID, "precision must be 0 or 1");
  if (n == 0 || m == 0) return BT_OK;
  BT_CHECK(a && b && out, BT_ERR_INVALID, "NULL buffer");
  return assoc_dense_common(ctx, nullptr, n, nullptr, m, a, b, d, nullptr, 1, out, nullptr, precision, loc);
}

int32_t bt_fused_cost(bt_ctx* ctx, const double* trk_tlbr, int32_t n, const double* det_tlbr, int32_t m,
                      const float* trk_feat, const float* det_feat, int32_t d, const float* face_sim,
                      int32_t stage, double* dists, int32_t precision, int32_t loc) {
  BT_ENTER(ctx);
  BT_LOC_OK(loc);
  BT_CHECK(n >= 0 && m >= 0 && d > 0, BT_ERR_INVALID, "bad size");
  BT_CHECK(stage == 1 || stage == 3, BT_ERR_INVALID, "stage must be 1 or 3");
  BT_CHECK(precision == 0 || precision == 1, BT_ERR_INVALID, "precision must be 0 or 1");
  if (n == 0 || m == 0) return BT_OK;
  BT_CHECK(trk_tlbr && det_tlbr && trk_feat && det_feat && dists, BT_ERR_INVALID, "NULL buffer");
  return assoc_dense_common(ctx, trk_tlbr, n, det_tlbr, m, trk_feat, det_feat, d, face_sim, stage, nullptr,
                            dists, precision, loc);
}

int32_t bt_linear_assignment(bt_ctx* ctx, const double* cost, int32_t n, int32_t m, double thresh, int32_t* x,
                             int32_t* y, int32_t loc) {
  BT_ENTER(ctx);
  BT_LOC_OK(loc);
  BT_CHECK(n >= 0 && m >= 0, BT_ERR_INVALID, "negative size");
  BT_CHECK(n <= ctx->max_tracks && m <= ctx->max_dets, BT_ERR_CAPACITY,
           "cost matrix %d x %d exceeds ctx capacity %d x %d", n, m, ctx->max_tracks, ctx->max_dets);
  if (n == 0 && m == 0) return BT_OK;
  BT_CHECK((n == 0 || x) && (m == 0 || y), BT_ERR_INVALID, "NULL output");
  BT_TRY(bt_arena_reserve(ctx, loc == BT_HOST ? 8 * (size_t)n * m + 4 * ((size_t)n + m) : 0));
  const double* d_cost = nullptr; int32_t *d_x = nullptr, *d_y = nullptr;
  if (n > 0 && m > 0) {
    BT_CHECK(cost != nullptr, BT_ERR_INVALID, "NULL cost");
    BT_TRY(bt_in(ctx, cost, (size_t)n * m, loc, &d_cost));
  }
  BT_TRY(bt_out(ctx, x, (size_t)n, loc, &d_x));
  BT_TRY(bt_out(ctx, y, (size_t)m, loc, &d_y));
  // video stream 0's candidate lists serve the stand-alone solver (its tracker keeps them zeroed between frames)
  const bt_cand& cand = *bt_lap_own_cand(ctx);
  BT_CUDA(cudaMemsetAsync(cand.cnt, 0, cand.clear_bytes, ctx->stream));
  BT_TRY(btk_lap_compact_dense(ctx, d_cost, n, m, thresh, cand, 0));
  BT_TRY(btk_lap_solve(ctx, cand, 0, n, m, thresh, d_x, d_y));
  // leave the candidate counters zeroed: the tracker path relies on it (its LAP kernel clears its own)
  BT_CUDA(cudaMemsetAsync(cand.cnt, 0, cand.clear_bytes, ctx->stream));
  BT_TRY(bt_unstage_out(ctx, x, d_x, sizeof(int32_t) * n, loc));
  BT_TRY(bt_unstage_out(ctx, y, d_y, sizeof(int32_t) * m, loc));
  return bt_finish(ctx, loc);
}

int32_t bt_feature_ema(bt_ctx* ctx, float* smooth, float* curr, const float* feat, const int32_t* track_idx,
                       const int32_t* feat_idx, const uint8_t* first, int32_t k, int32_t d, float alpha,
                       int32_t loc) {
  BT_ENTER(ctx);
  BT_LOC_OK(loc);
  BT_CHECK(k >= 0 && d > 0, BT_ERR_INVALID, "bad size");
  if (k == 0) return BT_OK;
  BT_CHECK(smooth && curr && feat, BT_ERR_INVALID, "NULL buffer");
  BT_CHECK(loc == BT_DEVICE || (track_idx == nullptr && feat_idx == nullptr), BT_ERR_INVALID,
           "index lists are only supported for device buffers");
  const size_t kd = (size_t)k * d;
  BT_TRY(bt_arena_reserve(ctx, loc == BT_HOST ? 12 * kd + k : 0));
  const float *d_s_in, *d_c_in, *d_f; const uint8_t* d_first = nullptr;
  BT_TRY(bt_in(ctx, (const float*)smooth, kd, loc, &d_s_in));
  BT_TRY(bt_in(ctx, (const float*)curr, kd, loc, &d_c_in));
  BT_TRY(bt_in(ctx, feat, kd, loc, &d_f));
  if (first) BT_TRY(bt_in(ctx, first, (size_t)k, loc, &d_first));
  float* d_s = const_cast<float*>(d_s_in);
  float* d_c = const_cast<float*>(d_c_in);
  // alpha and 1 - alpha are rounded to float32 separately (NEP 50 weak Python floats, demo:473, demo:499-501)
  const double alpha_d = (alpha == 0.9f) ? 0.9 : (double)alpha;
  BT_TRY(btk_feature_ema(ctx, d_s, d_c, d_f, track_idx, feat_idx, d_first, k, d, (float)alpha_d, (float)(1.0 - alpha_d)));
  BT_TRY(bt_unstage_out(ctx, smooth, d_s, sizeof(float) * kd, loc));
  BT_TRY(bt_unstage_out(ctx, curr, d_c, sizeof(float) * kd, loc));
  return bt_finish(ctx, loc);
}

}  // extern "C"
