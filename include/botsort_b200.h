/*
 * botsort_b200 -- C ABI of the B200-native BoT-SORT per-frame tracking hot path.
 *
 * This is the drop-in boundary.  The reference
 * (PINTO0309/BoT-SORT-ONNX-TensorRT, one Python file, cited below as demo:LINE =
 * demo_bottrack_onnx_tflite.py:LINE) has no FFI of its own: the path sits behind Python
 * classes/functions.  Every entry point below names the reference interface it replaces;
 * the ctypes binding a maintainer would add is shown in INTEGRATION.md and shipped in
 * bot-sort-onnx-tensorrt_b200/_lib.py.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no exceptions or aborts cross the boundary.
 *   - every function returns int32 status: BT_OK (0) or a negative bt_status; the message
 *     of the last failure is bt_last_error(ctx) (ctx may be NULL for create failures).
 *   - `loc` says where EVERY pointer argument of that call lives: BT_HOST (pageable or
 *     pinned host memory; the library stages through its own pinned/device workspaces and
 *     the call returns after the results are back in the host buffers) or BT_DEVICE (device
 *     memory on the ctx's device; work is enqueued on the ctx stream, the call returns
 *     without synchronising -- use bt_sync or bt_stream).
 *   - the caller owns every input/output buffer; a ctx is not re-entrant (one per host
 *     thread); different ctxs are independent.  No CPU fallback exists: without a CUDA
 *     device bt_create fails with BT_ERR_CUDA.
 *   - matrices are row-major and dense; boxes are tlbr = x1,y1,x2,y2.
 */
#ifndef BOTSORT_B200_H_
#define BOTSORT_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BT_VERSION 200 /* 0.2.0: multi-stream contexts, fp16 feature ingest, submit/step split */

typedef struct bt_ctx bt_ctx;

typedef enum bt_status {
  BT_OK = 0,
  BT_ERR_INVALID = -1,  /* bad argument (maps to ValueError in the Python mirror) */
  BT_ERR_CUDA = -2,     /* CUDA runtime/driver failure (RuntimeError) */
  BT_ERR_CAPACITY = -3, /* problem larger than the ctx was created for (ValueError) */
  BT_ERR_STATE = -4     /* call sequence error (RuntimeError) */
} bt_status;

enum { BT_HOST = 0, BT_DEVICE = 1 };

/* bt_create flags */
enum {
  BT_FLAG_SIMT_SIM = 1u << 0,     /* ReID similarity on the CUDA-core kernel instead of tcgen05 */
  BT_FLAG_NO_F32_FEATURES = 1u << 1 /* do not keep the fp32 smooth (EMA) feature bank (A10 exposed state) */
};

/* dtype of the ReID feature rows handed to the tracker */
enum {
  BT_F32 = 0, /* float32 rows (what onnxruntime returns, demo:829-837) */
  BT_F16 = 1  /* float16 rows (what the reference's fp16 TensorRT FastReID engine computes, demo:738, demo:35-49):
                 half the bytes over PCIe / NVLink and no conversion pass; semantics = the reference fed with
                 feats.astype(float32) */
};

/* most video streams one launch can serve (bt_update_streams batch size) */
#define BT_MAX_BATCH 32

/* Track states, demo:382-387 */
enum { BT_STATE_NEW = 0, BT_STATE_TRACKED = 1, BT_STATE_LOST = 2, BT_STATE_LONGLOST = 3, BT_STATE_REMOVED = 4 };

/* Tracker hyper-parameters; defaults = the constants hard-coded at demo:1268-1277, demo:1571,
 * demo:1604, demo:1667, demo:473 (bt_default_config fills them). */
typedef struct bt_config {
  /* score thresholds are Python floats (doubles) in the reference and are compared with float(score)
   * (demo:1022, demo:1501, demo:1531, demo:1617): kept as doubles so that a score equal to float32(0.4)
   * lands on the same side as in the reference */
  double track_high_thresh; /* 0.40  demo:1268 */
  double track_low_thresh;  /* 0.10  demo:1269 */
  double new_track_thresh;  /* 0.90  demo:1270 */
  double match_thresh;      /* 0.80  demo:1271, first association */
  double second_thresh;     /* 0.50  demo:1571 */
  double unconfirmed_thresh;/* 0.70  demo:1604 */
  double proximity_thresh;  /* 0.50  demo:1274 */
  double appearance_thresh; /* 0.25  demo:1275 (compared with float32 distances: weak Python float -> float32) */
  double duplicate_iou_dist;/* 0.15  demo:1667 */
  double ema_alpha;         /* 0.9   demo:473 (alpha and 1 - alpha are rounded to float32 separately, NEP 50) */
  int32_t track_buffer;     /* 300   demo:1272 */
  int32_t frame_rate;       /* 30    demo:1256 */
  int32_t with_reid;        /* 1: features are given every frame; 0: IoU-only (similarities 0) */
  int32_t reserved;
} bt_config;

int32_t bt_version(void);
const char* bt_last_error(const bt_ctx* ctx);
void bt_default_config(bt_config* cfg);

/* One ctx = one CUDA device + one CUDA stream + workspaces + n_streams trackers (video streams).
 * max_tracks bounds live track slots (tracked + lost + unconfirmed) PER video stream, max_dets the
 * detections per frame and stream, feat_dim the ReID feature size (2048 for Fast-ReID, demo:1060).
 * The video streams of one ctx are a leading batch dimension of every kernel: one
 * bt_update_streams call steps any subset of them with ONE launch per kernel.  The reference runs one
 * tracker per process (global id counter, demo:390, demo:1264); here ids are per video stream.
 * bt_create == bt_create_streams with n_streams = 1. */
int32_t bt_create(int32_t device, int32_t max_tracks, int32_t max_dets, int32_t feat_dim,
                  uint32_t flags, bt_ctx** out);
int32_t bt_create_streams(int32_t device, int32_t n_streams, int32_t max_tracks, int32_t max_dets,
                          int32_t feat_dim, uint32_t flags, bt_ctx** out);
int32_t bt_num_streams(const bt_ctx* ctx);
int32_t bt_destroy(bt_ctx* ctx);
int32_t bt_sync(bt_ctx* ctx);
/* cudaStream_t of the ctx as an opaque pointer (so callers can record CUDA events on it). */
void* bt_stream(bt_ctx* ctx);
/* Number of kernels this library launched on the ctx since creation (bench `gpu_launches`). */
int64_t bt_launch_count(const bt_ctx* ctx);

/* Segment timing of bt_update_arrays with CUDA events recorded on the ctx stream (what
 * bench.py's roofline figures are computed from).  Off by default. */
enum {
  BT_SEG_PREP = 0,    /* detection prep + feature normalisation/fp16 staging */
  BT_SEG_PREDICT = 1, /* batched Kalman predict */
  BT_SEG_ASSOC = 2,   /* fused association kernel (ReID GEMM + IoU + cost fusion + candidate emission) */
  BT_SEG_LAP = 3,     /* the three LAP solves */
  BT_SEG_UPDATE = 4,  /* Kalman update + initiate + feature EMA */
  BT_SEG_DUP = 5,     /* duplicate test + result gather */
  /* host wall-clock phases of bt_update_arrays (std::chrono), same units (ms) */
  BT_SEG_HOST_ENQUEUE1 = 6, /* input staging + enqueue up to the LAP */
  BT_SEG_HOST_WAIT1 = 7,    /* first stream synchronisation */
  BT_SEG_HOST_LISTS = 8,    /* list bookkeeping + enqueue of update/duplicate kernels */
  BT_SEG_HOST_WAIT2 = 9,    /* second stream synchronisation */
  BT_SEG_HOST_FINAL = 10,   /* duplicate resolution, slot recycling */
  BT_SEG_COUNT = 11
};
int32_t bt_profile_enable(bt_ctx* ctx, int32_t on);
/* Re-launches the fused association kernel of the LAST bt_update_arrays frame `iters` times back to
 * back on the ctx stream (same operands; the tensor-core kernel's candidate emission is idempotent)
 * inside one CUDA-event pair and returns the elapsed time in *total_ms (per launch: / iters).  Measurement aid for the kernel's
 * roofline figure: back-to-back launches are not inflated by host enqueue gaps.  The tracker state is
 * not changed (the candidate lists are left cleared, as the frame step leaves them). */
int32_t bt_profile_replay_assoc(bt_ctx* ctx, int32_t iters, double* total_ms);
/* accumulated device milliseconds and number of samples of a segment since bt_profile_enable(1) */
int32_t bt_profile_read(bt_ctx* ctx, int32_t segment, double* total_ms, int64_t* samples);

/* ---- Kalman filter (replaces KalmanFilter, demo:118-336) ------------------------------- */
/* KalmanFilter.initiate, demo:166-197: xywh[k,4] float32 -> mean[k,8], cov[k,64] float64.
 * The float32 rounding NumPy>=2 applies to the initial std/variance (float32 measurement,
 * NEP-50 weak Python floats) is reproduced. */
int32_t bt_kalman_initiate(bt_ctx* ctx, const float* xywh, double* mean, double* cov, int32_t k,
                           int32_t loc);
/* KalmanFilter.multi_predict (demo:265-302) fused with the velocity reset of
 * STrack.multi_predict (demo:529-532): in place on mean[n,8], cov[n,64].
 * state: int32[n] track states or NULL; rows with state != BT_STATE_TRACKED get
 * mean[6]=mean[7]=0 before the prediction.  noise_f32 != 0 evaluates the process noise in
 * float32 (what NumPy does when every pooled mean is still float32, i.e. on frame 2). */
int32_t bt_kalman_multi_predict(bt_ctx* ctx, double* mean, double* cov, const int32_t* state,
                                int32_t n, int32_t noise_f32, int32_t loc);
/* KalmanFilter.update (demo:304-336) batched over k (track, measurement) pairs:
 * track_idx[k] rows of mean/cov (NULL = 0..k-1) are updated in place with meas[meas_idx[k]]
 * (meas[.,4] float64 xywh; meas_idx NULL = 0..k-1).  noise_f32: uint8[k] or NULL, rows whose
 * projection noise NumPy evaluates in float32 (never-predicted float32 states). */
int32_t bt_kalman_update(bt_ctx* ctx, double* mean, double* cov, const double* meas,
                         const int32_t* track_idx, const int32_t* meas_idx, const uint8_t* noise_f32,
                         int32_t k, int32_t loc);
/* KalmanFilter.project (demo:236-263) batched: mean[n,8], cov[n,64] -> pmean[n,4], pcov[n,16]. */
int32_t bt_kalman_project(bt_ctx* ctx, const double* mean, const double* cov, double* pmean,
                          double* pcov, int32_t n, int32_t loc);

/* ---- matching primitives (replace demo:1682-1761 and the in-graph ReID cosine) ---------- */
/* iou_distance / bbox_ious, demo:1731-1761: out[n,m] = 1 - IoU (float64; strict `<=` empty
 * rule, no +1 pixel convention, demo:1702). */
int32_t bt_iou_distance(bt_ctx* ctx, const double* a_tlbr, int32_t n, const double* b_tlbr,
                        int32_t m, double* out, int32_t loc);
/* embedding distance, demo:1599 (and the in-graph cosine consumed at demo:1453-1460):
 * out[n,m] = 1 - max(0, A[n,d] . B[m,d]^T), float32.
 * precision 0: tcgen05 tensor cores, fp16 operands, fp32 accumulate (d % 64 == 0 required);
 * precision 1: fp32 CUDA-core kernel. */
int32_t bt_embedding_distance(bt_ctx* ctx, const float* a, int32_t n, const float* b, int32_t m,
                              int32_t d, float* out, int32_t precision, int32_t loc);
/* Fused association cost.  stage 1 = demo:1539-1554 (appearance overrides IoU),
 * stage 3 = demo:1599-1602 (unconfirmed tracks: appearance gate then IoU gate):
 * dists[n,m] float64 from boxes + features; face_sim[n,m] float32 or NULL (= 0, the face
 * encoder is out of scope).  trk_feat/det_feat float32 [.,d] unit-norm rows. */
int32_t bt_fused_cost(bt_ctx* ctx, const double* trk_tlbr, int32_t n, const double* det_tlbr,
                      int32_t m, const float* trk_feat, const float* det_feat, int32_t d,
                      const float* face_sim, int32_t stage, double* dists, int32_t precision,
                      int32_t loc);
/* fuse_score (upstream BoT-SORT name, absent from the reference):
 * out = 1 - (1 - iou_dists) * det_scores[None, :]. */
int32_t bt_fuse_score(bt_ctx* ctx, const double* iou_dists, const double* det_scores, int32_t n,
                      int32_t m, double* out, int32_t loc);
/* linear_assignment, demo:1682-1693 == lap.lapjv(cost, extend_cost=True, cost_limit=thresh):
 * exact minimiser of sum over matched pairs of (cost - thresh); x[n] column of each row or -1,
 * y[m] row of each column or -1. */
int32_t bt_linear_assignment(bt_ctx* ctx, const double* cost, int32_t n, int32_t m, double thresh,
                             int32_t* x, int32_t* y, int32_t loc);
/* STrack.update_body_features, demo:492-502, batched: for i<k, row t=track_idx[i] gets
 * curr[t] = feat[feat_idx[i]]; smooth[t] = normalise(alpha*smooth[t] + (1-alpha)*feat[...])
 * (first[i] != 0: smooth[t] = normalise(feat[...]) -- the first call, demo:497-498). */
int32_t bt_feature_ema(bt_ctx* ctx, float* smooth, float* curr, const float* feat,
                       const int32_t* track_idx, const int32_t* feat_idx, const uint8_t* first,
                       int32_t k, int32_t d, float alpha, int32_t loc);

/* ---- detector side (replaces YOLOX in-graph decode+NMS + YOLOX._postprocess demo:968-1030,
 *      and the crop + FastReID._preprocess of demo:1434-1436, demo:1101-1142) ------------- */
typedef struct bt_yolox_config {
  int32_t in_h, in_w;         /* model input size (480, 640) */
  int32_t img_h, img_w;       /* original frame size */
  int32_t num_classes;        /* 4: body, head, hand, face (demo:1304-1370) */
  float nms_score_thresh;     /* 0.15 in-graph, model file name demo:34 */
  float nms_iou_thresh;       /* 0.80 */
  int32_t max_per_class;      /* 50 */
  float post_score_thresh;    /* 0.35 demo:862 */
} bt_yolox_config;
void bt_default_yolox_config(bt_yolox_config* cfg);
/* raw_head[anchors, 5+num_classes] float32 (cx,cy,w,h logits per stride 8/16/32, obj logit,
 * class logits); out_boxes[max_out,6] int32/float packed as float64 rows
 * (classid, score, x1, y1, x2, y2 in image pixels, truncated like demo:1009-1012),
 * ordered by class then descending score; *out_count rows written. */
int32_t bt_yolox_postprocess(bt_ctx* ctx, const float* raw_head, const bt_yolox_config* cfg,
                             double* out_boxes, int32_t max_out, int32_t* out_count, int32_t loc);
/* frame uint8 [h,w,3] BGR; boxes int32[n,4] tlbr; out float32 [n,3,out_h,out_w] RGB,
 * cv2.resize INTER_LINEAR bit-exact then (x/255-mean)/std. */
int32_t bt_reid_crop_gather(bt_ctx* ctx, const uint8_t* frame, int32_t h, int32_t w,
                            const int32_t* boxes, int32_t n, int32_t out_h, int32_t out_w,
                            float* out, int32_t loc);

/* Device-chained detector side (the reference bounces every model output through host NumPy,
 * demo:829-837, demo:916-918, demo:1096-1098).  All pointers are DEVICE pointers, everything is enqueued on the ctx
 * stream, nothing is synchronised:
 *   raw YOLOX head -> decode + NMS + _postprocess (as bt_yolox_postprocess)
 *     -> the class-0 (body) detections land in the input buffers of the NEXT bt_submit_streams of `stream_id`
 *        (bt_input_buffers: boxes, scores); rows past the last body get score 0, which every association stage
 *        ignores, so the tracker can be stepped with m = cfg->max_per_class and no count visits the host
 *     -> their ReID crops (as bt_reid_crop_gather) in `crops` float32[max_per_class, 3, out_h, out_w] -- the input
 *        binding of the ReID engine, whose fp16 output binding is the feats16 buffer of bt_input_buffers.
 * det_out / det_count: optional device float64[max_out,6] + int32 with the whole decoded list (all classes), e.g. for
 * the host-side body-part grouping (demo:1372-1411).  Then:
 *   bt_update_streams(ctx, 1, &stream_id, &boxes, &scores, &feats16, &max_per_class, BT_F16, NULL, BT_DEVICE, info)
 * with the bt_input_buffers pointers. */
int32_t bt_detect_stage(bt_ctx* ctx, int32_t stream_id, const float* raw_head, const bt_yolox_config* cfg,
                        const uint8_t* frame, int32_t h, int32_t w, int32_t out_h, int32_t out_w, float* crops,
                        double* det_out, int32_t max_out, int32_t* det_count);

/* ---- the tracker (replaces BoTSORT.update demo:1291-1639 driven by arrays) --------------- */
/* Resets video stream 0 / one video stream / all of them (stream_id < 0). */
int32_t bt_tracker_reset(bt_ctx* ctx, const bt_config* cfg /* NULL = defaults */);
int32_t bt_tracker_reset_stream(bt_ctx* ctx, int32_t stream_id, const bt_config* cfg);

typedef struct bt_frame_info {
  int32_t frame_id;
  int32_t n_tracked;     /* len(self.tracked_stracks) after the frame (the returned list) */
  int32_t n_lost;
  int32_t n_removed_total;
  int32_t n_pool;        /* tracks x detections entering the first association */
  int32_t n_high;
  int32_t n_low;
  int32_t n_unconfirmed;
  int32_t n_matches1, n_matches2, n_matches3;
  int32_t n_births;
  int32_t n_births_skipped; /* births dropped because the track store of the stream was full (the reference
                               has no bound; the frame stays consistent) */
  int32_t reserved[3];
} bt_frame_info;

/* One BoTSORT.update on detector/encoder outputs of video stream 0: boxes int32[m,4] tlbr (as
 * YOLOX._postprocess emits them), scores float32[m], feats float32[m,feat_dim] (NULL when with_reid == 0).
 * All class 0 (body).  info may be NULL. */
int32_t bt_update_arrays(bt_ctx* ctx, const int32_t* boxes, const float* scores, const float* feats,
                         int32_t m, int32_t loc, bt_frame_info* info);

/* The same for `count` video streams of the ctx at once (stream_ids[k] distinct, count <= BT_MAX_BATCH):
 * boxes[k] / scores[k] / feats[k] / m[k] are stream stream_ids[k]'s detections; feats rows are
 * feat_dtype (BT_F32 / BT_F16); face_sims is NULL or an array whose entries are NULL or the face
 * similarity matrix float32[n_pool, m[k]] of that stream (demo:1465-1486: rows in pool order = activated
 * tracked tracks in list order, then lost tracks in list order; the term `face_emb_dists` of demo:1541-1546).
 * infos: NULL or bt_frame_info[count].  == bt_submit_streams + bt_step_streams. */
int32_t bt_update_streams(bt_ctx* ctx, int32_t count, const int32_t* stream_ids, const int32_t* const* boxes,
                          const float* const* scores, const void* const* feats, const int32_t* m,
                          int32_t feat_dtype, const float* const* face_sims, int32_t loc, bt_frame_info* infos);

/* Split form for pipelining: bt_submit_streams starts moving one frame of inputs into the ctx's
 * (double-buffered) association buffers on a copy stream and returns; bt_step_streams runs the frame step
 * of the OLDEST submitted frame of each listed stream and returns when the results are on the host.  A
 * stream may have at most two submitted-but-not-stepped frames, so
 *     submit(f0); loop { submit(f[k+1]); step() -> results of f[k]; }
 * overlaps frame k+1's host->device copy with frame k's association and bookkeeping.
 * Host feature buffers must stay valid until the step of that frame returns (pinned memory makes the copy
 * asynchronous); boxes and scores are copied at once. */
int32_t bt_submit_streams(bt_ctx* ctx, int32_t count, const int32_t* stream_ids, const int32_t* const* boxes,
                          const float* const* scores, const void* const* feats, const int32_t* m,
                          int32_t feat_dtype, const float* const* face_sims, int32_t loc);
int32_t bt_step_streams(bt_ctx* ctx, int32_t count, const int32_t* stream_ids, bt_frame_info* infos);

/* Zero-copy ingest: device pointers of the buffers the NEXT bt_submit_streams of this
 * video stream reads its inputs from when it is called with loc == BT_DEVICE and exactly these pointers --
 * a detector / ReID engine that writes its outputs here (bt_yolox_postprocess + bt_reid_crop_gather ->
 * encoder -> feats16) hands them over without any copy.  boxes int32[max_dets,4], scores
 * float32[max_dets], feats16 float16[max_dets, feat_dim]. */
int32_t bt_input_buffers(bt_ctx* ctx, int32_t stream_id, int32_t** boxes, float** scores, void** feats16);

/* Read back a track list after a frame: which = 0 tracked_stracks (the list update() returns,
 * in the reference's order), 1 lost_stracks.  Any output pointer may be NULL.  Host pointers only.
 * ids/state/activated/frame_id/start_frame/tracklet_len/det_index: int32[n]; score float32[n];
 * tlbr float64[n,4]; mean float64[n,8]; cov float64[n,64].  Returns the list length in *n
 * (capacity `cap` rows; BT_ERR_CAPACITY if smaller).  bt_get_tracks == stream 0. */
int32_t bt_get_tracks(bt_ctx* ctx, int32_t which, int32_t cap, int32_t* n, int32_t* ids,
                      int32_t* state, int32_t* activated, int32_t* frame_id, int32_t* start_frame,
                      int32_t* tracklet_len, int32_t* det_index, float* score, double* tlbr,
                      double* mean, double* cov);
int32_t bt_get_tracks_stream(bt_ctx* ctx, int32_t stream_id, int32_t which, int32_t cap, int32_t* n, int32_t* ids,
                             int32_t* state, int32_t* activated, int32_t* frame_id, int32_t* start_frame,
                             int32_t* tracklet_len, int32_t* det_index, float* score, double* tlbr,
                             double* mean, double* cov);
/* fp32 feature state of a list (A10 exposed state): curr/smooth float32[n,feat_dim]. */
int32_t bt_get_track_features(bt_ctx* ctx, int32_t which, int32_t cap, float* curr, float* smooth);
int32_t bt_get_track_features_stream(bt_ctx* ctx, int32_t stream_id, int32_t which, int32_t cap, float* curr,
                                     float* smooth);
/* Per-frame intermediates of the last frame step, for parity tests:
 * stage 1/2/3 matches as (track list index, detection list index) pairs in the reference's
 * index spaces (demo:1556, demo:1571, demo:1604); pairs int32[cap,2]. */
int32_t bt_get_matches(bt_ctx* ctx, int32_t stage, int32_t cap, int32_t* n, int32_t* pairs);
int32_t bt_get_matches_stream(bt_ctx* ctx, int32_t stream_id, int32_t stage, int32_t cap, int32_t* n, int32_t* pairs);

#ifdef __cplusplus
}
#endif
#endif /* BOTSORT_B200_H_ */
