"""ctypes binding of libbotsort_b200.so (C ABI: include/botsort_b200.h).

This is the only place the Python mirror touches native code.  There is no CPU fallback: if
the shared library is missing or no sm_100 device is present, loading / `Context()` raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libbotsort_b200.so")

BT_HOST, BT_DEVICE = 0, 1
BT_OK, BT_ERR_INVALID, BT_ERR_CUDA, BT_ERR_CAPACITY, BT_ERR_STATE = 0, -1, -2, -3, -4
BT_FLAG_SIMT_SIM = 1 << 0
BT_FLAG_NO_F32_FEATURES = 1 << 1
BT_F32, BT_F16 = 0, 1
BT_MAX_BATCH = 32

# every symbol include/botsort_b200.h declares (checked by tests/test_abi.py)
EXPORTS = [
    "bt_version", "bt_last_error", "bt_default_config", "bt_create", "bt_create_streams", "bt_num_streams",
    "bt_destroy", "bt_sync", "bt_stream",
    "bt_launch_count", "bt_kalman_initiate", "bt_kalman_multi_predict", "bt_kalman_update", "bt_kalman_project",
    "bt_iou_distance", "bt_embedding_distance", "bt_fused_cost", "bt_fuse_score", "bt_linear_assignment",
    "bt_feature_ema", "bt_default_yolox_config", "bt_yolox_postprocess", "bt_reid_crop_gather", "bt_detect_stage",
    "bt_tracker_reset", "bt_tracker_reset_stream", "bt_update_arrays", "bt_update_streams", "bt_submit_streams",
    "bt_step_streams", "bt_input_buffers", "bt_get_tracks", "bt_get_tracks_stream", "bt_get_track_features",
    "bt_get_track_features_stream", "bt_get_matches", "bt_get_matches_stream",
    "bt_profile_enable", "bt_profile_read", "bt_profile_replay_assoc",
]
SEGMENTS = ("prep", "predict", "assoc", "lap", "update", "dup",
            "host_enqueue1", "host_wait1", "host_lists", "host_wait2", "host_final")


class BtConfig(C.Structure):
    _fields_ = [
        ("track_high_thresh", C.c_double), ("track_low_thresh", C.c_double), ("new_track_thresh", C.c_double),
        ("match_thresh", C.c_double), ("second_thresh", C.c_double), ("unconfirmed_thresh", C.c_double),
        ("proximity_thresh", C.c_double), ("appearance_thresh", C.c_double), ("duplicate_iou_dist", C.c_double),
        ("ema_alpha", C.c_double),
        ("track_buffer", C.c_int32), ("frame_rate", C.c_int32), ("with_reid", C.c_int32), ("reserved", C.c_int32),
    ]


class BtYoloxConfig(C.Structure):
    _fields_ = [
        ("in_h", C.c_int32), ("in_w", C.c_int32), ("img_h", C.c_int32), ("img_w", C.c_int32),
        ("num_classes", C.c_int32), ("nms_score_thresh", C.c_float), ("nms_iou_thresh", C.c_float),
        ("max_per_class", C.c_int32), ("post_score_thresh", C.c_float),
    ]


class BtFrameInfo(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "frame_id", "n_tracked", "n_lost", "n_removed_total", "n_pool", "n_high", "n_low", "n_unconfirmed",
        "n_matches1", "n_matches2", "n_matches3", "n_births", "n_births_skipped")] + [("reserved", C.c_int32 * 3)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_ if n != "reserved"}


_lib = None


def load_library(path: Optional[str] = None) -> C.CDLL:
    """dlopen the shared library and declare the signatures (no CUDA call is made)."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise RuntimeError(
            f"{p} not found: build it with `python bot-sort-onnx-tensorrt_b200/build.py` "
            "(nvcc, sm_100a). There is no CPU fallback.")
    lib = C.CDLL(p)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    lib.bt_version.restype = i32
    lib.bt_last_error.restype = C.c_char_p
    lib.bt_last_error.argtypes = [vp]
    lib.bt_default_config.restype = None
    lib.bt_default_config.argtypes = [C.POINTER(BtConfig)]
    lib.bt_default_yolox_config.restype = None
    lib.bt_default_yolox_config.argtypes = [C.POINTER(BtYoloxConfig)]
    lib.bt_create.restype = i32
    lib.bt_create.argtypes = [i32, i32, i32, i32, C.c_uint32, C.POINTER(vp)]
    lib.bt_create_streams.restype = i32
    lib.bt_create_streams.argtypes = [i32, i32, i32, i32, i32, C.c_uint32, C.POINTER(vp)]
    lib.bt_num_streams.restype = i32
    lib.bt_num_streams.argtypes = [vp]
    lib.bt_destroy.restype = i32
    lib.bt_destroy.argtypes = [vp]
    lib.bt_sync.restype = i32
    lib.bt_sync.argtypes = [vp]
    lib.bt_stream.restype = vp
    lib.bt_stream.argtypes = [vp]
    lib.bt_launch_count.restype = i64
    lib.bt_launch_count.argtypes = [vp]
    sigs = {
        "bt_kalman_initiate": [vp, vp, vp, vp, i32, i32],
        "bt_kalman_multi_predict": [vp, vp, vp, vp, i32, i32, i32],
        "bt_kalman_update": [vp, vp, vp, vp, vp, vp, vp, i32, i32],
        "bt_kalman_project": [vp, vp, vp, vp, vp, i32, i32],
        "bt_iou_distance": [vp, vp, i32, vp, i32, vp, i32],
        "bt_embedding_distance": [vp, vp, i32, vp, i32, i32, vp, i32, i32],
        "bt_fused_cost": [vp, vp, i32, vp, i32, vp, vp, i32, vp, i32, vp, i32, i32],
        "bt_fuse_score": [vp, vp, vp, i32, i32, vp, i32],
        "bt_linear_assignment": [vp, vp, i32, i32, C.c_double, vp, vp, i32],
        "bt_feature_ema": [vp, vp, vp, vp, vp, vp, vp, i32, i32, C.c_float, i32],
        "bt_yolox_postprocess": [vp, vp, C.POINTER(BtYoloxConfig), vp, i32, vp, i32],
        "bt_reid_crop_gather": [vp, vp, i32, i32, vp, i32, i32, i32, vp, i32],
        "bt_detect_stage": [vp, i32, vp, C.POINTER(BtYoloxConfig), vp, i32, i32, i32, i32, vp, vp, i32, vp],
        "bt_tracker_reset": [vp, C.POINTER(BtConfig)],
        "bt_tracker_reset_stream": [vp, i32, C.POINTER(BtConfig)],
        "bt_update_arrays": [vp, vp, vp, vp, i32, i32, C.POINTER(BtFrameInfo)],
        "bt_update_streams": [vp, i32, vp, vp, vp, vp, vp, i32, vp, i32, vp],
        "bt_submit_streams": [vp, i32, vp, vp, vp, vp, vp, i32, vp, i32],
        "bt_step_streams": [vp, i32, vp, vp],
        "bt_input_buffers": [vp, i32, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)],
        "bt_get_tracks": [vp, i32, i32, vp] + [vp] * 11,
        "bt_get_tracks_stream": [vp, i32, i32, i32, vp] + [vp] * 11,
        "bt_get_track_features": [vp, i32, i32, vp, vp],
        "bt_get_track_features_stream": [vp, i32, i32, i32, vp, vp],
        "bt_get_matches": [vp, i32, i32, vp, vp],
        "bt_get_matches_stream": [vp, i32, i32, i32, vp, vp],
        "bt_profile_enable": [vp, i32],
        "bt_profile_read": [vp, i32, C.POINTER(C.c_double), C.POINTER(C.c_int64)],
        "bt_profile_replay_assoc": [vp, i32, C.POINTER(C.c_double)],
    }
    for name, args in sigs.items():
        fn = getattr(lib, name)
        fn.restype = i32
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _arr(a, dtype, shape=None) -> np.ndarray:
    out = np.ascontiguousarray(a, dtype=dtype)
    if shape is not None:
        out = out.reshape(shape)
    return out


class BotsortError(RuntimeError):
    pass


class Context:
    """One bt_ctx: a CUDA device + stream + workspaces + one tracker.  Host-buffer (NumPy) API."""

    def __init__(self, max_tracks: int = 4096, max_dets: int = 4096, feat_dim: int = 2048, device: int = 0,
                 flags: int = 0, n_streams: int = 1):
        self.lib = load_library()
        h = C.c_void_p()
        st = self.lib.bt_create_streams(device, n_streams, max_tracks, max_dets, feat_dim, flags, C.byref(h))
        if st != BT_OK:
            msg = self.lib.bt_last_error(None).decode()
            raise BotsortError(f"bt_create failed ({st}): {msg}")
        self.h = h
        self.max_tracks, self.max_dets, self.feat_dim, self.device = max_tracks, max_dets, feat_dim, device
        self.n_streams = n_streams

    def close(self):
        if getattr(self, "h", None):
            self.lib.bt_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, st: int):
        if st == BT_OK:
            return
        msg = self.lib.bt_last_error(self.h).decode()
        if st in (BT_ERR_INVALID, BT_ERR_CAPACITY):
            raise ValueError(msg)
        raise BotsortError(f"status {st}: {msg}")

    # ---- misc ----
    def sync(self):
        self._check(self.lib.bt_sync(self.h))

    @property
    def stream(self) -> int:
        return int(self.lib.bt_stream(self.h) or 0)

    @property
    def launch_count(self) -> int:
        return int(self.lib.bt_launch_count(self.h))

    def profile_enable(self, on: bool = True):
        self._check(self.lib.bt_profile_enable(self.h, int(on)))

    def profile_replay_assoc(self, iters: int = 50) -> float:
        """Average device ms of one launch of the last frame's association kernel (back-to-back replays)."""
        ms = C.c_double(0)
        self._check(self.lib.bt_profile_replay_assoc(self.h, iters, C.byref(ms)))
        return ms.value / max(1, iters)

    def profile_read(self) -> dict:
        """{segment: (total device ms, samples)} accumulated since profile_enable(True)."""
        out = {}
        for i, name in enumerate(SEGMENTS):
            ms, n = C.c_double(0), C.c_int64(0)
            self._check(self.lib.bt_profile_read(self.h, i, C.byref(ms), C.byref(n)))
            out[name] = (ms.value, n.value)
        return out

    # ---- Kalman ----
    def kalman_initiate(self, xywh):
        z = _arr(xywh, np.float32, (-1, 4))
        k = z.shape[0]
        mean = np.empty((k, 8), np.float64)
        cov = np.empty((k, 8, 8), np.float64)
        self._check(self.lib.bt_kalman_initiate(self.h, _ptr(z), _ptr(mean), _ptr(cov), k, BT_HOST))
        return mean, cov

    def kalman_multi_predict(self, mean, cov, state=None, noise_f32=False):
        m = np.array(mean, dtype=np.float64, order="C").reshape(-1, 8)
        c = np.array(cov, dtype=np.float64, order="C").reshape(-1, 8, 8)
        s = None if state is None else _arr(state, np.int32, (-1,))
        self._check(self.lib.bt_kalman_multi_predict(self.h, _ptr(m), _ptr(c), _ptr(s), m.shape[0],
                                                     int(bool(noise_f32)), BT_HOST))
        return m, c

    def kalman_update(self, mean, cov, meas, noise_f32=None):
        m = np.array(mean, dtype=np.float64, order="C").reshape(-1, 8)
        c = np.array(cov, dtype=np.float64, order="C").reshape(-1, 8, 8)
        z = _arr(meas, np.float64, (-1, 4))
        f = None if noise_f32 is None else _arr(noise_f32, np.uint8, (-1,))
        self._check(self.lib.bt_kalman_update(self.h, _ptr(m), _ptr(c), _ptr(z), None, None, _ptr(f), m.shape[0],
                                              BT_HOST))
        return m, c

    def kalman_project(self, mean, cov):
        m = _arr(mean, np.float64, (-1, 8))
        c = _arr(cov, np.float64, (-1, 8, 8))
        pm = np.empty((m.shape[0], 4), np.float64)
        pc = np.empty((m.shape[0], 4, 4), np.float64)
        self._check(self.lib.bt_kalman_project(self.h, _ptr(m), _ptr(c), _ptr(pm), _ptr(pc), m.shape[0], BT_HOST))
        return pm, pc

    # ---- matching ----
    def iou_distance(self, a_tlbr, b_tlbr):
        a = _arr(a_tlbr, np.float64, (-1, 4))
        b = _arr(b_tlbr, np.float64, (-1, 4))
        out = np.empty((a.shape[0], b.shape[0]), np.float64)
        self._check(self.lib.bt_iou_distance(self.h, _ptr(a), a.shape[0], _ptr(b), b.shape[0], _ptr(out), BT_HOST))
        return out

    def embedding_distance(self, a, b, precision: int = 0):
        a = _arr(a, np.float32)
        b = _arr(b, np.float32)
        d = a.shape[1] if a.ndim == 2 and a.shape[0] else (b.shape[1] if b.ndim == 2 else self.feat_dim)
        a = a.reshape(-1, d)
        b = b.reshape(-1, d)
        out = np.empty((a.shape[0], b.shape[0]), np.float32)
        self._check(self.lib.bt_embedding_distance(self.h, _ptr(a), a.shape[0], _ptr(b), b.shape[0], d, _ptr(out),
                                                   precision, BT_HOST))
        return out

    def fused_cost(self, trk_tlbr, det_tlbr, trk_feat, det_feat, stage: int = 1, face_sim=None, precision: int = 0):
        rt = _arr(trk_tlbr, np.float64, (-1, 4))
        ct = _arr(det_tlbr, np.float64, (-1, 4))
        a = _arr(trk_feat, np.float32)
        b = _arr(det_feat, np.float32)
        d = a.shape[-1]
        a = a.reshape(-1, d)
        b = b.reshape(-1, d)
        n, m = rt.shape[0], ct.shape[0]
        fs = None if face_sim is None else _arr(face_sim, np.float32, (n, m))
        out = np.empty((n, m), np.float64)
        self._check(self.lib.bt_fused_cost(self.h, _ptr(rt), n, _ptr(ct), m, _ptr(a), _ptr(b), d, _ptr(fs), stage,
                                           _ptr(out), precision, BT_HOST))
        return out

    def fuse_score(self, iou_dists, det_scores):
        d = _arr(iou_dists, np.float64)
        n, m = d.shape
        s = _arr(det_scores, np.float64, (m,))
        out = np.empty((n, m), np.float64)
        self._check(self.lib.bt_fuse_score(self.h, _ptr(d), _ptr(s), n, m, _ptr(out), BT_HOST))
        return out

    def lapjv(self, cost, thresh: float):
        c = _arr(cost, np.float64)
        n, m = c.shape
        x = np.empty(n, np.int32)
        y = np.empty(m, np.int32)
        self._check(self.lib.bt_linear_assignment(self.h, _ptr(c), n, m, float(thresh), _ptr(x), _ptr(y), BT_HOST))
        return x, y

    def feature_ema(self, smooth, curr, feat, first=None, alpha: float = 0.9):
        s = np.array(smooth, dtype=np.float32, order="C")
        c = np.array(curr, dtype=np.float32, order="C")
        f = _arr(feat, np.float32)
        k, d = f.shape
        fl = None if first is None else _arr(first, np.uint8, (k,))
        self._check(self.lib.bt_feature_ema(self.h, _ptr(s), _ptr(c), _ptr(f), None, None, _ptr(fl), k, d,
                                            float(alpha), BT_HOST))
        return s, c

    # ---- detector side ----
    def yolox_postprocess(self, raw_head, cfg: Optional[BtYoloxConfig] = None, max_out: int = 256):
        if cfg is None:
            cfg = BtYoloxConfig()
            self.lib.bt_default_yolox_config(C.byref(cfg))
        raw = _arr(raw_head, np.float32)
        out = np.zeros((max_out, 6), np.float64)
        cnt = np.zeros(1, np.int32)
        self._check(self.lib.bt_yolox_postprocess(self.h, _ptr(raw), C.byref(cfg), _ptr(out), max_out, _ptr(cnt),
                                                  BT_HOST))
        return out[: int(cnt[0])]

    def reid_crop_gather(self, frame, boxes, out_h: int = 256, out_w: int = 128):
        f = _arr(frame, np.uint8)
        h, w = f.shape[:2]
        b = _arr(boxes, np.int32, (-1, 4))
        out = np.empty((b.shape[0], 3, out_h, out_w), np.float32)
        self._check(self.lib.bt_reid_crop_gather(self.h, _ptr(f), h, w, _ptr(b), b.shape[0], out_h, out_w, _ptr(out),
                                                 BT_HOST))
        return out

    def detect_stage(self, raw_head_ptr: int, frame_ptr: int, h: int, w: int, crops_ptr: int,
                     cfg: Optional[BtYoloxConfig] = None, out_h: int = 256, out_w: int = 128, stream: int = 0,
                     det_out_ptr: int = 0, max_out: int = 0, det_count_ptr: int = 0):
        """Device-chained detector side (all device pointers, asynchronous): raw YOLOX head -> body detections in the
        stream's next input buffers + their ReID crops in `crops_ptr`."""
        if cfg is None:
            cfg = BtYoloxConfig()
            self.lib.bt_default_yolox_config(C.byref(cfg))
        self._check(self.lib.bt_detect_stage(self.h, stream, C.c_void_p(raw_head_ptr), C.byref(cfg), C.c_void_p(frame_ptr),
                                             h, w, out_h, out_w, C.c_void_p(crops_ptr),
                                             C.c_void_p(det_out_ptr) if det_out_ptr else None, max_out,
                                             C.c_void_p(det_count_ptr) if det_count_ptr else None))
        return cfg

    # ---- tracker ----
    def default_config(self) -> BtConfig:
        cfg = BtConfig()
        self.lib.bt_default_config(C.byref(cfg))
        return cfg

    def tracker_reset(self, cfg: Optional[BtConfig] = None, stream: int = -1):
        """Reset one video stream's tracker (stream >= 0) or all of them (default)."""
        self._check(self.lib.bt_tracker_reset_stream(self.h, stream, None if cfg is None else C.byref(cfg)))

    @staticmethod
    def _feat_array(feats, m, d):
        """float16 arrays are passed through as they are (BT_F16), everything else as float32 (BT_F32)."""
        if feats is None:
            return None, BT_F32
        a = np.asarray(feats)
        if a.dtype == np.float16:
            return np.ascontiguousarray(a).reshape(m, d), BT_F16
        return _arr(a, np.float32, (m, d)), BT_F32

    def update_arrays(self, boxes, scores, feats=None, face_sim=None, stream: int = 0) -> dict:
        """One BoTSORT.update of video stream `stream` on detector / encoder outputs (NumPy arrays)."""
        return self.update_streams([stream], [boxes], [scores], [feats],
                                   None if face_sim is None else [face_sim])[0]

    def _pack_batch(self, stream_ids, boxes, scores, feats, face_sims):
        n = len(stream_ids)
        ids = (C.c_int32 * n)(*[int(s) for s in stream_ids])
        keep, bp, sp, fp, mp, xp = [], (C.c_void_p * n)(), (C.c_void_p * n)(), (C.c_void_p * n)(), (C.c_int32 * n)(), (C.c_void_p * n)()
        dtype = None
        for k in range(n):
            b = _arr(boxes[k], np.int32, (-1, 4))
            s = _arr(scores[k], np.float32, (-1,))
            m = b.shape[0]
            f, dt = self._feat_array(None if feats is None else feats[k], m, self.feat_dim)
            if f is not None:
                if dtype is not None and dt != dtype:
                    raise ValueError("one batch cannot mix float16 and float32 features")
                dtype = dt
            x = None if (face_sims is None or face_sims[k] is None) else _arr(face_sims[k], np.float32)
            keep += [b, s, f, x]
            bp[k], sp[k], mp[k] = b.ctypes.data, s.ctypes.data, m
            fp[k] = None if f is None else f.ctypes.data
            xp[k] = None if x is None else x.ctypes.data
        return n, ids, bp, sp, fp, mp, (BT_F32 if dtype is None else dtype), (None if face_sims is None else xp), keep

    def update_streams(self, stream_ids, boxes, scores, feats=None, face_sims=None) -> list:
        """One frame step for several video streams of this ctx at once (one launch per kernel)."""
        n, ids, bp, sp, fp, mp, dtype, xp, keep = self._pack_batch(stream_ids, boxes, scores, feats, face_sims)
        infos = (BtFrameInfo * n)()
        self._check(self.lib.bt_update_streams(self.h, n, ids, bp, sp, fp, mp, dtype, xp, BT_HOST, infos))
        return [infos[k].as_dict() for k in range(n)]

    def submit_streams(self, stream_ids, boxes, scores, feats=None, face_sims=None):
        """Start moving one frame of inputs in (returns at once); the arrays must stay alive until the
        matching step_streams returns -- the returned handle keeps them referenced."""
        n, ids, bp, sp, fp, mp, dtype, xp, keep = self._pack_batch(stream_ids, boxes, scores, feats, face_sims)
        self._check(self.lib.bt_submit_streams(self.h, n, ids, bp, sp, fp, mp, dtype, xp, BT_HOST))
        return keep

    def step_streams(self, stream_ids) -> list:
        n = len(stream_ids)
        ids = (C.c_int32 * n)(*[int(s) for s in stream_ids])
        infos = (BtFrameInfo * n)()
        self._check(self.lib.bt_step_streams(self.h, n, ids, infos))
        return [infos[k].as_dict() for k in range(n)]

    def input_buffers(self, stream: int = 0):
        """Device pointers (boxes int32, scores float32, feats float16) the next submit of `stream` reads in place."""
        b, s, f = C.c_void_p(), C.c_void_p(), C.c_void_p()
        self._check(self.lib.bt_input_buffers(self.h, stream, C.byref(b), C.byref(s), C.byref(f)))
        return int(b.value), int(s.value), int(f.value)

    def update_arrays_raw(self, boxes_ptr: int, scores_ptr: int, feats_ptr: int, m: int, loc: int,
                          info: Optional[BtFrameInfo] = None, dtype: int = BT_F32, stream: int = 0):
        """Pointer-level call (pinned host or device buffers) for one stream, used by bench.py."""
        self.update_streams_raw([stream], [boxes_ptr], [scores_ptr], [feats_ptr], [m], loc, dtype,
                                None if info is None else C.pointer(info))

    def _raw_arrays(self, stream_ids, boxes_ptrs, scores_ptrs, feats_ptrs, ms):
        # the argument vectors of a steady pipeline repeat (double-buffered addresses, fixed m): keep the last few
        key = (tuple(stream_ids), tuple(boxes_ptrs), tuple(scores_ptrs), tuple(feats_ptrs), tuple(ms))
        cache = self.__dict__.setdefault("_raw_cache", {})
        hit = cache.get(key)
        if hit is not None:
            return hit
        n = len(stream_ids)
        ids = (C.c_int32 * n)(*stream_ids)
        bp = (C.c_void_p * n)(*boxes_ptrs)
        sp = (C.c_void_p * n)(*scores_ptrs)
        fp = (C.c_void_p * n)(*[p or None for p in feats_ptrs])
        mp = (C.c_int32 * n)(*ms)
        if len(cache) >= 64:
            cache.clear()
        cache[key] = out = (n, ids, bp, sp, fp, mp)
        return out

    def update_streams_raw(self, stream_ids, boxes_ptrs, scores_ptrs, feats_ptrs, ms, loc, dtype=BT_F32, infos=None):
        n, ids, bp, sp, fp, mp = self._raw_arrays(stream_ids, boxes_ptrs, scores_ptrs, feats_ptrs, ms)
        self._check(self.lib.bt_update_streams(self.h, n, ids, bp, sp, fp, mp, dtype, None, loc, infos))

    def submit_streams_raw(self, stream_ids, boxes_ptrs, scores_ptrs, feats_ptrs, ms, loc, dtype=BT_F32):
        n, ids, bp, sp, fp, mp = self._raw_arrays(stream_ids, boxes_ptrs, scores_ptrs, feats_ptrs, ms)
        self._check(self.lib.bt_submit_streams(self.h, n, ids, bp, sp, fp, mp, dtype, None, loc))

    def step_streams_raw(self, stream_ids, infos=None):
        n = len(stream_ids)
        ids = (C.c_int32 * n)(*stream_ids)
        self._check(self.lib.bt_step_streams(self.h, n, ids, infos))

    def get_tracks(self, which: int = 0, with_state: bool = False, stream: int = 0) -> dict:
        n = C.c_int32(0)
        hints = self.__dict__.setdefault("_list_size_hint", {})
        k = hints.get((stream, which), 0)   # steady state: the list is as long as last time -> one call
        for attempt in range(2):
            out = {name: np.empty(k, np.int32) for name in
                   ("ids", "state", "activated", "frame_id", "start_frame", "tracklet_len", "det_index")}
            out["score"] = np.empty(k, np.float32)
            out["tlbr"] = np.empty((k, 4), np.float64)
            mean = np.empty((k, 8), np.float64) if with_state else None
            cov = np.empty((k, 8, 8), np.float64) if with_state else None
            st = self.lib.bt_get_tracks_stream(
                self.h, stream, which, k, C.byref(n), _ptr(out["ids"]), _ptr(out["state"]), _ptr(out["activated"]),
                _ptr(out["frame_id"]), _ptr(out["start_frame"]), _ptr(out["tracklet_len"]), _ptr(out["det_index"]),
                _ptr(out["score"]), _ptr(out["tlbr"]), _ptr(mean), _ptr(cov))
            if attempt == 0 and n.value > k and st in (BT_OK, BT_ERR_CAPACITY):
                k = n.value                 # the list grew: the library reported its length, size the buffers again
                continue
            self._check(st)
            break
        hints[(stream, which)] = cnt = n.value
        if cnt < k:
            out = {name: a[:cnt] for name, a in out.items()}
            mean = None if mean is None else mean[:cnt]
            cov = None if cov is None else cov[:cnt]
        if with_state:
            out["mean"], out["cov"] = mean, cov
        return out

    def get_track_features(self, which: int = 0, stream: int = 0):
        n = C.c_int32(0)
        self._check(self.lib.bt_get_tracks_stream(self.h, stream, which, 0, C.byref(n), *([None] * 11)))
        k = n.value
        curr = np.empty((k, self.feat_dim), np.float32)
        smooth = np.empty((k, self.feat_dim), np.float32)
        self._check(self.lib.bt_get_track_features_stream(self.h, stream, which, k, _ptr(curr), _ptr(smooth)))
        return curr, smooth

    def get_matches(self, stage: int, stream: int = 0) -> np.ndarray:
        n = C.c_int32(0)
        self._check(self.lib.bt_get_matches_stream(self.h, stream, stage, 0, C.byref(n), None))
        pairs = np.empty((n.value, 2), np.int32)
        if n.value:
            self._check(self.lib.bt_get_matches_stream(self.h, stream, stage, n.value, C.byref(n), _ptr(pairs)))
        return pairs
