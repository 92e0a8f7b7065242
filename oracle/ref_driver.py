"""TEST INFRASTRUCTURE ONLY -- drives the REAL reference tracker on array inputs.

Runs `BoTSORT.update` of `/root/reference/demo_bottrack_onnx_tflite.py` (imported in place by
oracle/ref_loader.py) with stub models:
  detector      : returns the given int boxes as class-0 `Box` objects (what YOLOX._postprocess
                  would emit, demo:1015-1027)
  body encoder  : returns (sims[N_det, M_trk] = f_det . f_trk^T, feats[N_det, D])  (demo:1452-1460)
  face encoder  : returns (feats, sims) swapped (demo:1478-1480) with sims = 0 and non-zero feats
Only usable in the build container (needs /root/reference).  One tracker at a time: the
reference's id counter is process-global (demo:390, demo:1264).
"""
from __future__ import annotations

import numpy as np

from .ref_loader import load_reference


class _StubDetector:
    def __init__(self, ref):
        self.ref = ref
        self.boxes = None
        self.scores = None

    def __call__(self, image):
        out = []
        for b, s in zip(self.boxes, self.scores):
            out.append(self.ref.Box(trackid=0, classid=0, score=float(s), x1=int(b[0]), y1=int(b[1]),
                                    x2=int(b[2]), y2=int(b[3]), cx=0, cy=0, is_used=False))
        return out


class _StubBody:
    def __init__(self, dim):
        self.feature_size = dim
        self.feats = None

    def __call__(self, *, base_images, target_features):
        f = np.array(self.feats, dtype=np.float32, copy=True)
        t = np.asarray(target_features, dtype=np.float32).reshape(-1, self.feature_size)
        sims = np.matmul(f, t.T).astype(np.float32) if len(t) else np.zeros((len(f), 0), np.float32)
        return sims, f


class _StubFace:
    feature_size = 4
    _input_shapes = [[1, 3, 8, 8]]

    def __call__(self, *, base_images, target_features):
        n = len(base_images)
        t = np.asarray(target_features, dtype=np.float32).reshape(-1, self.feature_size)
        feats = np.full((n, self.feature_size), 0.5, dtype=np.float32)
        sims = np.zeros((n, len(t)), dtype=np.float32)
        return feats, sims


class ReferenceRunner:
    def __init__(self, feat_dim: int):
        self.ref = load_reference()
        self.det = _StubDetector(self.ref)
        self.body = _StubBody(feat_dim)
        self.face = _StubFace()
        self.tracker = self.ref.BoTSORT(self.det, self.body, self.face, frame_rate=30)
        self.image = np.zeros((4, 4, 3), dtype=np.uint8)

    def update_arrays(self, boxes, scores, feats):
        self.det.boxes = boxes
        self.det.scores = scores
        self.body.feats = feats
        return self.tracker.update(self.image)

    def snapshot(self):
        def pack(lst):
            return {
                "ids": np.array([t.track_id for t in lst], dtype=np.int64),
                "state": np.array([t.state for t in lst], dtype=np.int64),
                "activated": np.array([t.is_activated for t in lst], dtype=bool),
                "tlbr": np.array([t.tlbr for t in lst], dtype=np.float64).reshape(-1, 4),
                "score": np.array([t.score for t in lst], dtype=np.float64),
                "frame_id": np.array([t.frame_id for t in lst], dtype=np.int64),
                "start_frame": np.array([t.start_frame for t in lst], dtype=np.int64),
                "tracklet_len": np.array([t.tracklet_len for t in lst], dtype=np.int64),
                "mean": np.array([np.asarray(t.mean, dtype=np.float64) for t in lst]).reshape(-1, 8),
                "cov": np.array([np.asarray(t.covariance, dtype=np.float64) for t in lst]).reshape(-1, 8, 8),
            }
        return {"tracked": pack(self.tracker.tracked_stracks), "lost": pack(self.tracker.lost_stracks)}
