"""Seeded synthetic tracking scenes.

The reference ships no data and no tests, so every workload here is synthetic: N identities
on a jittered grid, per-frame random-walk motion, integer-truncated boxes exactly like
`YOLOX._postprocess` produces them (`/root/reference/demo_bottrack_onnx_tflite.py:1009-1012`),
a configurable fraction of low-score detections (second association stage, `demo:1568-1586`)
and of dropped detections (lost / re-found tracks, `demo:1582-1586`, `demo:1564-1566`),
and unit-norm ReID features `f = normalise(g_id + sigma * noise)` (same-id cosine ~0.99,
cross-id ~0 +- 0.02 at D=2048).

Used by bench.py, __graft_entry__.smoke() and the tests; it is plain NumPy and holds no
tracker logic.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Iterator

import numpy as np


@dataclass
class SceneConfig:
    n_ids: int = 64                 # identities in the scene (tracks)
    feat_dim: int = 2048            # ReID feature size (Fast-ReID: 2048, demo:1060)
    seed: int = 0
    pitch_x: float = 70.0           # grid pitch; smaller => more IoU overlap between neighbours
    pitch_y: float = 120.0
    w_range: tuple = (40.0, 90.0)
    h_range: tuple = (80.0, 160.0)
    walk: float = 3.0               # per-frame centre random walk U[-walk, walk] px
    size_jitter: float = 1.5        # per-frame w/h jitter U[-j, j] px
    low_frac: float = 0.0           # fraction of detections with score in (0.1, 0.40]
    drop_frac: float = 0.0          # fraction of identities not detected in a frame
    mid_frac: float = 0.0           # fraction with score in (0.40, 0.9): matchable, never born
    feat_noise: float = 0.005       # sigma of the per-frame feature noise (per component)
    newcomer_every: int = 0         # every k frames one hidden identity starts to appear (0 = off)
    with_features: bool = True      # False => IoU-only (features all-zero rows)
    emit_features: bool = True      # False => next_frame() returns feats=None (the features of a frame are produced
                                    # elsewhere from `gt` and `identity`, e.g. on the GPU that owns the stream)


class SyntheticScene:
    """Deterministic stream of per-frame detections for one video stream."""

    def __init__(self, cfg: SceneConfig):
        self.cfg = cfg
        rng = np.random.default_rng(cfg.seed)
        n = cfg.n_ids
        cols = int(np.ceil(np.sqrt(n * cfg.pitch_y / cfg.pitch_x)))
        cols = max(cols, 1)
        gx = (np.arange(n) % cols).astype(np.float64)
        gy = (np.arange(n) // cols).astype(np.float64)
        self.cx = 100.0 + gx * cfg.pitch_x + rng.uniform(-10, 10, n)
        self.cy = 150.0 + gy * cfg.pitch_y + rng.uniform(-10, 10, n)
        self.w = rng.uniform(*cfg.w_range, n)
        self.h = rng.uniform(*cfg.h_range, n)
        self.canvas_w = int(200 + cols * cfg.pitch_x + 100)
        self.canvas_h = int(300 + (n // cols + 1) * cfg.pitch_y + 200)
        if cfg.with_features:
            g = rng.standard_normal((n, cfg.feat_dim)).astype(np.float32)
            g /= np.linalg.norm(g, axis=1, keepdims=True)
            self.identity = g
        else:
            self.identity = None
        # identities that are hidden at the start and appear later (late births -> unconfirmed path)
        self.appear_frame = np.zeros(n, dtype=np.int64)
        if cfg.newcomer_every > 0:
            k = max(1, n // 8)
            late = rng.choice(n, size=k, replace=False)
            self.appear_frame[late] = 2 + cfg.newcomer_every * np.arange(k)
        self.frame_no = 0
        self._rng_seed = cfg.seed

    def next_frame(self) -> Dict[str, np.ndarray]:
        """Advance one frame. Returns dict with
        boxes  int32  [M,4]  x1,y1,x2,y2 (image pixels, truncated like demo:1009-1012)
        scores float32[M]
        feats  float32[M,D]  unit-norm rows (zeros when with_features=False)
        gt     int64  [M]    identity index of each detection (for diagnostics only)
        """
        cfg = self.cfg
        self.frame_no += 1
        rng = np.random.default_rng([self._rng_seed, self.frame_no])
        n = cfg.n_ids
        self.cx += rng.uniform(-cfg.walk, cfg.walk, n)
        self.cy += rng.uniform(-cfg.walk, cfg.walk, n)
        w = self.w + rng.uniform(-cfg.size_jitter, cfg.size_jitter, n)
        h = self.h + rng.uniform(-cfg.size_jitter, cfg.size_jitter, n)
        visible = self.appear_frame <= self.frame_no
        if cfg.drop_frac > 0 and self.frame_no > 1:
            visible &= rng.uniform(size=n) >= cfg.drop_frac
        ids = np.nonzero(visible)[0]
        ids = rng.permutation(ids)
        m = len(ids)
        x1 = np.maximum(0.0, self.cx[ids] - w[ids] / 2)
        y1 = np.maximum(0.0, self.cy[ids] - h[ids] / 2)
        x2 = np.minimum(self.cx[ids] + w[ids] / 2, float(self.canvas_w))
        y2 = np.minimum(self.cy[ids] + h[ids] / 2, float(self.canvas_h))
        boxes = np.stack([x1, y1, x2, y2], axis=1).astype(np.int32)  # truncation toward zero
        scores = np.full(m, 0.95, dtype=np.float32)
        if self.frame_no > 1:
            r = rng.uniform(size=m)
            low = r < cfg.low_frac
            mid = (~low) & (r < cfg.low_frac + cfg.mid_frac)
            scores[low] = rng.uniform(0.15, 0.35, int(low.sum())).astype(np.float32)
            scores[mid] = rng.uniform(0.50, 0.85, int(mid.sum())).astype(np.float32)
        if not cfg.emit_features:
            feats = None
        elif cfg.with_features:
            f = self.identity[ids] + cfg.feat_noise * rng.standard_normal((m, cfg.feat_dim)).astype(np.float32)
            f /= np.linalg.norm(f, axis=1, keepdims=True)
            feats = np.ascontiguousarray(f, dtype=np.float32)
        else:
            feats = np.zeros((m, cfg.feat_dim), dtype=np.float32)
        return {"boxes": boxes, "scores": scores, "feats": feats, "gt": ids.astype(np.int64)}

    def frames(self, count: int) -> Iterator[Dict[str, np.ndarray]]:
        for _ in range(count):
            yield self.next_frame()


def steady_state_pair(n: int, m: int, feat_dim: int, seed: int = 0, with_features: bool = True,
                      pitch_x: float = 70.0, pitch_y: float = 120.0):
    """Two consecutive frames of a scene with n identities (m = n detections are produced;
    if m < n the detection list is truncated).  Convenience for kernel-level tests/benches."""
    sc = SyntheticScene(SceneConfig(n_ids=n, feat_dim=feat_dim, seed=seed, with_features=with_features,
                                    pitch_x=pitch_x, pitch_y=pitch_y))
    f1 = sc.next_frame()
    f2 = sc.next_frame()
    if m < n:
        for k in ("boxes", "scores", "feats", "gt"):
            f2[k] = f2[k][:m]
    return f1, f2


# ------------------------------------------------------------------------------------------------
# detector-side synthetic inputs (BASELINE config 4): a camera frame and the raw YOLOX head for it
# ------------------------------------------------------------------------------------------------
def yolox_anchor_grid(in_h: int, in_w: int):
    """(grid_x, grid_y, stride) of every anchor, strides 8 / 16 / 32 in that order (YOLOX head layout)."""
    gx, gy, st = [], [], []
    for s in (8, 16, 32):
        hh, ww = in_h // s, in_w // s
        yy, xx = np.meshgrid(np.arange(hh), np.arange(ww), indexing="ij")
        gx.append(xx.reshape(-1)); gy.append(yy.reshape(-1)); st.append(np.full(hh * ww, s))
    return np.concatenate(gx), np.concatenate(gy), np.concatenate(st)


class DetectorScene:
    """`k` pedestrians random-walking over a (h, w) camera frame.  next_frame() returns the BGR uint8 frame (every
    identity carries its own fixed texture, so a crop-based ReID stub sees the same person in every frame) and the
    raw head [anchors, 5 + C] float32: background logits, threshold-level clutter, and for every person a cluster of
    `dups` neighbouring anchors that all decode to (almost) the same box -- the work NMS exists for."""

    def __init__(self, k: int = 40, h: int = 480, w: int = 640, num_classes: int = 4, dups: int = 6, clutter: int = 300,
                 seed: int = 0):
        self.rng = np.random.default_rng(seed)
        self.k, self.h, self.w, self.C, self.dups, self.clutter = k, h, w, num_classes, dups, clutter
        r = self.rng
        cols = int(np.ceil(np.sqrt(k * w / h)))
        rows = (k + cols - 1) // cols
        self.bw = r.uniform(28, 0.8 * w / cols, k)
        self.bh = r.uniform(60, 0.9 * h / rows, k)
        self.cx = ((np.arange(k) % cols) + 0.5) * w / cols + r.uniform(-4, 4, k)
        self.cy = ((np.arange(k) // cols) + 0.5) * h / rows + r.uniform(-4, 4, k)
        self.tex = []
        for i in range(k):                                   # blocky low-frequency texture: survives the encoder stub's pooling
            tr = np.random.default_rng(seed * 1000 + 17 + i)
            th, tw = int(self.bh[i]) + 8, int(self.bw[i]) + 8
            coarse = tr.integers(0, 256, (16, 8, 3)).astype(np.float64)
            yy = np.minimum(np.arange(th) * 16 // th, 15); xx = np.minimum(np.arange(tw) * 8 // tw, 7)
            t = coarse[yy][:, xx] + tr.normal(0, 6, (th, tw, 3))
            self.tex.append(np.clip(t, 0, 255).astype(np.uint8))
        self.gx, self.gy, self.st = yolox_anchor_grid(h, w)
        self.background = r.integers(0, 256, (h, w, 3), dtype=np.uint8)

    def boxes(self):
        x1 = np.clip(self.cx - self.bw / 2, 0, self.w - 2); y1 = np.clip(self.cy - self.bh / 2, 0, self.h - 2)
        x2 = np.clip(self.cx + self.bw / 2, x1 + 1, self.w - 1); y2 = np.clip(self.cy + self.bh / 2, y1 + 1, self.h - 1)
        return np.stack([x1, y1, x2, y2], axis=1)

    def next_frame(self):
        r = self.rng
        self.cx += r.uniform(-2, 2, self.k); self.cy += r.uniform(-2, 2, self.k)
        b = self.boxes()
        frame = self.background.copy()
        for i, (x1, y1, x2, y2) in enumerate(b.astype(int)):
            frame[y1:y2, x1:x2] = self.tex[i][: y2 - y1, : x2 - x1]
        n, C = len(self.gx), self.C
        raw = np.empty((n, 5 + C), np.float32)
        raw[:, 0:2] = r.uniform(0, 1, (n, 2))
        raw[:, 2:4] = r.normal(0.5, 0.5, (n, 2))
        raw[:, 4] = r.normal(-6, 1, n)
        raw[:, 5:] = r.normal(-3, 1, (n, C))
        idx = r.choice(n, size=min(self.clutter, n), replace=False)
        raw[idx, 4] = r.normal(-0.5, 0.7, len(idx))
        raw[idx, 5:] = r.normal(0.0, 1.0, (len(idx), C))
        logit = lambda p: np.log(p / (1 - p))
        for i, (x1, y1, x2, y2) in enumerate(b):
            cx, cy, w_, h_ = (x1 + x2) / 2, (y1 + y2) / 2, x2 - x1, y2 - y1
            s = 8 if max(w_, h_) < 64 else (16 if max(w_, h_) < 160 else 32)
            base = 0 if s == 8 else ((self.h // 8) * (self.w // 8) if s == 16 else (self.h // 8) * (self.w // 8) + (self.h // 16) * (self.w // 16))
            ww, hh = self.w // s, self.h // s
            gxi, gyi = min(int(cx // s), ww - 1), min(int(cy // s), hh - 1)
            for d in range(self.dups):                       # the cell itself, then its neighbours
                ox, oy = ((0, 0), (1, 0), (-1, 0), (0, 1), (0, -1), (1, 1), (-1, -1), (1, -1), (-1, 1))[d % 9]
                ax, ay = min(max(gxi + ox, 0), ww - 1), min(max(gyi + oy, 0), hh - 1)
                if d and (ax, ay) == (gxi, gyi):
                    continue
                a = base + ay * ww + ax
                jit = 0.0 if d == 0 else r.uniform(-1.5, 1.5)
                p = np.sqrt(0.96 * (1.0 if d == 0 else r.uniform(0.5, 0.9)))
                raw[a, 0] = (cx + jit) / s - self.gx[a]; raw[a, 1] = (cy + jit) / s - self.gy[a]
                raw[a, 2] = np.log(w_ / s); raw[a, 3] = np.log(h_ / s)
                raw[a, 4] = logit(p); raw[a, 5:] = -6.0; raw[a, 5] = logit(p)
        return frame, raw
