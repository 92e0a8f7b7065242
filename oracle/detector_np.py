"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the detector-side pieces of the hot path.

Only `tests/`, `__graft_entry__.smoke()` and bench.py's CPU legs may import this module.

  yolox_decode_nms / yolox_postprocess
      The reference keeps YOLOX decode + NMS INSIDE its ONNX graph
      (/root/reference/README.md:179-183, README.md:197-244; model file name
      `..._post_..._score015_iou080_box050.onnx`, demo:34) and the graph is not in the reference
      tree, so this restates the published definitions: standard YOLOX head decode
      ((xy + grid) * stride, exp(wh) * stride, score = sigmoid(obj) * sigmoid(cls)) and ONNX
      NonMaxSuppression-11 (per class, score > score_threshold, descending score -- equal scores:
      lower anchor index first --, suppress when IoU > iou_threshold, at most
      max_output_boxes_per_class; output ordered by class then score).  PARITY UNPINNED by the
      reference for this part.  The Python tail follows YOLOX._postprocess, demo:996-1030.
  crop_preprocess
      demo:1434-1436 (crop) + FastReID._preprocess demo:1101-1142.  `resize_linear_u8` restates
      cv2.resize(INTER_LINEAR, uint8) and is pinned bit-exactly against cv2 itself in
      tests/test_oracle_detector.py.
(demo = /root/reference/demo_bottrack_onnx_tflite.py)
"""
from __future__ import annotations

import numpy as np


# --------------------------------------------------------------------------------------------
# YOLOX
# --------------------------------------------------------------------------------------------
def yolox_grid(in_h: int, in_w: int):
    gx, gy, st = [], [], []
    for s in (8, 16, 32):
        h, w = in_h // s, in_w // s
        yy, xx = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
        gx.append(xx.reshape(-1))
        gy.append(yy.reshape(-1))
        st.append(np.full(h * w, s))
    return (np.concatenate(gx).astype(np.float32), np.concatenate(gy).astype(np.float32),
            np.concatenate(st).astype(np.float32))


def _sigmoid32(x):
    x = np.asarray(x, dtype=np.float32)
    return (np.float32(1.0) / (np.float32(1.0) + np.exp(-x))).astype(np.float32)


def _iou32(a, b):
    a = a.astype(np.float32)
    b = b.astype(np.float32)
    area1 = (a[2] - a[0]) * (a[3] - a[1])
    area2 = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    ix1 = np.maximum(a[0], b[:, 0]); iy1 = np.maximum(a[1], b[:, 1])
    ix2 = np.minimum(a[2], b[:, 2]); iy2 = np.minimum(a[3], b[:, 3])
    iw = np.maximum(ix2 - ix1, np.float32(0)); ih = np.maximum(iy2 - iy1, np.float32(0))
    inter = (iw * ih).astype(np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        iou = inter / (area1 + area2 - inter)
    iou = np.where((area1 <= 0) | (area2 <= 0), np.float32(0), iou)
    return iou.astype(np.float32)


def yolox_decode_nms(raw: np.ndarray, in_h: int = 480, in_w: int = 640, score_thresh: float = 0.15,
                     iou_thresh: float = 0.80, max_per_class: int = 50):
    """raw [anchors, 5+C] float32 -> rows [class, score, x1, y1, x2, y2] (float32, model pixels)."""
    raw = np.asarray(raw, dtype=np.float32)
    gx, gy, st = yolox_grid(in_h, in_w)
    cx = (raw[:, 0] + gx) * st
    cy = (raw[:, 1] + gy) * st
    bw = np.exp(raw[:, 2]) * st
    bh = np.exp(raw[:, 3]) * st
    half = np.float32(0.5)
    boxes = np.stack([cx - bw * half, cy - bh * half, cx + bw * half, cy + bh * half], axis=1).astype(np.float32)
    obj = _sigmoid32(raw[:, 4])
    cls = _sigmoid32(raw[:, 5:])
    scores = (obj[:, None] * cls).astype(np.float32)
    out = []
    for c in range(scores.shape[1]):
        s = scores[:, c]
        cand = np.nonzero(s > np.float32(score_thresh))[0]
        order = cand[np.lexsort((cand, -s[cand].astype(np.float64)))]     # score desc, index asc
        keep = []
        suppressed = np.zeros(len(order), bool)
        for k, a in enumerate(order):
            if suppressed[k]:
                continue
            if len(keep) >= max_per_class:
                break
            keep.append(a)
            rest = np.arange(k + 1, len(order))
            if len(rest):
                iou = _iou32(boxes[a], boxes[order[rest]])
                suppressed[rest] |= iou > np.float32(iou_thresh)
        for a in keep:
            out.append([c, s[a], *boxes[a]])
    return np.asarray(out, dtype=np.float32).reshape(-1, 6)


def yolox_postprocess(raw, img_h: int, img_w: int, in_h: int = 480, in_w: int = 640, post_score: float = 0.35,
                      **nms_kw):
    """Decode + NMS, then YOLOX._postprocess (demo:1001-1027): score > 0.35, float32 multiply then
    divide, int() truncation.  Rows [class, score, x1, y1, x2, y2] (float64 container)."""
    det = yolox_decode_nms(raw, in_h, in_w, **nms_kw)
    rows = []
    for box in det:
        score = box[1]
        if not score > post_score:
            continue
        x_min = int(max(0, box[2]) * img_w / in_w)
        y_min = int(max(0, box[3]) * img_h / in_h)
        x_max = int(min(box[4], in_w) * img_w / in_w)
        y_max = int(min(box[5], in_h) * img_h / in_h)
        rows.append([int(box[0]), float(score), x_min, y_min, x_max, y_max])
    return np.asarray(rows, dtype=np.float64).reshape(-1, 6)


def synth_yolox_head(rng, boxes_xyxy, classes, scores, in_h=480, in_w=640, num_classes=4, clutter=300):
    """Raw head [anchors, 5+C] whose decode yields the planted boxes (inverse of the decode) plus
    low-score clutter and near-duplicate anchors that NMS has to suppress."""
    gx, gy, st = yolox_grid(in_h, in_w)
    n = len(gx)
    raw = np.zeros((n, 5 + num_classes), np.float32)
    raw[:, 0:2] = rng.uniform(0, 1, (n, 2))
    raw[:, 2:4] = rng.normal(0.5, 0.5, (n, 2))
    raw[:, 4] = rng.normal(-6, 1, n)            # objectness logit: background
    raw[:, 5:] = rng.normal(-3, 1, (n, num_classes))
    idx = rng.choice(n, size=min(clutter, n), replace=False)
    raw[idx, 4] = rng.normal(-0.5, 0.7, len(idx))     # clutter around the 0.15/0.35 thresholds
    raw[idx, 5:] = rng.normal(0.0, 1.0, (len(idx), num_classes))

    def logit(p):
        return np.log(p / (1 - p))

    for (x1, y1, x2, y2), c, s in zip(boxes_xyxy, classes, scores):
        cx, cy, w, h = (x1 + x2) / 2, (y1 + y2) / 2, x2 - x1, y2 - y1
        stride = 8 if max(w, h) < 64 else (16 if max(w, h) < 160 else 32)
        sel = np.nonzero(st == stride)[0]
        ww = in_w // stride
        gxi = min(int(cx // stride), ww - 1)
        gyi = min(int(cy // stride), in_h // stride - 1)
        a = sel[gyi * ww + gxi]
        for dup, jitter in ((a, 0.0), (a + 1 if (a + 1) in sel else a - 1, 1.5)):   # near-duplicate neighbour
            raw[dup, 0] = (cx + jitter) / stride - gx[dup]
            raw[dup, 1] = (cy + jitter) / stride - gy[dup]
            raw[dup, 2] = np.log(w / stride)
            raw[dup, 3] = np.log(h / stride)
            p = np.sqrt(s) if jitter == 0.0 else np.sqrt(s * 0.8)
            raw[dup, 4] = logit(p)
            raw[dup, 5:] = -6.0
            raw[dup, 5 + c] = logit(p)
    return raw


# --------------------------------------------------------------------------------------------
# crop + resize + normalise
# --------------------------------------------------------------------------------------------
def _coords(dst, src, is_x):
    d = np.arange(dst, dtype=np.float64)
    inv = np.float64(dst) / np.float64(src)
    scale = 1.0 / inv
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    f = (f - s.astype(np.float32)).astype(np.float32)
    if is_x:
        lo = s < 0
        f[lo] = 0; s[lo] = 0
        hi = s >= src - 1
        f[hi] = 0; s[hi] = src - 1
        s0 = s
        s1 = np.minimum(s + 1, src - 1)
    else:
        s0 = np.clip(s, 0, src - 1)
        s1 = np.clip(s + 1, 0, src - 1)
    a0 = np.rint((np.float32(1.0) - f) * np.float32(2048)).astype(np.int64)
    a1 = np.rint(f * np.float32(2048)).astype(np.int64)
    return s0, s1, a0, a1


def resize_linear_u8(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """cv2.resize(img, (out_w, out_h)) for uint8 HxWxC, INTER_LINEAR, restated."""
    h, w = img.shape[:2]
    sx0, sx1, ax0, ax1 = _coords(out_w, w, True)
    sy0, sy1, by0, by1 = _coords(out_h, h, False)
    src = img.astype(np.int64)
    hor = src[:, sx0, :] * ax0[None, :, None] + src[:, sx1, :] * ax1[None, :, None]     # [h, out_w, C]
    r0 = hor[sy0] >> 4
    r1 = hor[sy1] >> 4
    out = (((by0[:, None, None] * r0) >> 16) + ((by1[:, None, None] * r1) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


MEAN = np.array([0.485, 0.456, 0.406], dtype=np.float32).reshape(1, 3, 1, 1)
STD = np.array([0.229, 0.224, 0.225], dtype=np.float32).reshape(1, 3, 1, 1)


def crop_preprocess(frame: np.ndarray, boxes: np.ndarray, out_h: int = 256, out_w: int = 128,
                    resize=resize_linear_u8) -> np.ndarray:
    """demo:1434-1436 + demo:1124-1142.  frame uint8 [H,W,3] BGR, boxes int [N,4] -> float32 [N,3,H,W]."""
    outs = []
    for x1, y1, x2, y2 in np.asarray(boxes, dtype=np.int64):
        crop = frame[y1:y2, x1:x2, :]
        if crop.shape[0] == 0 or crop.shape[1] == 0:
            outs.append(None)
            continue
        r = resize(crop, out_h, out_w)
        r = r[..., ::-1].transpose(2, 0, 1)
        outs.append(r)
    res = np.zeros((len(outs), 3, out_h, out_w), np.float32)
    for i, r in enumerate(outs):
        if r is None:
            continue
        res[i] = ((r[None] / 255.0 - MEAN) / STD).astype(np.float32)[0]
    return res
