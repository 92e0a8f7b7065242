"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the BoT-SORT hot path.

Only `tests/`, `__graft_entry__.smoke()` and bench.py's CPU legs (`cpu_baseline`,
`--impl reference`) may import this module, and only as the checker / the timed CPU
baseline.  Nothing in the product package (`bot-sort-onnx-tensorrt_b200/`) imports it.

It is a NumPy restatement of the reference's per-frame tracker arithmetic, each function
citing the lines of `/root/reference/demo_bottrack_onnx_tflite.py` ("demo:NNN") it follows.
Parity pinning: the reference has no tests / golden vectors, so this
restatement is pinned against the reference ITSELF executed in the build container
(`oracle/ref_loader.py` + `tests/test_oracle_vs_reference.py`) and against fixtures generated
from it (`oracle/gen_golden.py` -> `tests/golden/*.npz`).  Pieces whose arithmetic is not in
the reference tree (lap's solver, the in-graph ReID cosine) are "parity unpinned by the
reference"; see oracle/lapjv_port.c and `embedding_distance` below.

Two cost structures are offered because the reference's CPU time is dominated by a pure
Python IoU loop (demo:1742):
  mode="faithful"    -- same loop structure as the reference (Python double loop IoU,
                        one Kalman update call per match); used on small cases and as the
                        "as-is" CPU baseline sample.
  mode="vectorized"  -- broadcast IoU, batched Kalman update; the strongest CPU baseline.
Both produce identical decisions; floats agree to ~1e-12.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import List, Optional, Sequence, Tuple

import numpy as np
import scipy.linalg

_HERE = os.path.dirname(os.path.abspath(__file__))

# --------------------------------------------------------------------------------------
# constants (demo:163-164, demo:1268-1277, demo:1571, demo:1604, demo:1667, demo:473)
# --------------------------------------------------------------------------------------
STD_POS = 1.0 / 20
STD_VEL = 1.0 / 160
TRACK_HIGH_THRESH = 0.40
TRACK_LOW_THRESH = 0.1
NEW_TRACK_THRESH = 0.9
MATCH_THRESH = 0.8
SECOND_THRESH = 0.5
UNCONF_THRESH = 0.7
PROXIMITY_THRESH = 0.5
APPEARANCE_THRESH = 0.25
DUP_IOU_DIST = 0.15
TRACK_BUFFER = 300
ALPHA = 0.9

ST_NEW, ST_TRACKED, ST_LOST, ST_LONGLOST, ST_REMOVED = 0, 1, 2, 3, 4

_MOTION = np.eye(8, 8)
for _i in range(4):
    _MOTION[_i, 4 + _i] = 1.0
_UPDATE = np.eye(4, 8)


# --------------------------------------------------------------------------------------
# Kalman filter (demo:118-336)
# --------------------------------------------------------------------------------------
def kf_initiate(measurement: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """demo:166-197.  dtype follows NumPy promotion of the input (float32 in the tracker,
    demo:465/561): under numpy>=2 both outputs are float32, under the pinned numpy 1.24 the
    covariance is float64."""
    mean_pos = measurement
    mean_vel = np.zeros_like(mean_pos)
    mean = np.r_[mean_pos, mean_vel]
    std = [
        2 * STD_POS * measurement[2], 2 * STD_POS * measurement[3],
        2 * STD_POS * measurement[2], 2 * STD_POS * measurement[3],
        10 * STD_VEL * measurement[2], 10 * STD_VEL * measurement[3],
        10 * STD_VEL * measurement[2], 10 * STD_VEL * measurement[3]]
    covariance = np.diag(np.square(std))
    return mean, covariance


def kf_multi_predict(mean: np.ndarray, covariance: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """demo:265-302, closed form: with P = [[A,B],[C,D]] (4x4 blocks),
    F P F^T = [[(A+C)+(B+D), B+D],[C+D, D]]; the association order equals the two np.dot
    calls at demo:299-300 (all other terms of the dgemm sums are exact zeros), so this is
    bit-identical to the reference for finite inputs (checked in tests).
    dtype quirk kept: the motion noise is evaluated in the dtype of `mean` (demo:281-291);
    on frame 2 every pool track still carries the float32 mean of `initiate`, so the noise is
    float32 arithmetic there, float64 afterwards."""
    mean = np.asarray(mean)
    if mean.dtype != np.float32:
        mean = mean.astype(np.float64)
    covariance = np.asarray(covariance, dtype=np.float64)
    w = mean[:, 2]
    h = mean[:, 3]
    sp = np.stack([STD_POS * w, STD_POS * h, STD_POS * w, STD_POS * h,
                   STD_VEL * w, STD_VEL * h, STD_VEL * w, STD_VEL * h], axis=1)
    q = np.square(sp).astype(np.float64)
    mean = mean.astype(np.float64)
    new_mean = mean.copy()
    new_mean[:, :4] = mean[:, :4] + mean[:, 4:]
    left = covariance.copy()
    left[:, :4, :] = covariance[:, :4, :] + covariance[:, 4:, :]
    out = left.copy()
    out[:, :, :4] = left[:, :, :4] + left[:, :, 4:]
    idx = np.arange(8)
    out[:, idx, idx] += q
    return new_mean, out


def kf_project(mean: np.ndarray, covariance: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """demo:236-263."""
    std = [STD_POS * mean[2], STD_POS * mean[3], STD_POS * mean[2], STD_POS * mean[3]]
    innovation_cov = np.diag(np.square(std))
    pm = np.dot(_UPDATE, mean)
    pc = np.linalg.multi_dot((_UPDATE, covariance, _UPDATE.T))
    return pm, pc + innovation_cov


def kf_update(mean: np.ndarray, covariance: np.ndarray, measurement: np.ndarray):
    """demo:304-336 (one track)."""
    projected_mean, projected_cov = kf_project(mean, covariance)
    chol_factor, lower = scipy.linalg.cho_factor(projected_cov, lower=True, check_finite=False)
    kalman_gain = scipy.linalg.cho_solve((chol_factor, lower), np.dot(covariance, _UPDATE.T).T,
                                         check_finite=False).T
    innovation = measurement - projected_mean
    new_mean = mean + np.dot(innovation, kalman_gain.T)
    new_covariance = covariance - np.linalg.multi_dot((kalman_gain, projected_cov, kalman_gain.T))
    return new_mean, new_covariance


def kf_update_batch(mean: np.ndarray, covariance: np.ndarray, measurement: np.ndarray,
                    f32_state: Optional[np.ndarray] = None):
    """Vectorised demo:304-336 over K tracks: mean[K,8], covariance[K,8,8], measurement[K,4].
    `f32_state[k]` marks tracks whose mean is still the float32 output of `initiate` (never
    predicted): for those the reference evaluates the projection noise in float32
    (demo:253-258 on a float32 mean under numpy>=2)."""
    mean = np.asarray(mean)
    k = mean.shape[0]
    m64 = mean.astype(np.float64)
    p = np.asarray(covariance, dtype=np.float64)
    w = m64[:, 2]
    h = m64[:, 3]
    noise = np.square(np.stack([STD_POS * w, STD_POS * h, STD_POS * w, STD_POS * h], axis=1))
    if f32_state is not None and np.any(f32_state):
        m32 = mean.astype(np.float32)
        w32, h32 = m32[:, 2], m32[:, 3]
        sp32 = np.float32(STD_POS)
        n32 = np.square(np.stack([sp32 * w32, sp32 * h32, sp32 * w32, sp32 * h32], axis=1))
        noise = np.where(np.asarray(f32_state, bool)[:, None], n32.astype(np.float64), noise)
    s = p[:, :4, :4].copy()
    idx = np.arange(4)
    s[:, idx, idx] += noise
    pht = p[:, :, :4]                                   # P H^T  [K,8,4]
    gain = np.linalg.solve(s, pht.transpose(0, 2, 1)).transpose(0, 2, 1)   # [K,8,4]
    innovation = np.asarray(measurement, dtype=np.float64) - m64[:, :4]
    new_mean = m64 + np.einsum("ka,kra->kr", innovation, gain)
    new_cov = p - gain @ s @ gain.transpose(0, 2, 1)
    return new_mean, new_cov


# --------------------------------------------------------------------------------------
# IoU (demo:1695-1761)
# --------------------------------------------------------------------------------------
def bbox_iou(a, b) -> float:
    """demo:1695-1713 (scalar).  Strict `<=` empty rule, no +1 pixel convention."""
    ixmin = max(a[0], b[0]); iymin = max(a[1], b[1])
    ixmax = min(a[2], b[2]); iymax = min(a[3], b[3])
    if ixmax <= ixmin or iymax <= iymin:
        return 0.0
    inter = (ixmax - ixmin) * (iymax - iymin)
    area1 = (a[2] - a[0]) * (a[3] - a[1])
    area2 = (b[2] - b[0]) * (b[3] - b[1])
    return inter / float(area1 + area2 - inter)


def bbox_ious_loop(atlbrs, btlbrs) -> np.ndarray:
    """demo:1731-1743, the reference's own structure (pure Python double loop)."""
    ious = np.zeros((len(atlbrs), len(btlbrs)), dtype=np.float32)
    if ious.size == 0:
        return ious
    return np.array([[bbox_iou(a, b) for b in btlbrs] for a in atlbrs])


def bbox_ious_vec(atlbrs, btlbrs) -> np.ndarray:
    """Vectorised restatement of demo:1695-1743 in float64."""
    a = np.asarray(atlbrs, dtype=np.float64).reshape(-1, 4)
    b = np.asarray(btlbrs, dtype=np.float64).reshape(-1, 4)
    if a.shape[0] == 0 or b.shape[0] == 0:
        return np.zeros((a.shape[0], b.shape[0]), dtype=np.float32)
    ixmin = np.maximum(a[:, None, 0], b[None, :, 0])
    iymin = np.maximum(a[:, None, 1], b[None, :, 1])
    ixmax = np.minimum(a[:, None, 2], b[None, :, 2])
    iymax = np.minimum(a[:, None, 3], b[None, :, 3])
    empty = (ixmax <= ixmin) | (iymax <= iymin)
    inter = (ixmax - ixmin) * (iymax - iymin)
    area1 = ((a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1]))[:, None]
    area2 = ((b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1]))[None, :]
    with np.errstate(divide="ignore", invalid="ignore"):
        iou = inter / (area1 + area2 - inter)
    return np.where(empty, 0.0, iou)


def iou_distance(atlbrs, btlbrs, mode: str = "vectorized") -> np.ndarray:
    """demo:1745-1761 on tlbr arrays."""
    ious = bbox_ious_loop(atlbrs, btlbrs) if mode == "faithful" else bbox_ious_vec(atlbrs, btlbrs)
    return 1 - ious


# --------------------------------------------------------------------------------------
# ReID distance and cost fusion (demo:1453-1460, demo:1539-1554, demo:1593-1602)
# --------------------------------------------------------------------------------------
def embedding_distance(track_feats: np.ndarray, det_feats: np.ndarray) -> np.ndarray:
    """`1 - max(0, A @ B^T)` in float32 (demo:1599).  The in-graph variant
    (README.md:185-195, consumed at demo:1453-1460, demo:1543) is not visible in the
    reference tree: parity unpinned; a clamp cannot matter for the fused cost."""
    a = np.asarray(track_feats, dtype=np.float32)
    b = np.asarray(det_feats, dtype=np.float32)
    return (1.0 - np.maximum(0.0, np.matmul(a, b.transpose(1, 0)))).astype(np.float32)


def fuse_stage1(ious_dists: np.ndarray, body_sim: np.ndarray, face_sim: Optional[np.ndarray] = None,
                proximity: float = PROXIMITY_THRESH, appearance: float = APPEARANCE_THRESH) -> np.ndarray:
    """demo:1539-1554.  body_sim / face_sim are [tracks, dets] similarities (float32)."""
    ious_dists_mask = ious_dists > proximity
    emb_dists = 1.0 - np.asarray(body_sim, dtype=np.float32)
    if face_sim is None:
        face_sim = np.zeros_like(emb_dists)
    face_emb_dists = 1.0 - np.asarray(face_sim, dtype=np.float32)
    emb_dists_comp = np.minimum(emb_dists, face_emb_dists)
    emb_dists_mask = emb_dists_comp > appearance
    emb_dists[emb_dists_mask] = 1.0
    ious_dists_mask = np.logical_and(emb_dists_mask, ious_dists_mask)
    emb_dists[ious_dists_mask] = 1.0
    return np.minimum(ious_dists, emb_dists)


def fuse_stage3(ious_dists: np.ndarray, emb_dists: np.ndarray,
                proximity: float = PROXIMITY_THRESH, appearance: float = APPEARANCE_THRESH) -> np.ndarray:
    """demo:1599-1602 (unconfirmed tracks): appearance gate then IoU gate, then min."""
    emb = np.array(emb_dists, dtype=np.float32, copy=True)
    emb[emb > appearance] = 1.0
    emb[ious_dists > proximity] = 1.0
    return np.minimum(ious_dists, emb)


def fuse_score(iou_dists: np.ndarray, det_scores: np.ndarray) -> np.ndarray:
    """Upstream BoT-SORT name, absent from the reference: 1 - (1-d) * score."""
    if iou_dists.size == 0:
        return iou_dists
    iou_sim = 1 - iou_dists
    return 1 - iou_sim * np.asarray(det_scores)[None, :]


# --------------------------------------------------------------------------------------
# linear assignment (demo:1682-1693 -> lap.lapjv)
# --------------------------------------------------------------------------------------
_CLIB = None


def build_c(force: bool = False) -> str:
    """Compile oracle/lapjv_port.c -> oracle/_build/liboracle.so (gcc)."""
    out_dir = os.path.join(_HERE, "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "liboracle.so")
    src = os.path.join(_HERE, "lapjv_port.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", so, src])
    return so


def _clib():
    global _CLIB
    if _CLIB is None:
        lib = ctypes.CDLL(build_c())
        lib.oracle_lapjv_extended.restype = ctypes.c_int
        lib.oracle_lapjv_extended.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_double,
                                              ctypes.c_void_p, ctypes.c_void_p]
        _CLIB = lib
    return _CLIB


def lapjv_extended(cost: np.ndarray, limit: float, solver: str = "jv") -> Tuple[np.ndarray, np.ndarray]:
    """lap.lapjv(cost, extend_cost=True, cost_limit=limit) -> (x, y) with -1 = unmatched.
    solver="jv": C restatement of lap's dense Jonker-Volgenant (oracle/lapjv_port.c);
    solver="scipy": same extended matrix solved by scipy.optimize.linear_sum_assignment."""
    c = np.ascontiguousarray(cost, dtype=np.float64)
    n, m = c.shape
    if solver == "jv":
        x = np.empty(n, dtype=np.int32)
        y = np.empty(m, dtype=np.int32)
        rc = _clib().oracle_lapjv_extended(n, m, c.ctypes.data, float(limit), x.ctypes.data, y.ctypes.data)
        if rc != 0:
            raise MemoryError("oracle_lapjv_extended")
        return x, y
    from scipy.optimize import linear_sum_assignment
    k = n + m
    ext = np.empty((k, k), dtype=np.float64)
    ext[:] = limit / 2.0
    ext[n:, m:] = 0
    ext[:n, :m] = c
    ri, ci = linear_sum_assignment(ext)
    x = np.empty(k, dtype=np.int32); y = np.empty(k, dtype=np.int32)
    x[ri] = ci; y[ci] = ri
    x[x >= m] = -1; y[y >= n] = -1
    return x[:n], y[:m]


def linear_assignment(cost_matrix: np.ndarray, thresh: float, solver: str = "jv"):
    """demo:1682-1693 including its quirks (tuples in the empty branch, shape (0,) matches)."""
    if cost_matrix.size == 0:
        return (np.empty((0, 2), dtype=int), tuple(range(cost_matrix.shape[0])),
                tuple(range(cost_matrix.shape[1])))
    x, y = lapjv_extended(cost_matrix, thresh, solver)
    matches = [[ix, mx] for ix, mx in enumerate(x) if mx >= 0]
    unmatched_a = np.where(x < 0)[0]
    unmatched_b = np.where(y < 0)[0]
    return np.asarray(matches), unmatched_a, unmatched_b


def assignment_objective(cost: np.ndarray, thresh: float, x: np.ndarray) -> float:
    """Objective lap minimises, up to a constant: sum over matches of (c_ij - thresh)."""
    rows = np.nonzero(np.asarray(x) >= 0)[0]
    return float(np.sum(cost[rows, np.asarray(x)[rows]] - thresh))


# --------------------------------------------------------------------------------------
# tracker restatement (demo:439-688 STrack, demo:1252-1680 BoTSORT + list utils)
# --------------------------------------------------------------------------------------
class OTrack:
    """Array-backed restatement of STrack (demo:439-688).  Holds exactly the state the
    association and lifecycle read."""
    __slots__ = ("tlwh0", "score", "mean", "cov", "is_activated", "state", "track_id", "frame_id",
                 "start_frame", "tracklet_len", "curr_feat", "smooth_feat", "det_index", "f32_state")

    def __init__(self, tlbr_int, score, feat, det_index):
        t = np.asarray(tlbr_int)
        tlwh = t.copy()
        tlwh[2:] -= tlwh[:2]                                  # demo:676-679 on the int array (demo:1495)
        self.tlwh0 = np.asarray(tlwh, dtype=np.float32)       # demo:465
        self.score = score
        self.mean = None
        self.cov = None
        self.is_activated = False
        self.state = ST_NEW
        self.track_id = 0
        self.frame_id = 0
        self.start_frame = 0
        self.tracklet_len = 0
        self.curr_feat = None
        self.smooth_feat = None
        self.det_index = det_index
        self.f32_state = False
        if feat is not None:
            self.update_features(feat)

    def update_features(self, feat):
        """demo:492-502 (body features; in-place normalisation, aliasing on first call)."""
        self.curr_feat = feat
        if self.smooth_feat is None:
            self.smooth_feat = feat
        else:
            self.smooth_feat = ALPHA * self.smooth_feat + (1 - ALPHA) * feat
        self.smooth_feat /= np.linalg.norm(self.smooth_feat)

    @property
    def tlwh(self):
        if self.mean is None:
            return self.tlwh0.copy()
        ret = self.mean[:4].copy()
        ret[:2] -= ret[2:] / 2
        return ret

    @property
    def tlbr(self):
        ret = self.tlwh.copy()
        ret[2:] += ret[:2]
        return ret

    @staticmethod
    def tlwh_to_xywh(tlwh):
        ret = np.asarray(tlwh).copy()
        ret[:2] += ret[2:] / 2
        return ret


def _joint(a: List[OTrack], b: List[OTrack]) -> List[OTrack]:
    """demo:1642-1653."""
    exists = {}
    res = []
    for t in a:
        exists[t.track_id] = 1
        res.append(t)
    for t in b:
        if not exists.get(t.track_id, 0):
            exists[t.track_id] = 1
            res.append(t)
    return res


def _sub(a: List[OTrack], b: List[OTrack]) -> List[OTrack]:
    """demo:1655-1663."""
    d = {}
    for t in a:
        d[t.track_id] = t
    for t in b:
        if d.get(t.track_id, 0):
            del d[t.track_id]
    return list(d.values())


class OracleBoTSORT:
    """Restatement of BoTSORT.update (demo:1291-1639) driven by arrays instead of an image:
    the detector output (int boxes + scores, all class 0 = body) and the ReID features are
    given, the in-graph similarity is computed here as f_det . f_trk^T (float32).
    Face similarities are 0 (SURVEY: face encoder out of scope, term vanishes)."""

    def __init__(self, frame_rate: int = 30, mode: str = "vectorized", lap_solver: str = "jv",
                 use_features: bool = True, iou_mode: Optional[str] = None):
        self.iou_mode = iou_mode or mode       # bench.py: reference loop structure with a sampled IoU loop
        self.tracked: List[OTrack] = []
        self.lost: List[OTrack] = []
        self.removed: List[OTrack] = []
        self._count = 0                                      # BaseTrack._count, per tracker
        self.frame_id = 0
        self.max_time_lost = int(frame_rate / 30.0 * TRACK_BUFFER)
        self.mode = mode
        self.lap_solver = lap_solver
        self.use_features = use_features
        self.last = {}                                       # per-frame intermediates for tests

    # -- helpers -------------------------------------------------------------------
    def _next_id(self):
        self._count += 1
        return self._count

    def _iou_dist(self, tracks: Sequence[OTrack], dets: Sequence[OTrack]) -> np.ndarray:
        a = [t.tlbr for t in tracks]
        b = [t.tlbr for t in dets]
        return iou_distance(a, b, self.iou_mode)

    def _multi_predict(self, pool: List[OTrack]):
        """demo:524-536."""
        if len(pool) == 0:
            return
        mm = np.asarray([t.mean.copy() for t in pool])
        cc = np.asarray([t.cov for t in pool])
        for i, t in enumerate(pool):
            if t.state != ST_TRACKED:
                mm[i][6] = 0
                mm[i][7] = 0
        mm, cc = kf_multi_predict(mm, cc)
        for i, t in enumerate(pool):
            t.mean = mm[i]
            t.cov = cc[i]
            t.f32_state = False

    def _apply_matches(self, tracks: List[OTrack], dets: List[OTrack], matches, activated, refind):
        """demo:1558-1566 / 1572-1580 / 1605-1608."""
        if len(matches) == 0:
            return
        if self.mode == "faithful":
            for it, idet in matches:
                trk, det = tracks[it], dets[idet]
                z = OTrack.tlwh_to_xywh(det.tlwh)
                trk.mean, trk.cov = kf_update(trk.mean, trk.cov, z)
        else:
            mm = np.stack([tracks[it].mean for it, _ in matches])
            cc = np.stack([np.asarray(tracks[it].cov, dtype=np.float64) for it, _ in matches])
            zz = np.stack([OTrack.tlwh_to_xywh(dets[idet].tlwh) for _, idet in matches])
            f32 = np.array([tracks[it].f32_state for it, _ in matches])
            nm, nc = kf_update_batch(mm, cc, zz, f32)
            for k, (it, _) in enumerate(matches):
                tracks[it].mean = nm[k]
                tracks[it].cov = nc[k]
        for it, idet in matches:
            trk, det = tracks[it], dets[idet]
            trk.f32_state = False
            if det.curr_feat is not None:
                trk.update_features(det.curr_feat)
            if trk.state == ST_TRACKED:
                trk.tracklet_len += 1                        # demo:595
                activated.append(trk)
            else:
                trk.tracklet_len = 0                         # demo:577
                refind.append(trk)
            trk.frame_id = self.frame_id
            trk.state = ST_TRACKED
            trk.is_activated = True
            trk.score = det.score
            trk.det_index = det.det_index

    # -- the frame step ------------------------------------------------------------
    def update_arrays(self, boxes: np.ndarray, scores: np.ndarray, feats: Optional[np.ndarray],
                      face_sims: Optional[np.ndarray] = None):
        """boxes int [M,4] tlbr, scores float32 [M], feats float32 [M,D] (or None).
        face_sims: optional float32 [n_pool, M] face similarities in pool order (what the face encoder returns
        at demo:1473-1486, transposed like demo:1480; default 0 = the face encoder is out of scope)."""
        self.frame_id += 1
        activated: List[OTrack] = []
        refind: List[OTrack] = []
        lost_now: List[OTrack] = []
        removed_now: List[OTrack] = []
        m_all = len(boxes)
        if feats is not None:
            feats = np.array(feats, dtype=np.float32, copy=True)   # rows are normalised in place (demo:502)

        unconfirmed = [t for t in self.tracked if not t.is_activated]          # demo:1415-1421
        tracked = [t for t in self.tracked if t.is_activated]
        pool = _joint(tracked, self.lost)                                      # demo:1423
        self._multi_predict(pool)                                              # demo:1426

        # in-graph similarity, computed BEFORE the detection STracks normalise their rows
        if feats is not None and len(pool) > 0 and m_all > 0:
            pf = np.asarray([t.curr_feat for t in pool], dtype=np.float32)
            sims_all = np.matmul(pf, feats.T)                                  # [P, M_all] (demo:1459)
        else:
            sims_all = np.zeros((len(pool), m_all), dtype=np.float32)

        # body.score is a Python float (float(score), demo:1022): float32(0.4) > 0.4 is True
        fs = [float(v) for v in scores]
        hi_idx = [i for i in range(m_all) if fs[i] > TRACK_HIGH_THRESH]    # demo:1501
        lo_idx = [i for i in range(m_all)
                  if fs[i] <= TRACK_HIGH_THRESH and fs[i] >= TRACK_LOW_THRESH]   # demo:1531
        dets_hi = [OTrack(boxes[i], float(scores[i]), None if feats is None else feats[i], i) for i in hi_idx]
        dets_lo = [OTrack(boxes[i], float(scores[i]), None if feats is None else feats[i], i) for i in lo_idx]
        body_sim = sims_all[:, hi_idx] if len(hi_idx) else np.zeros((len(pool), 0), np.float32)
        face_sim = None
        if face_sims is not None and len(pool) > 0 and m_all > 0:
            face_sim = np.asarray(face_sims, dtype=np.float32).reshape(len(pool), m_all)[:, hi_idx]   # demo:1509-1514

        # first association (demo:1538-1566)
        ious_d = self._iou_dist(pool, dets_hi)
        dists = fuse_stage1(ious_d, body_sim, face_sim)
        self.last["dists1"] = dists
        self.last["iou1"] = ious_d
        self.last["emb1"] = 1.0 - np.asarray(body_sim, dtype=np.float32)
        self.last["scores"] = np.asarray(scores, dtype=np.float32)
        matches, u_track, u_det = linear_assignment(dists, MATCH_THRESH, self.lap_solver)
        self.last["matches1"] = np.asarray(matches).reshape(-1, 2)
        self._apply_matches(pool, dets_hi, matches, activated, refind)

        # second association (demo:1568-1586)
        r_tracked = [pool[i] for i in u_track if pool[i].state == ST_TRACKED]
        dists2 = self._iou_dist(r_tracked, dets_lo)
        self.last["dists2"] = dists2
        matches2, u_track2, _u_det2 = linear_assignment(dists2, SECOND_THRESH, self.lap_solver)
        self.last["matches2"] = np.asarray(matches2).reshape(-1, 2)
        self._apply_matches(r_tracked, dets_lo, matches2, activated, refind)
        for it in u_track2:
            trk = r_tracked[it]
            if trk.state != ST_LOST:
                trk.state = ST_LOST
                lost_now.append(trk)

        # unconfirmed (demo:1588-1612)
        u_boxes = [dets_hi[i] for i in u_det]
        ious_d3 = self._iou_dist(unconfirmed, u_boxes)
        d_feat = feats.shape[1] if feats is not None else 1
        uf = (np.asarray([t.curr_feat for t in unconfirmed], dtype=np.float32)
              if len(unconfirmed) > 0 and feats is not None else np.zeros((len(unconfirmed), d_feat), np.float32))
        bf = (np.asarray([t.curr_feat for t in u_boxes], dtype=np.float32)
              if len(u_boxes) > 0 and feats is not None else np.zeros((len(u_boxes), d_feat), np.float32))
        emb3 = embedding_distance(uf, bf)
        dists3 = fuse_stage3(ious_d3, emb3)
        self.last["dists3"] = dists3
        self.last["iou3"] = ious_d3
        self.last["emb3"] = emb3
        matches3, u_unconf, u_det3 = linear_assignment(dists3, UNCONF_THRESH, self.lap_solver)
        self.last["matches3"] = np.asarray(matches3).reshape(-1, 2)
        self._apply_matches(unconfirmed, u_boxes, matches3, activated, refind)
        for it in u_unconf:
            trk = unconfirmed[it]
            trk.state = ST_REMOVED
            removed_now.append(trk)

        # births (demo:1614-1621, activate demo:556-568)
        for inew in u_det3:
            trk = u_boxes[inew]
            if trk.score < NEW_TRACK_THRESH:
                continue
            trk.track_id = self._next_id()
            trk.mean, trk.cov = kf_initiate(OTrack.tlwh_to_xywh(trk.tlwh0))
            trk.f32_state = True
            trk.tracklet_len = 0
            trk.state = ST_TRACKED
            if self.frame_id == 1:
                trk.is_activated = True
            trk.frame_id = self.frame_id
            trk.start_frame = self.frame_id
            activated.append(trk)

        # expiry (demo:1623-1627)
        for trk in self.lost:
            if self.frame_id - trk.frame_id > self.max_time_lost:
                trk.state = ST_REMOVED
                removed_now.append(trk)

        # merge (demo:1629-1637)
        self.tracked = [t for t in self.tracked if t.state == ST_TRACKED]
        self.tracked = _joint(self.tracked, activated)
        self.tracked = _joint(self.tracked, refind)
        self.lost = _sub(self.lost, self.tracked)
        self.lost.extend(lost_now)
        self.lost = _sub(self.lost, self.removed)
        self.removed.extend(removed_now)
        self.tracked, self.lost = self._remove_duplicates(self.tracked, self.lost)
        return self.tracked

    def _remove_duplicates(self, sa: List[OTrack], sb: List[OTrack]):
        """demo:1665-1680."""
        pdist = self._iou_dist(sa, sb)
        self.last["dup_dist"] = pdist
        pairs = np.where(pdist < DUP_IOU_DIST)
        dupa, dupb = set(), set()
        for p, q in zip(*pairs):
            timep = sa[p].frame_id - sa[p].start_frame
            timeq = sb[q].frame_id - sb[q].start_frame
            if timep > timeq:
                dupb.add(q)
            else:
                dupa.add(p)
        return ([t for i, t in enumerate(sa) if i not in dupa],
                [t for i, t in enumerate(sb) if i not in dupb])

    # -- snapshot for comparisons --------------------------------------------------
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
                "det_index": np.array([t.det_index for t in lst], dtype=np.int64),
                "mean": np.array([np.asarray(t.mean, dtype=np.float64) for t in lst]).reshape(-1, 8),
                "cov": np.array([np.asarray(t.cov, dtype=np.float64) for t in lst]).reshape(-1, 8, 8),
            }
        return {"tracked": pack(self.tracked), "lost": pack(self.lost)}
