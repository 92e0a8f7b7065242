"""Host-side mirror of the reference's tracker API over the B200 kernels.

Same names, argument meaning and return types as `demo_bottrack_onnx_tflite.py` of
PINTO0309/BoT-SORT-ONNX-TensorRT (cited as demo:LINE):

    KalmanFilter (demo:118-380), TrackState / BaseTrack (demo:382-437), STrack (demo:439-688),
    BoTSORT (demo:1252-1639), joint_stracks / sub_stracks / remove_duplicate_stracks
    (demo:1642-1680), linear_assignment (demo:1682-1693), bbox_iou / bbox_ious / iou_distance
    (demo:1695-1761), plus `embedding_distance` and `fuse_score` (upstream BoT-SORT names the
    north star asks for; semantics in SURVEY.md A13/A14).

Every numeric operation goes through the C ABI (`_lib.Context`) to the sm_100a kernels; there is
no NumPy re-implementation behind these classes.  Two ways to use it:

  * object API  -- `STrack.multi_predict(list)`, `track.update(det, frame_id)`,
    `iou_distance(a, b)`, `linear_assignment(cost, thresh)`: call-for-call compatible, one small
    GPU round trip per call (parity tests read like the reference).
  * frame API   -- `BoTSORT.update(image)` / `BoTSORT.update_arrays(boxes, scores, feats)`:
    the whole frame step on the device-resident track store (one C call per frame);
    `STrack` objects handed back are lightweight views refreshed from the store.
"""
from __future__ import annotations

from collections import OrderedDict, deque
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import _lib

_default_ctx: Optional[_lib.Context] = None


def get_context(max_tracks: int = 4096, max_dets: int = 4096, feat_dim: int = 2048, device: int = 0) -> _lib.Context:
    """Shared context of the object API (created on first use; raises without a B200)."""
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = _lib.Context(max_tracks=max_tracks, max_dets=max_dets, feat_dim=feat_dim, device=device)
    return _default_ctx


def set_context(ctx: Optional[_lib.Context]) -> None:
    global _default_ctx
    _default_ctx = ctx


# --------------------------------------------------------------------------------------------
# detector data classes (demo:84-116)
# --------------------------------------------------------------------------------------------
class Box:
    def __init__(self, trackid, classid, score, x1, y1, x2, y2, cx, cy, is_used):
        self.trackid, self.classid, self.score = trackid, classid, score
        self.x1, self.y1, self.x2, self.y2, self.cx, self.cy = x1, y1, x2, y2, cx, cy
        self.is_used = is_used


class Body(Box):
    def __init__(self, *, head=None, hand1=None, hand2=None, **kw):
        super().__init__(**kw)
        self.head, self.hand1, self.hand2 = head, hand1, hand2


class Head(Box):
    def __init__(self, *, face=None, face_landmarks=None, **kw):
        super().__init__(**kw)
        self.face, self.face_landmarks = face, face_landmarks


class Face(Box):
    pass


class Hand(Box):
    pass


# --------------------------------------------------------------------------------------------
# Kalman filter (demo:118-380)
# --------------------------------------------------------------------------------------------
chi2inv95 = {1: 3.8415, 2: 5.9915, 3: 7.8147, 4: 9.4877, 5: 11.070, 6: 12.592, 7: 14.067, 8: 15.507, 9: 16.919}


class KalmanFilter:
    """8-state constant-velocity filter, xywh parametrisation; arithmetic on the GPU (fp64)."""

    def __init__(self, ctx: Optional[_lib.Context] = None):
        self._ctx = ctx
        ndim, dt = 4, 1.0
        self._motion_mat = np.eye(2 * ndim, 2 * ndim)
        for i in range(ndim):
            self._motion_mat[i, ndim + i] = dt
        self._update_mat = np.eye(ndim, 2 * ndim)
        self._std_weight_position = 1.0 / 20
        self._std_weight_velocity = 1.0 / 160

    @property
    def ctx(self) -> _lib.Context:
        return self._ctx or get_context()

    def initiate(self, measurement):
        mean, cov = self.ctx.kalman_initiate(np.asarray(measurement, dtype=np.float32).reshape(1, 4))
        return mean[0], cov[0]

    def predict(self, mean, covariance):
        m, c = self.ctx.kalman_multi_predict(np.asarray(mean)[None], np.asarray(covariance)[None])
        return m[0], c[0]

    def project(self, mean, covariance):
        pm, pc = self.ctx.kalman_project(np.asarray(mean)[None], np.asarray(covariance)[None])
        return pm[0], pc[0]

    def multi_predict(self, mean, covariance):
        mean = np.asarray(mean)
        noise_f32 = mean.dtype == np.float32      # NumPy evaluates the noise in the dtype of `mean`
        return self.ctx.kalman_multi_predict(mean, covariance, None, noise_f32=noise_f32)

    def update(self, mean, covariance, measurement):
        mean = np.asarray(mean)
        f32 = np.array([mean.dtype == np.float32], dtype=np.uint8)
        m, c = self.ctx.kalman_update(mean[None], np.asarray(covariance)[None],
                                      np.asarray(measurement, dtype=np.float64)[None], noise_f32=f32)
        return m[0], c[0]

    def gating_distance(self, mean, covariance, measurements, only_position=False, metric="maha"):
        """demo:338-380 -- never called by the reference tracker (dead code); the
        projection runs on the GPU, the K x 4 residual algebra stays on the host."""
        pm, pc = self.project(mean, covariance)
        measurements = np.asarray(measurements, dtype=np.float64)
        if only_position:
            pm, pc = pm[:2], pc[:2, :2]
            measurements = measurements[:, :2]
        d = measurements - pm
        if metric == "gaussian":
            return np.sum(d * d, axis=1)
        if metric == "maha":
            chol = np.linalg.cholesky(pc)
            z = np.linalg.solve(chol, d.T)
            return np.sum(z * z, axis=0)
        raise ValueError("invalid distance metric")


# --------------------------------------------------------------------------------------------
# track objects (demo:382-688)
# --------------------------------------------------------------------------------------------
class TrackState:
    New = 0
    Tracked = 1
    Lost = 2
    LongLost = 3
    Removed = 4


class BaseTrack:
    _count = 0
    track_id = 0
    is_activated = False
    state = TrackState.New
    history = OrderedDict()
    features: list = []
    body_curr_feature = None
    face_curr_feature = None
    score = 0
    start_frame = 0
    frame_id = 0
    time_since_update = 0
    location = (np.inf, np.inf)

    @property
    def end_frame(self):
        return self.frame_id

    @staticmethod
    def next_id():
        BaseTrack._count += 1
        return BaseTrack._count

    def activate(self, *args):
        raise NotImplementedError

    def predict(self):
        raise NotImplementedError

    def update(self, *args, **kwargs):
        raise NotImplementedError

    def mark_lost(self):
        self.state = TrackState.Lost

    def mark_long_lost(self):
        self.state = TrackState.LongLost

    def mark_removed(self):
        self.state = TrackState.Removed

    @staticmethod
    def clear_count():
        BaseTrack._count = 0


class STrack(BaseTrack):
    shared_kalman = KalmanFilter()

    def __init__(self, tlwh, score, feature_history, body, body_feature=None, face_feature=None):
        self._tlwh = np.asarray(tlwh, dtype=np.float32)
        self.kalman_filter: Optional[KalmanFilter] = None
        self.mean, self.covariance = None, None
        self.is_activated = False
        self.score = score
        self.tracklet_len = 0
        self.alpha = 0.9
        self.feature_history = feature_history
        self.body = body
        self.body_smooth_feature = None
        self.body_curr_feature = None
        self.body_features = deque([], maxlen=feature_history)
        if body_feature is not None:
            self.update_body_features(body_feature)
        self.face_smooth_feature = None
        self.face_curr_feature = None
        self.face_features = deque([], maxlen=feature_history)
        if face_feature is not None:
            self.update_face_features(face_feature)

    # -- features (demo:492-514): EMA + renormalisation on the GPU -------------------------
    def _ema(self, smooth, feature):
        ctx = STrack.shared_kalman.ctx
        f = np.asarray(feature, dtype=np.float32).reshape(1, -1)
        if smooth is None:
            s, _ = ctx.feature_ema(np.zeros_like(f), np.zeros_like(f), f, first=np.ones(1, np.uint8), alpha=self.alpha)
        else:
            s, _ = ctx.feature_ema(np.asarray(smooth, np.float32).reshape(1, -1), np.zeros_like(f), f, alpha=self.alpha)
        return s[0]

    def update_body_features(self, feature):
        first = self.body_smooth_feature is None
        self.body_smooth_feature = self._ema(self.body_smooth_feature, feature)
        # the reference normalises the caller's array in place on the first call (aliasing, demo:497-502)
        self.body_curr_feature = self.body_smooth_feature if first else feature
        self.body_features.append(self.body_curr_feature)

    def update_face_features(self, feature):
        first = self.face_smooth_feature is None
        self.face_smooth_feature = self._ema(self.face_smooth_feature, feature)
        self.face_curr_feature = self.face_smooth_feature if first else feature
        self.face_features.append(self.face_curr_feature)

    # -- Kalman -------------------------------------------------------------------------------
    def predict(self):
        mean_state = self.mean.copy()
        if self.state != TrackState.Tracked:
            mean_state[6] = 0
            mean_state[7] = 0
        self.mean, self.covariance = self.kalman_filter.predict(mean_state, self.covariance)

    @staticmethod
    def multi_predict(stracks: Sequence["STrack"]):
        """demo:524-536: one batched GPU call for the whole list (velocity reset fused in-kernel)."""
        if len(stracks) == 0:
            return
        multi_mean = np.asarray([st.mean.copy() for st in stracks])
        multi_cov = np.asarray([st.covariance for st in stracks])
        state = np.asarray([st.state for st in stracks], dtype=np.int32)
        mean, cov = STrack.shared_kalman.ctx.kalman_multi_predict(
            multi_mean, multi_cov, state, noise_f32=(multi_mean.dtype == np.float32))
        for i, st in enumerate(stracks):
            st.mean = mean[i]
            st.covariance = cov[i]

    @staticmethod
    def multi_gmc(stracks, H=np.eye(2, 3)):
        """demo:538-554 -- camera-motion compensation; its only call site is commented out in the
        reference (demo:1534-1536).  Out of the hot path: host arithmetic."""
        if len(stracks) > 0:
            R = H[:2, :2]
            R8x8 = np.kron(np.eye(4, dtype=float), R)
            t = H[:2, 2]
            for st in stracks:
                mean = R8x8.dot(st.mean)
                mean[:2] += t
                st.mean = mean
                st.covariance = R8x8.dot(st.covariance).dot(R8x8.transpose())

    def activate(self, kalman_filter: KalmanFilter, frame_id: int):
        self.kalman_filter = kalman_filter
        self.track_id = self.next_id()
        self.mean, self.covariance = self.kalman_filter.initiate(self.tlwh_to_xywh(self._tlwh))
        # the reference's just-initiated state is float32 under NumPy >= 2; keep the
        # dtype so later noise terms follow the same promotion
        self.mean = self.mean.astype(np.float32)
        self.covariance = self.covariance.astype(np.float32)
        self.tracklet_len = 0
        self.state = TrackState.Tracked
        if frame_id == 1:
            self.is_activated = True
        self.frame_id = frame_id
        self.start_frame = frame_id

    def re_activate(self, new_track: "STrack", frame_id: int, new_id: bool = False):
        self.mean, self.covariance = self.kalman_filter.update(self.mean, self.covariance,
                                                               self.tlwh_to_xywh(new_track.tlwh))
        if new_track.body_curr_feature is not None:
            self.update_body_features(new_track.body_curr_feature)
        if new_track.face_curr_feature is not None:
            self.update_face_features(new_track.face_curr_feature)
        self.tracklet_len = 0
        self.state = TrackState.Tracked
        self.is_activated = True
        self.frame_id = frame_id
        if new_id:
            self.track_id = self.next_id()
        self.score = new_track.score
        self.body = new_track.body

    def update(self, new_track: "STrack", frame_id: int):
        self.frame_id = frame_id
        self.tracklet_len += 1
        self.mean, self.covariance = self.kalman_filter.update(self.mean, self.covariance,
                                                               self.tlwh_to_xywh(new_track.tlwh))
        if new_track.body_curr_feature is not None:
            self.update_body_features(new_track.body_curr_feature)
        if new_track.face_curr_feature is not None:
            self.update_face_features(new_track.face_curr_feature)
        self.state = TrackState.Tracked
        self.is_activated = True
        self.score = new_track.score
        self.body = new_track.body

    def propagate_trackid_to_related_objects(self):
        b = self.body
        if b is not None:
            b.trackid = self.track_id
            head = getattr(b, "head", None)
            if head is not None:
                head.trackid = self.track_id
                if getattr(head, "face", None) is not None:
                    head.face.trackid = self.track_id
            for hand in (getattr(b, "hand1", None), getattr(b, "hand2", None)):
                if hand is not None:
                    hand.trackid = self.track_id

    # -- box formats (demo:624-685) ------------------------------------------------------------
    @property
    def tlwh(self):
        if self.mean is None:
            return self._tlwh.copy()
        ret = self.mean[:4].copy()
        ret[:2] -= ret[2:] / 2
        return ret

    @property
    def tlbr(self):
        ret = self.tlwh.copy()
        ret[2:] += ret[:2]
        return ret

    @property
    def xywh(self):
        ret = self.tlwh.copy()
        ret[:2] += ret[2:] / 2.0
        return ret

    @staticmethod
    def tlwh_to_xyah(tlwh):
        ret = np.asarray(tlwh).copy()
        ret[:2] += ret[2:] / 2
        ret[2] /= ret[3]
        return ret

    @staticmethod
    def tlwh_to_xywh(tlwh):
        ret = np.asarray(tlwh).copy()
        ret[:2] += ret[2:] / 2
        return ret

    def to_xywh(self):
        return self.tlwh_to_xywh(self.tlwh)

    @staticmethod
    def tlbr_to_tlwh(tlbr):
        ret = np.asarray(tlbr).copy()
        ret[2:] -= ret[:2]
        return ret

    @staticmethod
    def tlwh_to_tlbr(tlwh):
        ret = np.asarray(tlwh).copy()
        ret[2:] += ret[:2]
        return ret

    def __repr__(self):
        return "OT_{}_({}-{})".format(self.track_id, self.start_frame, self.end_frame)


class STrackView(STrack):
    """An STrack whose numbers live in the device track store of a `BoTSORT`; the tracker
    refreshes `track_id/state/score/...` and the box every frame, `mean`/`covariance` and the
    feature banks are fetched on first access."""

    def __init__(self, owner: "BoTSORT"):
        self._owner = owner
        self._tlbr = np.zeros(4)
        self._lazy: Dict[str, np.ndarray] = {}
        self.body = None
        self.tracklet_len = 0
        self.alpha = 0.9
        self.kalman_filter = owner.kalman_filter
        self._list = 0
        self._pos = 0

    @property
    def tlbr(self):
        return self._tlbr.copy()

    @property
    def tlwh(self):
        ret = self._tlbr.copy()
        ret[2:] -= ret[:2]
        return ret

    def _fetch(self, key):
        if key not in self._lazy:
            self._owner._fill_lazy(self._list)
        return self._lazy[key]

    def _read_only(name):
        def setter(self, value):
            raise AttributeError(
                f"STrackView.{name} is a read-only view of the device track store: the BoTSORT that owns the track "
                f"updates it on the GPU every frame.  Copy the numbers into a plain STrack (or call the KalmanFilter / "
                f"matching functions on arrays) to edit them on the host.")
        return setter

    mean = property(lambda self: self._fetch("mean"), _read_only("mean"))
    covariance = property(lambda self: self._fetch("cov"), _read_only("covariance"))
    body_curr_feature = property(lambda self: self._fetch("curr"), _read_only("body_curr_feature"))
    body_smooth_feature = property(lambda self: self._fetch("smooth"), _read_only("body_smooth_feature"))
    del _read_only


# --------------------------------------------------------------------------------------------
# matching primitives (demo:1682-1761)
# --------------------------------------------------------------------------------------------
def joint_stracks(tlista, tlistb):
    exists, res = {}, []
    for t in tlista:
        exists[t.track_id] = 1
        res.append(t)
    for t in tlistb:
        if not exists.get(t.track_id, 0):
            exists[t.track_id] = 1
            res.append(t)
    return res


def sub_stracks(tlista, tlistb):
    stracks = {t.track_id: t for t in tlista}
    for t in tlistb:
        if stracks.get(t.track_id, 0):
            del stracks[t.track_id]
    return list(stracks.values())


def bbox_ious(atlbrs, btlbrs):
    """demo:1731-1743: N x M IoU on the GPU (float64); empty -> float32 zeros like the reference."""
    ious = np.zeros((len(atlbrs), len(btlbrs)), dtype=np.float32)
    if ious.size == 0:
        return ious
    return 1.0 - get_context().iou_distance(np.asarray(atlbrs, dtype=np.float64), np.asarray(btlbrs, dtype=np.float64))


def bbox_iou(box1, box2):
    """demo:1695-1713 (scalar): a 1 x 1 launch of the same kernel."""
    return float(bbox_ious([box1], [box2])[0, 0])


def iou_distance(atracks, btracks):
    """demo:1745-1761: accepts lists of STrack (uses .tlbr) or lists of arrays."""
    if (len(atracks) > 0 and isinstance(atracks[0], np.ndarray)) or (len(btracks) > 0 and isinstance(btracks[0], np.ndarray)):
        atlbrs, btlbrs = atracks, btracks
    else:
        atlbrs = [t.tlbr for t in atracks]
        btlbrs = [t.tlbr for t in btracks]
    return 1 - bbox_ious(atlbrs, btlbrs)


def embedding_distance(tracks, detections, precision: int = 0):
    """`1 - max(0, f_trk . f_det^T)` [N_trk, M_det] float32 (demo:1599; in-graph cosine consumed at
    demo:1453-1460).  Accepts STrack lists (uses body_curr_feature) or 2-D arrays."""
    def feats(x):
        if isinstance(x, np.ndarray):
            return np.asarray(x, dtype=np.float32)
        return np.asarray([t.body_curr_feature for t in x], dtype=np.float32)
    a, b = feats(tracks), feats(detections)
    if a.size == 0 or b.size == 0:
        return np.zeros((len(a), len(b)), dtype=np.float32)
    ctx = get_context()
    if precision == 0 and a.shape[1] % 64 != 0:
        precision = 1
    return ctx.embedding_distance(a, b, precision=precision)


def fuse_score(cost_matrix, detections):
    """Upstream BoT-SORT `fuse_score` (absent from the reference)."""
    cost_matrix = np.asarray(cost_matrix, dtype=np.float64)
    if cost_matrix.size == 0:
        return cost_matrix
    scores = np.asarray([d.score if hasattr(d, "score") else d for d in detections], dtype=np.float64)
    return get_context().fuse_score(cost_matrix, scores)


def linear_assignment(cost_matrix, thresh):
    """demo:1682-1693 (lap.lapjv(extend_cost=True, cost_limit=thresh)) on the GPU solver, including
    the reference's return quirks (tuples in the empty branch, shape-(0,) matches array)."""
    cost_matrix = np.asarray(cost_matrix)
    if cost_matrix.size == 0:
        return (np.empty((0, 2), dtype=int), tuple(range(cost_matrix.shape[0])), tuple(range(cost_matrix.shape[1])))
    x, y = get_context().lapjv(cost_matrix, thresh)
    matches = [[ix, mx] for ix, mx in enumerate(x) if mx >= 0]
    unmatched_a = np.where(x < 0)[0]
    unmatched_b = np.where(y < 0)[0]
    return np.asarray(matches), unmatched_a, unmatched_b


def remove_duplicate_stracks(stracksa, stracksb):
    pdist = iou_distance(stracksa, stracksb)
    pairs = np.where(pdist < 0.15)
    dupa, dupb = [], []
    for p, q in zip(*pairs):
        timep = stracksa[p].frame_id - stracksa[p].start_frame
        timeq = stracksb[q].frame_id - stracksb[q].start_frame
        if timep > timeq:
            dupb.append(q)
        else:
            dupa.append(p)
    resa = [t for i, t in enumerate(stracksa) if i not in dupa]
    resb = [t for i, t in enumerate(stracksb) if i not in dupb]
    return resa, resb


# --------------------------------------------------------------------------------------------
# part grouping (demo:1372-1411, demo:1715-1729, demo:1763-1791) -- host glue, <= 50 boxes/class
# --------------------------------------------------------------------------------------------
def _iou_box(a: Box, b: Box) -> float:
    ix1, iy1 = max(a.x1, b.x1), max(a.y1, b.y1)
    ix2, iy2 = min(a.x2, b.x2), min(a.y2, b.y2)
    if ix2 <= ix1 or iy2 <= iy1:
        return 0.0
    inter = (ix2 - ix1) * (iy2 - iy1)
    return inter / float((a.x2 - a.x1) * (a.y2 - a.y1) + (b.x2 - b.x1) * (b.y2 - b.y1) - inter)


def find_most_relevant_object(base_obj: Box, target_objs: List[Box]):
    """Greedy pick: highest IoU, ties by centre distance; marks the winner used (order-dependent)."""
    best, best_iou, best_dist = None, 0.0, float("inf")
    for cand in target_objs:
        if cand is None or cand.is_used:
            continue
        iou = _iou_box(base_obj, cand)
        dist = ((base_obj.cx - cand.cx) ** 2 + (base_obj.cy - cand.cy) ** 2) ** 0.5
        if iou > best_iou:
            best, best_iou, best_dist = cand, iou, dist
        elif iou > 0.0 and iou == best_iou and dist < best_dist:
            best, best_dist = cand, dist
    if best:
        best.is_used = True
    return best


def group_parts(detected_boxes: Sequence[Box]):
    """Split detector output by class (0 body, 1 head, 2 hand, 3 face) and attach face->head,
    head->body, two hands->body, in the reference's order."""
    common = lambda b: dict(trackid=0, classid=b.classid, score=b.score, x1=b.x1, y1=b.y1, x2=b.x2, y2=b.y2,
                            cx=b.cx, cy=b.cy, is_used=False)
    bodies = [Body(**common(b)) for b in detected_boxes if b.classid == 0]
    heads = [Head(**common(b)) for b in detected_boxes if b.classid == 1]
    hands = [Hand(**common(b)) for b in detected_boxes if b.classid == 2]
    faces = [Face(**common(b)) for b in detected_boxes if b.classid == 3]
    if faces:
        for h in heads:
            f = find_most_relevant_object(h, faces)
            if f is not None:
                h.face = f
    if heads:
        for b in bodies:
            h = find_most_relevant_object(b, heads)
            if h is not None:
                b.head = h
    if hands:
        for b in bodies:
            for attr in ("hand1", "hand2"):
                h = find_most_relevant_object(b, hands)
                if h is not None:
                    setattr(b, attr, h)
    return bodies


# --------------------------------------------------------------------------------------------
# the tracker (demo:1252-1639)
# --------------------------------------------------------------------------------------------
class BoTSORT:
    """Drop-in for the reference's BoTSORT.  `update(image)` keeps the reference contract
    (detector and encoder are user callables: ORT/TensorRT sessions in production, stubs in
    tests); the tracker state and all arithmetic live on the GPU."""

    def __init__(self, object_detection_model, body_feature_extractor_model, face_feature_extractor_model,
                 frame_rate: int = 30, ctx: Optional[_lib.Context] = None, max_tracks: int = 4096,
                 max_dets: int = 4096, feat_dim: Optional[int] = None, device: int = 0):
        self.tracked_stracks: List[STrack] = []
        self.lost_stracks: List[STrack] = []
        self.removed_stracks: List[STrack] = []
        BaseTrack.clear_count()
        self.frame_id = 0
        self.track_high_thresh = 0.40
        self.track_low_thresh = 0.1
        self.new_track_thresh = 0.9
        self.match_thresh = 0.8
        self.track_buffer = 300
        self.feature_history = 300
        self.proximity_thresh = 0.5
        self.appearance_thresh = 0.25
        self.buffer_size = int(frame_rate / 30.0 * self.track_buffer)
        self.max_time_lost = self.buffer_size
        self.detector = object_detection_model
        self.body_encoder = body_feature_extractor_model
        self.face_encoder = face_feature_extractor_model
        if feat_dim is None:
            feat_dim = int(getattr(body_feature_extractor_model, "feature_size", 2048) or 2048)
        self._own_ctx = ctx is None
        self.ctx = ctx or _lib.Context(max_tracks=max_tracks, max_dets=max_dets, feat_dim=feat_dim, device=device)
        self.kalman_filter = KalmanFilter(self.ctx)
        cfg = self.ctx.default_config()
        cfg.frame_rate = int(frame_rate)
        cfg.with_reid = 0 if body_feature_extractor_model is None else 1
        self._cfg = cfg
        self.ctx.tracker_reset(cfg)
        self._views: Dict[int, STrackView] = {}
        self._face_curr: Dict[int, np.ndarray] = {}      # track id -> face_curr_feature (host: the face model is the user's)
        self.last_info: dict = {}

    # -- frame API on arrays ------------------------------------------------------------------
    def update_arrays(self, boxes, scores, feats=None, bodies: Optional[Sequence[Box]] = None,
                      face_sim=None) -> List[STrack]:
        """One BoTSORT.update on detector/encoder outputs (class-0 boxes, int tlbr).  face_sim: None or the
        float32 [n_pool, m] face similarity matrix of demo:1465-1486 (rows = `strack_pool()` order)."""
        self.frame_id += 1
        if not self._cfg.with_reid:
            feats = None
        self.last_info = self.ctx.update_arrays(boxes, scores, feats, face_sim=face_sim)
        self._refresh_views(bodies)
        return self.tracked_stracks

    def strack_pool(self) -> List[STrack]:
        """The reference's pool of this frame (demo:1415-1423): activated tracked tracks, then lost tracks."""
        return [t for t in self.tracked_stracks if t.is_activated] + list(self.lost_stracks)

    def _refresh_views(self, bodies):
        lists = []
        for which in (0, 1):
            tr = self.ctx.get_tracks(which)
            out = []
            for i in range(len(tr["ids"])):
                tid = int(tr["ids"][i])
                v = self._views.get(tid)
                if v is None:
                    v = self._views[tid] = STrackView(self)
                v.track_id = tid
                v.state = int(tr["state"][i])
                v.is_activated = bool(tr["activated"][i])
                v.frame_id = int(tr["frame_id"][i])
                v.start_frame = int(tr["start_frame"][i])
                v.tracklet_len = int(tr["tracklet_len"][i])
                v.score = float(tr["score"][i])
                v._tlbr = tr["tlbr"][i].copy()
                v._lazy = {}
                v._list, v._pos = which, i
                v._det_index = int(tr["det_index"][i])
                if bodies is not None and v.frame_id == self.frame_id and 0 <= tr["det_index"][i] < len(bodies):
                    v.body = bodies[int(tr["det_index"][i])]
                out.append(v)
            lists.append(out)
        live = {v.track_id for lst in lists for v in lst}
        for tid in [t for t in self._views if t not in live]:
            gone = self._views.pop(tid)
            gone.state = TrackState.Removed
            self.removed_stracks.append(gone)
        self.tracked_stracks, self.lost_stracks = lists
        for v in self.tracked_stracks:
            v.propagate_trackid_to_related_objects()

    def _fill_lazy(self, which):
        tr = self.ctx.get_tracks(which, with_state=True)
        lst = self.tracked_stracks if which == 0 else self.lost_stracks
        curr = smooth = None
        if self._cfg.with_reid:
            curr, smooth = self.ctx.get_track_features(which)
        for i, v in enumerate(lst):
            v._lazy = {"mean": tr["mean"][i], "cov": tr["cov"][i]}
            if curr is not None:
                v._lazy["curr"], v._lazy["smooth"] = curr[i], smooth[i]

    # -- the reference entry point ------------------------------------------------------------
    def update(self, image: np.ndarray) -> List[STrack]:
        """demo:1291-1639: detect -> group parts -> crops -> body encoder -> device frame step."""
        detected = self.detector(image=image)
        bodies = group_parts(detected)
        m = len(bodies)
        boxes = np.array([[b.x1, b.y1, b.x2, b.y2] for b in bodies], dtype=np.int32).reshape(m, 4)
        scores = np.array([b.score for b in bodies], dtype=np.float32)
        feats = None
        if self._cfg.with_reid and m > 0:
            crops = [image[b.y1:b.y2, b.x1:b.x2, :] for b in bodies]                 # demo:1434-1436
            d = self.ctx.feat_dim
            # the similarity GEMM runs on the device against the track store, so the encoder is
            # not asked to compare against previous features (an all-zero target, demo:1445-1448)
            out = self.body_encoder(base_images=crops, target_features=[np.zeros((d,), np.float32)])
            feats = np.asarray(out[1], dtype=np.float32).reshape(m, d)
        elif self._cfg.with_reid:
            feats = np.zeros((0, self.ctx.feat_dim), np.float32)
        face_sim, face_feats = None, None
        if self.face_encoder is not None and self._cfg.with_reid and m > 0:
            face_sim, face_feats = self._face_term(image, bodies)
        out = self.update_arrays(boxes, scores, feats, bodies=bodies, face_sim=face_sim)
        if face_feats is not None:
            self._adopt_face_features(face_feats)
        return out

    # -- face similarity term (demo:1437-1441, demo:1465-1486, demo:1541-1546) ---------------------------------
    def _face_term(self, image, bodies):
        """Face crops (or the reference's all-zero stand-in) -> user face encoder against the pool's current face
        features -> [n_pool, m] similarity matrix.  The encoder returns (similarities[N, M], features[N, F]); the
        reference reads them swapped for this model (demo:1478-1480), which is kept: a face encoder written for
        the reference returns (features, similarities)."""
        enc = self.face_encoder
        fs = int(getattr(enc, "feature_size", 256) or 256)
        shape = getattr(enc, "_input_shapes", [[1, 3, 128, 128]])[0][1:]
        blank = np.zeros([d if isinstance(d, int) else 1 for d in shape], dtype=np.float32).transpose(1, 2, 0)
        crops = [image[b.head.face.y1:b.head.face.y2, b.head.face.x1:b.head.face.x2, :]
                 if (b.head is not None and b.head.face is not None) else blank for b in bodies]
        pool = self.strack_pool()
        none = np.zeros(fs, dtype=np.float32)
        targets = [self._face_curr.get(t.track_id, none) for t in pool] if pool else np.zeros([0, fs], dtype=np.float32)
        res = enc(base_images=crops, target_features=targets)
        sims = np.asarray(res[1], dtype=np.float32).reshape(len(bodies), len(pool)).transpose(1, 0).copy()   # [M, N]
        feats = np.asarray(res[0], dtype=np.float32).reshape(len(bodies), -1)
        sims[np.isclose(sims, 0.9999999, atol=1e-08, rtol=1e-08)] = 0.0                                       # demo:1482-1483
        return (sims if len(pool) else None), feats

    def _adopt_face_features(self, face_feats):
        """STrack.update_face_features (demo:504-514) for the tracks that took a detection this frame: the
        detection's STrack normalised its face row in place when it was built, so that is what the track keeps."""
        live = set()
        for lst in (self.tracked_stracks, self.lost_stracks):
            for t in lst:
                live.add(t.track_id)
        for t in self.tracked_stracks:
            j = getattr(t, "_det_index", -1)
            if t.frame_id == self.frame_id and 0 <= j < len(face_feats):
                f = face_feats[j].copy()
                f /= np.linalg.norm(f)
                self._face_curr[t.track_id] = f
        for tid in [k for k in self._face_curr if k not in live]:
            del self._face_curr[tid]

    def close(self):
        if self._own_ctx:
            self.ctx.close()


class _Matching:
    """`matching` namespace of upstream BoT-SORT that the north star names."""
    iou_distance = staticmethod(iou_distance)
    embedding_distance = staticmethod(embedding_distance)
    fuse_score = staticmethod(fuse_score)
    linear_assignment = staticmethod(linear_assignment)
    bbox_ious = staticmethod(bbox_ious)


matching = _Matching()
