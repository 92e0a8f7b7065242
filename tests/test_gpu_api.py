"""The host-side mirror of the reference API (botsort_b200.tracker) on the GPU: the tests read
like calls into demo_bottrack_onnx_tflite.py and are checked against the oracle."""
import numpy as np
import pytest

from oracle import oracle_np as O
from botsort_b200.synthetic import SceneConfig, SyntheticScene

pytestmark = pytest.mark.gpu


@pytest.fixture()
def api(ctx):
    from botsort_b200 import tracker as T
    T.set_context(ctx)
    T.STrack.shared_kalman = T.KalmanFilter(ctx)
    yield T
    T.set_context(None)


def test_kalman_filter_object(api):
    kf = api.KalmanFilter()
    z = np.array([100.5, 200.0, 40.0, 90.0], dtype=np.float32)
    mean, cov = kf.initiate(z)
    om, oc = O.kf_initiate(z)
    np.testing.assert_array_equal(mean, om.astype(np.float64))
    np.testing.assert_array_equal(cov, oc.astype(np.float64))
    m1, c1 = kf.predict(mean, cov)
    rm, rc = O.kf_multi_predict(mean[None], cov[None])
    assert np.max(np.abs(m1 - rm[0])) <= 1e-10 and np.max(np.abs(c1 - rc[0])) <= 1e-10
    pm, pc = kf.project(m1, c1)
    qm, qc = O.kf_project(m1, c1)
    assert np.max(np.abs(pm - qm)) <= 1e-10 and np.max(np.abs(pc - qc)) <= 1e-10
    m2, c2 = kf.update(m1, c1, np.array([102.0, 203.0, 41.0, 88.0]))
    um, uc = O.kf_update(m1, c1, np.array([102.0, 203.0, 41.0, 88.0]))
    assert np.max(np.abs(m2 - um)) <= 1e-9 and np.max(np.abs(c2 - uc)) <= 1e-9
    with pytest.raises(ValueError):
        kf.gating_distance(m1, c1, np.zeros((2, 4)), metric="bogus")
    assert kf.gating_distance(m1, c1, np.array([[102.0, 203.0, 41.0, 88.0]])).shape == (1,)


def test_strack_lifecycle_and_multi_predict(api):
    api.BaseTrack.clear_count()
    kf = api.KalmanFilter()
    rng = np.random.default_rng(0)
    f = rng.standard_normal((3, 2048)).astype(np.float32)
    f /= np.linalg.norm(f, axis=1, keepdims=True)
    a = api.STrack(api.STrack.tlbr_to_tlwh(np.array([10, 20, 60, 140])), 0.95, 300, None, body_feature=f[0].copy())
    b = api.STrack(api.STrack.tlbr_to_tlwh(np.array([200, 50, 260, 190])), 0.93, 300, None, body_feature=f[1].copy())
    a.activate(kf, 1); b.activate(kf, 1)
    assert (a.track_id, b.track_id) == (1, 2) and a.is_activated and a.state == api.TrackState.Tracked
    b.mark_lost()
    m0 = [t.mean.copy() for t in (a, b)]
    c0 = [t.covariance.copy() for t in (a, b)]
    api.STrack.multi_predict([a, b])
    ref_in = np.asarray(m0); ref_in[1, 6:] = 0
    rm, rc = O.kf_multi_predict(ref_in, np.asarray(c0))
    assert np.max(np.abs(np.asarray([a.mean, b.mean]) - rm)) <= 1e-9
    assert np.max(np.abs(np.asarray([a.covariance, b.covariance]) - rc)) <= 1e-9
    api.STrack.multi_predict([])                                 # no-op on an empty list
    det = api.STrack(api.STrack.tlbr_to_tlwh(np.array([12, 22, 63, 141])), 0.9, 300, None, body_feature=f[2].copy())
    pm, pc = a.mean.copy(), a.covariance.copy()
    a.update(det, 2)
    um, uc = O.kf_update(pm, pc, api.STrack.tlwh_to_xywh(det.tlwh))
    assert np.max(np.abs(a.mean - um)) <= 1e-9 and np.max(np.abs(a.covariance - uc)) <= 1e-9
    assert a.tracklet_len == 1 and a.frame_id == 2 and a.score == 0.9
    ref_s = 0.9 * f[0] + 0.1 * f[2]
    ref_s /= np.linalg.norm(ref_s)
    assert np.max(np.abs(a.body_smooth_feature - ref_s)) <= 1e-6
    b.re_activate(det, 3)
    assert b.state == api.TrackState.Tracked and b.tracklet_len == 0 and b.track_id == 2
    np.testing.assert_allclose(a.tlbr, a.tlwh_to_tlbr(a.tlwh))


def test_matching_functions(api):
    rng = np.random.default_rng(1)
    a = rng.uniform(0, 300, (30, 2)); a = np.hstack([a, a + rng.uniform(20, 120, (30, 2))])
    b = np.floor(rng.uniform(0, 300, (25, 2))); b = np.hstack([b, b + np.floor(rng.uniform(20, 120, (25, 2)))])
    d = api.iou_distance(list(a), list(b))
    assert np.max(np.abs(d - O.iou_distance(a, b))) <= 1e-12
    assert abs(api.bbox_iou(a[0], b[0]) - O.bbox_iou(a[0], b[0])) <= 1e-12
    assert api.iou_distance([], list(b)).shape == (0, 25) and api.bbox_ious([], []).dtype == np.float32
    for thresh in (0.8, 0.5):
        m, ua, ub = api.linear_assignment(d, thresh)
        om, oua, oub = O.linear_assignment(d, thresh, "jv")
        np.testing.assert_array_equal(m, om); np.testing.assert_array_equal(ua, oua); np.testing.assert_array_equal(ub, oub)
    m, ua, ub = api.linear_assignment(np.zeros((0, 3)), 0.8)
    assert m.shape == (0, 2) and ua == () and ub == (0, 1, 2)
    m, ua, ub = api.linear_assignment(np.ones((2, 2)), 0.8)
    assert m.shape == (0,)
    f1 = rng.standard_normal((20, 2048)).astype(np.float32); f1 /= np.linalg.norm(f1, axis=1, keepdims=True)
    f2 = rng.standard_normal((31, 2048)).astype(np.float32); f2 /= np.linalg.norm(f2, axis=1, keepdims=True)
    e = api.matching.embedding_distance(f1, f2)
    assert np.max(np.abs(e - O.embedding_distance(f1, f2))) <= 1e-4
    s = rng.uniform(0.1, 1, 25)
    np.testing.assert_allclose(api.matching.fuse_score(d, s), O.fuse_score(d, s), atol=1e-14)


class _Det:
    def __init__(self, api):
        self.api, self.boxes, self.scores, self.extra = api, None, None, []

    def __call__(self, image):
        out = [self.api.Box(trackid=0, classid=0, score=float(s), x1=int(b[0]), y1=int(b[1]), x2=int(b[2]), y2=int(b[3]),
                            cx=int((b[0] + b[2]) // 2), cy=int((b[1] + b[3]) // 2), is_used=False)
               for b, s in zip(self.boxes, self.scores)]
        return out + list(self.extra)


class _Enc:
    feature_size = 2048

    def __init__(self):
        self.feats = None

    def __call__(self, *, base_images, target_features):
        assert len(base_images) == len(self.feats)
        return np.zeros((len(self.feats), len(target_features)), np.float32), self.feats


def test_botsort_update_with_stub_models(api):
    """BoTSORT(detector, body_encoder, face_encoder).update(image) -> List[STrack], like demo:2093-2130."""
    det, enc = _Det(api), _Enc()
    trk = api.BoTSORT(det, enc, None, frame_rate=30, max_tracks=256, max_dets=256)
    oracle = O.OracleBoTSORT()
    scene = SyntheticScene(SceneConfig(n_ids=48, feat_dim=2048, seed=4, low_frac=0.1, drop_frac=0.1, newcomer_every=3))
    image = np.zeros((32, 32, 3), np.uint8)
    try:
        for k in range(15):
            fr = scene.next_frame()
            det.boxes, det.scores, enc.feats = fr["boxes"], fr["scores"], fr["feats"]
            if k == 3:      # a head + face near the first body: grouped on the host, ids propagated
                b = fr["boxes"][0]
                det.extra = [api.Box(0, 1, 0.9, int(b[0]), int(b[1]), int(b[2]), int(b[1] + 20), int(b[0]), int(b[1]), False),
                             api.Box(0, 3, 0.8, int(b[0]) + 2, int(b[1]) + 2, int(b[2]) - 2, int(b[1] + 15), int(b[0]), int(b[1]), False)]
            else:
                det.extra = []
            out = trk.update(image)
            oracle.update_arrays(fr["boxes"], fr["scores"], fr["feats"])
            snap = oracle.snapshot()
            assert [t.track_id for t in out] == list(snap["tracked"]["ids"])
            assert [t.track_id for t in trk.lost_stracks] == list(snap["lost"]["ids"])
            if len(out):
                assert np.max(np.abs(np.array([t.tlbr for t in out]) - snap["tracked"]["tlbr"])) <= 1e-4
                assert [t.is_activated for t in out] == list(snap["tracked"]["activated"])
            if k == 3:
                withhead = [t for t in out if t.body is not None and t.body.head is not None]
                assert len(withhead) == 1 and withhead[0].body.head.trackid == withhead[0].track_id
                assert withhead[0].body.head.face is not None
        t0 = out[0]
        assert t0.mean.shape == (8,) and t0.covariance.shape == (8, 8)
        assert np.max(np.abs(t0.mean - snap["tracked"]["mean"][0])) <= 1e-4
        assert abs(np.linalg.norm(t0.body_smooth_feature) - 1) <= 1e-5
    finally:
        trk.close()


class _FaceEnc:
    """Face encoder stub with the reference's contract (demo:1207-1209 returns (similarities, base_features); the
    tracker reads them swapped, demo:1478-1480, so a model built for the reference returns (features, similarities))."""
    feature_size = 64
    _input_shapes = [[1, 3, 128, 128]]

    def __init__(self):
        self.feats = None
        self.calls = []

    def __call__(self, *, base_images, target_features):
        f = self.feats
        t = np.asarray(target_features, dtype=np.float32).reshape(-1, self.feature_size)
        sims = f @ t.T if len(t) else np.zeros((len(f), 0), np.float32)     # [N dets, M pool]
        self.calls.append((len(base_images), len(t)))
        return f.copy(), sims.astype(np.float32)


def test_botsort_update_feeds_the_face_similarity_term(api):
    """BoTSORT.update with a face encoder -> face crops / zero images -> encoder against the pool's
    face_curr_feature -> face_sim[n_pool, m] into the frame step (demo:1465-1486, demo:1541-1546).  The oracle gets
    the matrix from an independent mirror of the same bookkeeping."""
    det, enc, fenc = _Det(api), _Enc(), _FaceEnc()
    trk = api.BoTSORT(det, enc, fenc, frame_rate=30, max_tracks=256, max_dets=256)
    oracle = O.OracleBoTSORT()
    scene = SyntheticScene(SceneConfig(n_ids=40, feat_dim=2048, seed=9, low_frac=0.1, drop_frac=0.15, pitch_x=40.0, pitch_y=70.0,
                                       feat_noise=0.03))
    rng = np.random.default_rng(3)
    ident_face = rng.standard_normal((400, 64)).astype(np.float32)
    image = np.zeros((32, 32, 3), np.uint8)
    face_curr = {}
    try:
        for k in range(12):
            fr = scene.next_frame()
            det.boxes, det.scores, enc.feats = fr["boxes"], fr["scores"], fr["feats"]
            m = len(fr["boxes"])
            # a (noisy, unnormalised) face feature per detection; the body's identity decides it
            ff = (ident_face[fr["gt"]] + 0.05 * rng.standard_normal((m, 64))).astype(np.float32) if "gt" in fr \
                else rng.standard_normal((m, 64)).astype(np.float32)
            fenc.feats = ff
            # --- oracle side: pool order, similarity against the mirror's current face features ---
            pool = [t for t in oracle.tracked if t.is_activated] + list(oracle.lost)
            if pool and m:
                tf = np.stack([face_curr[t.track_id] for t in pool])
                sim = (ff @ tf.T).astype(np.float32).T.copy()
                sim[np.isclose(sim, 0.9999999, atol=1e-08, rtol=1e-08)] = 0.0
            else:
                sim = None
            out = trk.update(image)
            oracle.update_arrays(fr["boxes"], fr["scores"], fr["feats"], face_sims=sim)
            snap = oracle.snapshot()
            for i, tid in enumerate(snap["tracked"]["ids"]):
                if snap["tracked"]["frame_id"][i] == k + 1 and snap["tracked"]["det_index"][i] >= 0:
                    f = ff[snap["tracked"]["det_index"][i]].copy()
                    face_curr[int(tid)] = f / np.linalg.norm(f)
            assert [t.track_id for t in out] == list(snap["tracked"]["ids"]), f"frame {k + 1}"
            assert [t.track_id for t in trk.lost_stracks] == list(snap["lost"]["ids"])
            if m:
                assert fenc.calls[-1] == (m, len(pool))
        assert len(out) >= 30
    finally:
        trk.close()
