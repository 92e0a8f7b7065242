"""Pins the CPU oracle (oracle/oracle_np.py) against the REAL reference executed in place
(/root/reference, build container only; skipped on the GPU box where it does not exist).
The same comparison over more frames generates tests/golden/*.npz (oracle/gen_golden.py),
which travel to the GPU box.
"""
import numpy as np
import pytest

from oracle import oracle_np as O
from oracle.ref_loader import reference_available
from botsort_b200.synthetic import SceneConfig, SyntheticScene

pytestmark = pytest.mark.skipif(not reference_available(), reason="/root/reference not present")

FIELDS_EXACT = ("ids", "state", "activated", "frame_id", "start_frame", "tracklet_len")


def _compare(ref_snap, ora_snap, frame):
    for name in ("tracked", "lost"):
        r, o = ref_snap[name], ora_snap[name]
        for f in FIELDS_EXACT:
            np.testing.assert_array_equal(r[f], o[f], err_msg=f"frame {frame} {name}.{f}")
        np.testing.assert_allclose(r["score"], o["score"], atol=0, err_msg=f"frame {frame} score")
        if len(r["ids"]):
            assert np.max(np.abs(r["mean"] - o["mean"])) <= 1e-9, f"frame {frame} {name}.mean"
            assert np.max(np.abs(r["cov"] - o["cov"])) <= 1e-9, f"frame {frame} {name}.cov"
            assert np.max(np.abs(r["tlbr"] - o["tlbr"])) <= 1e-9, f"frame {frame} {name}.tlbr"


@pytest.mark.parametrize("mode", ["vectorized", "faithful"])
@pytest.mark.parametrize("scene", [
    SceneConfig(n_ids=24, feat_dim=64, seed=1, low_frac=0.2, drop_frac=0.15, mid_frac=0.1, newcomer_every=2),
    SceneConfig(n_ids=40, feat_dim=128, seed=2, pitch_x=30.0, pitch_y=50.0, low_frac=0.2, drop_frac=0.2, walk=5.0),
])
def test_tracker_restatement_matches_reference(scene, mode):
    from oracle.ref_driver import ReferenceRunner
    ref = ReferenceRunner(scene.feat_dim)
    ora = O.OracleBoTSORT(mode=mode, lap_solver="jv")
    sc = SyntheticScene(scene)
    for k in range(25):
        fr = sc.next_frame()
        ref.update_arrays(fr["boxes"], fr["scores"], fr["feats"].copy())
        ora.update_arrays(fr["boxes"], fr["scores"], fr["feats"].copy())
        _compare(ref.snapshot(), ora.snapshot(), k + 1)


def test_kalman_restatement_matches_reference():
    from oracle.ref_loader import load_reference
    ref = load_reference()
    kf = ref.KalmanFilter()
    rng = np.random.default_rng(0)
    for _ in range(20):
        z = np.array([rng.uniform(0, 2000), rng.uniform(0, 2000), rng.integers(20, 100), rng.integers(40, 200)],
                     dtype=np.float32)
        m, c = kf.initiate(z)
        om, oc = O.kf_initiate(z)
        np.testing.assert_array_equal(m, om)
        np.testing.assert_array_equal(c, oc)
        m = m.astype(np.float64); c = c.astype(np.float64)
        for _ in range(4):
            rm, rc = kf.multi_predict(m[None].copy(), c[None].copy())
            qm, qc = O.kf_multi_predict(m[None].copy(), c[None].copy())
            np.testing.assert_array_equal(rm, qm)
            np.testing.assert_array_equal(rc, qc)
            zz = (rm[0][:4] + rng.normal(0, 2, 4)).astype(np.float32)
            m, c = kf.update(rm[0], rc[0], zz)
            om, oc = O.kf_update(qm[0], qc[0], zz)
            np.testing.assert_array_equal(m, om)
            np.testing.assert_array_equal(c, oc)
            bm, bc = O.kf_update_batch(qm, qc, zz[None])
            assert np.max(np.abs(bm[0] - m)) <= 1e-10 and np.max(np.abs(bc[0] - c)) <= 1e-10


def test_iou_and_assignment_match_reference():
    from oracle.ref_loader import load_reference
    ref = load_reference()
    rng = np.random.default_rng(1)
    a = rng.uniform(0, 300, (30, 2)); a = np.hstack([a, a + rng.uniform(20, 120, (30, 2))])
    b = np.floor(rng.uniform(0, 300, (25, 2))); b = np.hstack([b, b + np.floor(rng.uniform(20, 120, (25, 2)))])
    r = ref.iou_distance(list(a), list(b))
    np.testing.assert_array_equal(r, O.iou_distance(list(a), list(b), "faithful"))
    assert np.max(np.abs(r - O.iou_distance(a, b, "vectorized"))) <= 1e-15
    for thresh in (0.8, 0.5):
        rm, ru, rv = ref.linear_assignment(r, thresh)
        om, ou, ov = O.linear_assignment(r, thresh, "jv")
        np.testing.assert_array_equal(rm, om)
        np.testing.assert_array_equal(ru, ou)
        np.testing.assert_array_equal(rv, ov)
    # empty-matrix branch quirks (demo:1683-1684)
    rm, ru, rv = ref.linear_assignment(np.zeros((0, 3)), 0.8)
    om, ou, ov = O.linear_assignment(np.zeros((0, 3)), 0.8)
    assert rm.shape == om.shape == (0, 2) and ru == ou == () and rv == ov == (0, 1, 2)
