"""The oracle against the committed golden fixtures (tests/golden/*.npz, produced by running the
real reference: oracle/gen_golden.py).  Runs on CPU everywhere, including the GPU box."""
import glob
import os

import numpy as np
import pytest

from oracle import detector_np as Dn
from oracle import oracle_np as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TRACKER_FILES = sorted(glob.glob(os.path.join(GOLDEN, "tracker_*.npz")))
EXACT = ("ids", "state", "activated", "frame_id", "start_frame", "tracklet_len")


def test_fixtures_present():
    assert len(TRACKER_FILES) >= 3 and os.path.exists(os.path.join(GOLDEN, "primitives.npz"))


@pytest.mark.parametrize("path", TRACKER_FILES, ids=[os.path.basename(p) for p in TRACKER_FILES])
@pytest.mark.parametrize("mode", ["vectorized", "faithful"])
def test_tracker_sequences(path, mode):
    g = np.load(path)
    trk = O.OracleBoTSORT(mode=mode, lap_solver="jv")
    for k in range(int(g["frames"])):
        trk.update_arrays(g[f"f{k}_boxes"], g[f"f{k}_scores"], g[f"f{k}_feats"].copy())
        snap = trk.snapshot()
        for lst in ("tracked", "lost"):
            for key in EXACT:
                np.testing.assert_array_equal(snap[lst][key], g[f"f{k}_{lst}_{key}"], err_msg=f"frame {k + 1} {lst}.{key}")
            if len(snap[lst]["ids"]):
                for key in ("mean", "cov", "tlbr"):
                    assert np.max(np.abs(snap[lst][key] - g[f"f{k}_{lst}_{key}"])) <= 1e-9, f"frame {k + 1} {lst}.{key}"


def test_kalman_primitives():
    g = np.load(os.path.join(GOLDEN, "primitives.npz"))
    ms, cs = zip(*[O.kf_initiate(z) for z in g["kf_z0"]])
    mean, cov = np.asarray(ms), np.asarray(cs)
    np.testing.assert_array_equal(mean, g["kf_init_mean"])
    np.testing.assert_array_equal(cov, g["kf_init_cov"])
    for step in range(3):
        mean, cov = O.kf_multi_predict(mean, cov)
        np.testing.assert_array_equal(mean, g[f"kf_pred{step}_mean"])
        np.testing.assert_array_equal(cov, g[f"kf_pred{step}_cov"])
        z = g[f"kf_z{step + 1}"]
        upd = [O.kf_update(mean[i], cov[i], z[i]) for i in range(len(z))]
        mean = np.asarray([u[0] for u in upd]); cov = np.asarray([u[1] for u in upd])
        np.testing.assert_array_equal(mean, g[f"kf_upd{step}_mean"])
        np.testing.assert_array_equal(cov, g[f"kf_upd{step}_cov"])


def test_iou_and_assignment_primitives():
    g = np.load(os.path.join(GOLDEN, "primitives.npz"))
    a, b = g["iou_a"], g["iou_b"]
    np.testing.assert_array_equal(O.iou_distance(list(a), list(b), "faithful"), g["iou_dist"])
    assert np.max(np.abs(O.iou_distance(a, b, "vectorized") - g["iou_dist"])) <= 1e-15
    for thresh in (0.8, 0.5, 0.7):
        tag = str(thresh).replace(".", "")
        m, ua, ub = O.linear_assignment(g["iou_dist"], thresh, "jv")
        np.testing.assert_array_equal(np.asarray(m).reshape(-1, 2), g[f"lap_{tag}_matches"])
        np.testing.assert_array_equal(ua, g[f"lap_{tag}_ua"])
        np.testing.assert_array_equal(ub, g[f"lap_{tag}_ub"])


def test_crop_preprocess_primitive():
    g = np.load(os.path.join(GOLDEN, "primitives.npz"))
    got = Dn.crop_preprocess(g["crop_frame"], g["crop_boxes"])
    np.testing.assert_array_equal(got, g["crop_out"])
