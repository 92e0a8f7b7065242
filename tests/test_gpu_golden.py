"""CUDA path against the committed golden fixtures (outputs of the REAL reference, generated in
the build container by oracle/gen_golden.py).  ids / states / list order exact; floats 1e-4."""
import glob
import os

import numpy as np
import pytest

import botsort_b200 as bs

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TRACKER_FILES = sorted(glob.glob(os.path.join(GOLDEN, "tracker_*.npz")))
EXACT = ("ids", "state", "activated", "frame_id", "start_frame", "tracklet_len")
TOL = 1e-4


@pytest.mark.parametrize("path", TRACKER_FILES, ids=[os.path.basename(p) for p in TRACKER_FILES])
@pytest.mark.parametrize("flags", [0, bs.BT_FLAG_SIMT_SIM], ids=["tcgen05", "simt"])
def test_tracker_sequences_match_reference(path, flags):
    g = np.load(path)
    c = bs.Context(max_tracks=256, max_dets=256, feat_dim=int(g["feat_dim"]), flags=flags)
    try:
        c.tracker_reset()
        for k in range(int(g["frames"])):
            c.update_arrays(g[f"f{k}_boxes"], g[f"f{k}_scores"], g[f"f{k}_feats"])
            for which, lst in ((0, "tracked"), (1, "lost")):
                got = c.get_tracks(which, with_state=True)
                for key in EXACT:
                    np.testing.assert_array_equal(got[key], g[f"f{k}_{lst}_{key}"].astype(np.int32),
                                                  err_msg=f"frame {k + 1} {lst}.{key}")
                if len(got["ids"]):
                    np.testing.assert_allclose(got["score"], g[f"f{k}_{lst}_score"], atol=1e-7)
                    for key, name in (("mean", "mean"), ("cov", "cov"), ("tlbr", "tlbr")):
                        err = np.max(np.abs(got[key] - g[f"f{k}_{lst}_{name}"]))
                        assert err <= TOL, f"frame {k + 1} {lst}.{key} err {err}"
    finally:
        c.close()


def test_kalman_primitives(ctx):
    g = np.load(os.path.join(GOLDEN, "primitives.npz"))
    mean, cov = ctx.kalman_initiate(g["kf_z0"])
    np.testing.assert_array_equal(mean, g["kf_init_mean"].astype(np.float64))
    np.testing.assert_array_equal(cov, g["kf_init_cov"].astype(np.float64))
    for step in range(3):
        mean, cov = ctx.kalman_multi_predict(mean, cov, None, noise_f32=(step == 0))
        assert np.max(np.abs(mean - g[f"kf_pred{step}_mean"])) <= 1e-9
        assert np.max(np.abs(cov - g[f"kf_pred{step}_cov"])) <= 1e-9
        mean, cov = ctx.kalman_update(mean, cov, g[f"kf_z{step + 1}"].astype(np.float64))
        assert np.max(np.abs(mean - g[f"kf_upd{step}_mean"])) <= 1e-8
        assert np.max(np.abs(cov - g[f"kf_upd{step}_cov"])) <= 1e-8


def test_iou_and_assignment_primitives(ctx):
    g = np.load(os.path.join(GOLDEN, "primitives.npz"))
    d = ctx.iou_distance(g["iou_a"], g["iou_b"])
    assert np.max(np.abs(d - g["iou_dist"])) <= 1e-12
    for thresh in (0.8, 0.5, 0.7):
        tag = str(thresh).replace(".", "")
        x, y = ctx.lapjv(g["iou_dist"], thresh)
        matches = np.array([[i, j] for i, j in enumerate(x) if j >= 0]).reshape(-1, 2)
        np.testing.assert_array_equal(matches, g[f"lap_{tag}_matches"])
        np.testing.assert_array_equal(np.where(x < 0)[0], g[f"lap_{tag}_ua"])
        np.testing.assert_array_equal(np.where(y < 0)[0], g[f"lap_{tag}_ub"])


def test_crop_gather_primitive(ctx):
    g = np.load(os.path.join(GOLDEN, "primitives.npz"))
    got = ctx.reid_crop_gather(g["crop_frame"], g["crop_boxes"])
    np.testing.assert_array_equal(got, g["crop_out"])
