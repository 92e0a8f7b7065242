"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz by running the REAL reference
(/root/reference/demo_bottrack_onnx_tflite.py, imported in place through oracle/ref_loader.py)
on seeded synthetic inputs.  Run in the build container only:

    python -m oracle.gen_golden

The fixtures travel to the GPU box (where /root/reference does not exist) and pin both the
oracle (tests/test_oracle_golden.py, CPU) and the CUDA path (tests/test_gpu_golden.py).
Library versions used are recorded in each file.
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle.ref_driver import ReferenceRunner  # noqa: E402
from oracle.ref_loader import load_reference  # noqa: E402
from botsort_b200.synthetic import SceneConfig, SyntheticScene  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

TRACKER_CASES = {
    # name: (scene config, frames)
    "tracker_c1_small": (SceneConfig(n_ids=24, feat_dim=64, seed=21, low_frac=0.2, drop_frac=0.15, mid_frac=0.1,
                                     newcomer_every=2), 30),
    "tracker_crowded": (SceneConfig(n_ids=40, feat_dim=64, seed=22, pitch_x=30.0, pitch_y=50.0, low_frac=0.2,
                                    drop_frac=0.2, walk=5.0), 30),
    "tracker_c1_64": (SceneConfig(n_ids=64, feat_dim=128, seed=23, low_frac=0.1, drop_frac=0.05), 12),
}


def versions():
    import cv2
    import scipy
    return np.array([f"numpy {np.__version__}", f"scipy {scipy.__version__}", f"cv2 {cv2.__version__}",
                     "lap shim: scipy LSA on lap's extended matrix"])


def gen_tracker(name, cfg, frames):
    ref = ReferenceRunner(cfg.feat_dim)
    scene = SyntheticScene(cfg)
    out = {"versions": versions(), "feat_dim": cfg.feat_dim, "frames": frames}
    for k in range(frames):
        fr = scene.next_frame()
        out[f"f{k}_boxes"] = fr["boxes"]
        out[f"f{k}_scores"] = fr["scores"]
        out[f"f{k}_feats"] = fr["feats"]
        ref.update_arrays(fr["boxes"], fr["scores"], fr["feats"].copy())
        snap = ref.snapshot()
        for lst in ("tracked", "lost"):
            for key, val in snap[lst].items():
                out[f"f{k}_{lst}_{key}"] = val
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "frames", frames, "tracked at end", len(snap["tracked"]["ids"]))


def gen_primitives():
    ref = load_reference()
    rng = np.random.default_rng(31)
    kf = ref.KalmanFilter()
    out = {"versions": versions()}
    # Kalman: initiate -> (multi_predict, update) x 3 on 16 tracks
    z0 = np.stack([rng.uniform(50, 3000, 16), rng.uniform(50, 2000, 16), rng.integers(20, 150, 16),
                   rng.integers(40, 250, 16)], axis=1).astype(np.float32)
    out["kf_z0"] = z0
    means, covs = zip(*[kf.initiate(z) for z in z0])
    mean = np.asarray(means); cov = np.asarray(covs)
    out["kf_init_mean"], out["kf_init_cov"] = mean, cov
    for step in range(3):
        mean, cov = kf.multi_predict(mean, cov)
        out[f"kf_pred{step}_mean"], out[f"kf_pred{step}_cov"] = mean, cov
        z = (mean[:, :4] + rng.normal(0, 2, (16, 4))).astype(np.float32)
        out[f"kf_z{step + 1}"] = z
        upd = [kf.update(mean[i], cov[i], z[i]) for i in range(16)]
        mean = np.asarray([u[0] for u in upd]); cov = np.asarray([u[1] for u in upd])
        out[f"kf_upd{step}_mean"], out[f"kf_upd{step}_cov"] = mean, cov
    # IoU distance + linear_assignment
    a = rng.uniform(0, 400, (40, 2)); a = np.hstack([a, a + rng.uniform(20, 150, (40, 2))])
    b = np.floor(rng.uniform(0, 400, (35, 2))); b = np.hstack([b, b + np.floor(rng.uniform(20, 150, (35, 2)))])
    b[0] = [a[0, 2], a[0, 1], a[0, 2] + 10, a[0, 3]]
    d = ref.iou_distance(list(a), list(b))
    out["iou_a"], out["iou_b"], out["iou_dist"] = a, b, d
    for thresh in (0.8, 0.5, 0.7):
        m, ua, ub = ref.linear_assignment(d, thresh)
        tag = str(thresh).replace(".", "")
        out[f"lap_{tag}_matches"] = np.asarray(m).reshape(-1, 2)
        out[f"lap_{tag}_ua"], out[f"lap_{tag}_ub"] = np.asarray(ua), np.asarray(ub)
    # crop + FastReID._preprocess through the reference's own method (bound to a bare object)
    frame = rng.integers(0, 256, (480, 640, 3), dtype=np.uint8)
    boxes = np.array([[10, 20, 138, 276], [300, 100, 364, 228], [0, 0, 640, 480], [600, 400, 640, 480],
                      [50, 60, 51, 300], [200, 50, 456, 562], [100, 100, 228, 356], [5, 5, 37, 69]], dtype=np.int32)

    class _Bare:
        _input_shapes = [[1, 3, 256, 128]]
        _h_index, _w_index = 2, 3
        _swap = (2, 0, 1)
        _mean = np.array([0.485, 0.456, 0.406], dtype=np.float32).reshape([1, 3, 1, 1])
        _std = np.array([0.229, 0.224, 0.225], dtype=np.float32).reshape([1, 3, 1, 1])
        _input_dtypes = [np.float32]

    crops = [frame[y1:y2, x1:x2, :] for x1, y1, x2, y2 in boxes]
    pre = ref.FastReID._preprocess(_Bare(), base_images=crops)
    out["crop_frame"], out["crop_boxes"], out["crop_out"] = frame, boxes, pre
    np.savez_compressed(os.path.join(OUT, "primitives.npz"), **out)
    print("primitives ok", pre.shape)


def gen_host_glue():
    """KalmanFilter.gating_distance (demo:338-380) and STrack.multi_gmc (demo:538-554): dead code in the reference's
    frame loop but part of its API -- pinned here so the host mirror cannot drift."""
    import types
    ref = load_reference()
    rng = np.random.default_rng(47)
    kf = ref.KalmanFilter()
    out = {"versions": versions()}
    z0 = np.stack([rng.uniform(50, 1500, 6), rng.uniform(50, 900, 6), rng.uniform(20, 150, 6), rng.uniform(40, 250, 6)], axis=1)
    means, covs = [], []
    for z in z0:
        m, c = kf.initiate(z)
        for _ in range(3):
            m, c = kf.predict(m, c)
        means.append(np.asarray(m, dtype=np.float64)); covs.append(np.asarray(c, dtype=np.float64))
    out["gd_mean"], out["gd_cov"] = np.asarray(means), np.asarray(covs)
    meas = np.stack([m[:4] + rng.normal(0, 6, (9, 4)) for m in means])            # [6, 9, 4]
    out["gd_meas"] = meas
    for metric in ("maha", "gaussian"):
        for only_pos in (False, True):
            out[f"gd_{metric}_{int(only_pos)}"] = np.asarray(
                [kf.gating_distance(means[i], covs[i], meas[i].copy(), only_position=only_pos, metric=metric) for i in range(6)])
    # multi_gmc on bare objects (the method only touches .mean / .covariance)
    ang = 0.03
    H = np.array([[np.cos(ang), -np.sin(ang), 4.5], [np.sin(ang), np.cos(ang), -2.25]])
    tracks = [types.SimpleNamespace(mean=means[i].copy(), covariance=covs[i].copy()) for i in range(6)]
    ref.STrack.multi_gmc(tracks, H)
    out["gmc_H"] = H
    out["gmc_mean"] = np.asarray([t.mean for t in tracks]); out["gmc_cov"] = np.asarray([t.covariance for t in tracks])
    np.savez_compressed(os.path.join(OUT, "host_glue.npz"), **out)
    print("host glue ok")


def main():
    os.makedirs(OUT, exist_ok=True)
    if len(sys.argv) > 1 and sys.argv[1] == "host_glue":
        return gen_host_glue()
    gen_primitives()
    gen_host_glue()
    for name, (cfg, frames) in TRACKER_CASES.items():
        gen_tracker(name, cfg, frames)


if __name__ == "__main__":
    main()
