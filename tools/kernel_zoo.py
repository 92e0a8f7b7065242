"""Launch every stand-alone kernel once at BASELINE-config sizes (for the per-kernel ncu capture:
profiles/r01_kernels.csv).  Not a benchmark."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import botsort_b200 as bs
from botsort_b200.synthetic import SceneConfig, SyntheticScene
from oracle import detector_np as Dn

rng = np.random.default_rng(0)
n = 2000
ctx = bs.Context(max_tracks=2304, max_dets=2304, feat_dim=2048)
z = np.stack([rng.uniform(0, 4000, n), rng.uniform(0, 3000, n), rng.integers(20, 200, n), rng.integers(40, 300, n)], 1).astype(np.float32)
mean, cov = ctx.kalman_initiate(z)
for _ in range(2):
    mean, cov = ctx.kalman_multi_predict(mean, cov, np.ones(n, np.int32))
    mean, cov = ctx.kalman_update(mean, cov, mean[:, :4] + 1.0)
ctx.kalman_project(mean, cov)
a = np.hstack([z[:, :2], z[:, :2] + z[:, 2:]]).astype(np.float64)
ctx.iou_distance(a, np.floor(a))
f1 = rng.standard_normal((n, 2048)).astype(np.float32); f1 /= np.linalg.norm(f1, axis=1, keepdims=True)
f2 = rng.standard_normal((n, 2048)).astype(np.float32); f2 /= np.linalg.norm(f2, axis=1, keepdims=True)
ctx.embedding_distance(f1, f2, precision=0)
ctx.fused_cost(a, np.floor(a), f1, f2, stage=1)
cost = rng.uniform(0, 1, (512, 512)); cost[rng.uniform(size=cost.shape) > 0.02] = 1.0
ctx.lapjv(cost, 0.8)
ctx.feature_ema(f1, f1, f2)
raw = Dn.synth_yolox_head(rng, [(50, 60, 120, 260), (300, 100, 380, 300)], [0, 0], [0.95, 0.6])
ctx.yolox_postprocess(raw)
frame = rng.integers(0, 256, (480, 640, 3), dtype=np.uint8)
bx = np.stack([rng.integers(0, 500, 50), rng.integers(0, 300, 50)], 1); bx = np.hstack([bx, bx + rng.integers(30, 140, (50, 2))]).astype(np.int32)
ctx.reid_crop_gather(frame, bx)
# the tracker path (C3 steady state, a few frames)
scene = SyntheticScene(SceneConfig(n_ids=n, feat_dim=2048, seed=1))
ctx.tracker_reset()
for _ in range(4):
    f = scene.next_frame()
    ctx.update_arrays(f["boxes"], f["scores"], f["feats"])
print("zoo done", ctx.launch_count)
