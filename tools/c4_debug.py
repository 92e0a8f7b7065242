"""Debug aid: C4-like scene (12 identities, 256-d stub features with high cross similarity); prints the stage-1
cost matrix of frame 2 (dense kernel), the dense-LAP answer and the tracker's own stage-1 matches."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import botsort_b200 as bs
from oracle import detector_np as Dn

sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_gpu_c4_pipeline import _scene, _frame_and_head, D

np.set_printoptions(linewidth=250, precision=3, suppress=True)
ctx = bs.Context(max_tracks=512, max_dets=512, feat_dim=2048)
rng = np.random.default_rng(0)
proj = np.random.default_rng(7).standard_normal((3 * 256 * 128, D)).astype(np.float32) / 300.0
k = 12
base = _scene(rng, k)
trk = bs.Context(max_tracks=256, max_dets=256, feat_dim=D)
trk.tracker_reset()
for f in range(3):
    boxes = base + rng.uniform(-3, 3, base.shape)
    scores = np.full(k, 0.96)
    frame, raw = _frame_and_head(rng, boxes, scores)
    det = ctx.yolox_postprocess(raw)
    body = det[det[:, 0] == 0]
    b_int = body[:, 2:6].astype(np.int32)
    sc = body[:, 1].astype(np.float32)
    crops = ctx.reid_crop_gather(frame, b_int)
    feats = crops.reshape(len(b_int), -1) @ proj
    feats /= np.linalg.norm(feats, axis=1, keepdims=True)
    feats = feats.astype(np.float32)
    if f >= 1:
        pre = trk.get_tracks(0, with_state=True)
        curr, _ = trk.get_track_features(0)
        # predicted boxes are not exposed before the update; frame-1 boxes are close enough for a look
        cost = ctx.fused_cost(pre["tlbr"], b_int.astype(np.float64), curr, feats, stage=1)
        print(f"frame {f+1}: stage-1 cost (tracks x dets), entries < 0.8 per row:", (cost < 0.8).sum(1))
        print(cost)
        x, y = ctx.lapjv(cost, 0.8)
        print("dense LAP x:", x.tolist())
    trk.update_arrays(b_int, sc, feats)
    print(f"frame {f+1}: tracker stage-1 matches:", trk.get_matches(1).tolist(), "ids", trk.get_tracks(0)["ids"].tolist())
