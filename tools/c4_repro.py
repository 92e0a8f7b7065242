"""Debug aid: runs the C4 pipeline test body and prints the tracker's matches when an id goes missing."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import botsort_b200 as bs
from oracle import detector_np as Dn
from oracle import oracle_np as O
from test_gpu_c4_pipeline import _scene, _frame_and_head, D

ctx = bs.Context(max_tracks=2304, max_dets=2304, feat_dim=2048)
rng = np.random.default_rng(0)
proj = np.random.default_rng(7).standard_normal((3 * 256 * 128, D)).astype(np.float32) / 300.0
k = 12
base = _scene(rng, k)
trk = bs.Context(max_tracks=256, max_dets=256, feat_dim=D)
oracle = O.OracleBoTSORT()
trk.tracker_reset()
for f in range(8):
    boxes = base + rng.uniform(-3, 3, base.shape)
    scores = np.full(k, 0.96)
    if f > 2:
        scores[f % k] = 0.3
    frame, raw = _frame_and_head(rng, boxes, scores)
    det = ctx.yolox_postprocess(raw)
    det_o = Dn.yolox_postprocess(raw, img_h=480, img_w=640)
    body = det[det[:, 0] == 0]
    b_int = body[:, 2:6].astype(np.int32)
    sc = body[:, 1].astype(np.float32)
    crops = ctx.reid_crop_gather(frame, b_int)
    crops_o = Dn.crop_preprocess(frame, b_int)
    feats = crops.reshape(len(b_int), -1) @ proj
    feats /= np.linalg.norm(feats, axis=1, keepdims=True)
    feats = feats.astype(np.float32)
    trk.update_arrays(b_int, sc, feats)
    oracle.update_arrays(b_int, det_o[det_o[:, 0] == 0][:, 1].astype(np.float32), feats.copy())
    got = trk.get_tracks(0, with_state=True)
    ref = oracle.snapshot()["tracked"]
    ok = np.array_equal(got["ids"], ref["ids"].astype(np.int32))
    print(f"frame {f+1}: {'ok' if ok else 'MISMATCH'} ids {got['ids'].tolist()} stage1 {trk.get_matches(1).tolist()} "
          f"stage2 {trk.get_matches(2).tolist()} stage3 {trk.get_matches(3).tolist()}")
    if not ok:
        break
