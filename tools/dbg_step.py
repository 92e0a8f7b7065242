import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import botsort_b200 as bs
from botsort_b200.synthetic import SceneConfig, SyntheticScene
from oracle import oracle_np as O
from test_gpu_tracker import _compare_frame
ctx = bs.Context(max_tracks=256, max_dets=256, feat_dim=2048)
scene = SyntheticScene(SceneConfig(n_ids=64, feat_dim=2048, seed=3, low_frac=0.15, drop_frac=0.1, mid_frac=0.05, newcomer_every=3))
oracle = O.OracleBoTSORT()
for k in range(12):
    fr = scene.next_frame()
    oracle.update_arrays(fr["boxes"], fr["scores"], fr["feats"])
    try:
        info = ctx.update_arrays(fr["boxes"], fr["scores"], fr["feats"])
        print("frame", k + 1, info, flush=True)
        _compare_frame(ctx, oracle, k + 1)
    except Exception as e:
        print("frame", k + 1, "FAILED:", str(e)[:600], flush=True)
        break
