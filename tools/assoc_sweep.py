"""Tile-width sweep of the fused association kernel on the GPU box: for BN = 224 / 256 run the C3 tracker for a
few frames, replay the last frame's association launch 50 times back to back (bench.py's roofline measurement)
and check that every variant returns the same track ids.  python tools/assoc_sweep.py [n]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import botsort_b200 as bs  # noqa: E402
from botsort_b200.synthetic import SceneConfig, SyntheticScene  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
scene = SyntheticScene(SceneConfig(n_ids=n, feat_dim=2048, seed=1))
frames = [scene.next_frame() for _ in range(8)]
ctx = bs.Context(max_tracks=n + 256, max_dets=n + 256, feat_dim=2048)
ref_ids = None
for bn in ("224", "256"):
    os.environ["BT_ASSOC_BN"] = bn
    ctx.tracker_reset()
    for f in frames:
        ctx.update_arrays(f["boxes"], f["scores"], f["feats"].astype(np.float16))
    ids = ctx.get_tracks(0)["ids"]
    ref_ids = ids if ref_ids is None else ref_ids
    us = 1e3 * min(ctx.profile_replay_assoc(50) for _ in range(3))
    print(f"BN={bn}: {us:7.2f} us per launch  {2.0 * n * n * 2048 / (us * 1e-6) / 1e12:7.1f} TFLOP/s  ids_match={np.array_equal(ids, ref_ids)}", flush=True)
