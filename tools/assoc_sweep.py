"""Profiling sweep of the fused association kernel on the GPU box: for each cluster shape run the
C3 tracker for a few frames with segment timing on and print the assoc kernel's average time.
Also checks that every variant returns the same track ids (correctness guard)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import botsort_b200 as bs  # noqa: E402
from botsort_b200.synthetic import SceneConfig, SyntheticScene  # noqa: E402

n = int(os.environ.get("SWEEP_N", "2000"))
frames_n = 14
scene = SyntheticScene(SceneConfig(n_ids=n, feat_dim=2048, seed=1))
frames = [scene.next_frame() for _ in range(frames_n)]
ctx = bs.Context(max_tracks=n + 256, max_dets=n + 256, feat_dim=2048)
ref_ids = None
for variant in sys.argv[1:] or ["1x1", "1x2", "2x1", "2x2", "4x2", "2x4", "4x1", "1x4"]:
    os.environ["BT_ASSOC_CLUSTER"] = variant
    try:
        ctx.tracker_reset()
        for f in frames[:4]:
            ctx.update_arrays(f["boxes"], f["scores"], f["feats"])
        ctx.profile_enable(True)
        for f in frames[4:]:
            ctx.update_arrays(f["boxes"], f["scores"], f["feats"])
        prof = ctx.profile_read()
        ctx.profile_enable(False)
        ids = ctx.get_tracks(0)["ids"]
        if ref_ids is None:
            ref_ids = ids
        ok = np.array_equal(ids, ref_ids)
        ms, cnt = prof["assoc"]
        tf = 2.0 * n * n * 2048 / (ms / cnt * 1e-3) / 1e12
        print(f"{variant}: assoc {1e3 * ms / cnt:8.2f} us  {tf:7.1f} TFLOP/s  ids_match={ok}  "
              + " ".join(f"{k}={1e3 * v[0] / max(1, v[1]):.1f}us" for k, v in prof.items()), flush=True)
    except Exception as e:  # noqa: BLE001
        print(f"{variant}: FAILED {e}", flush=True)
