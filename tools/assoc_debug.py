"""One C3 frame with the association / LAP kernels' device-side phase stamps switched on for the LAST frame:
    python tools/assoc_debug.py [BT_ASSOC_DEBUG bits, default 1]
(bits: csrc/reid_gemm.cu EpiParams::debug; BT_LAP_DEBUG=1 prints the LAP stage timeline of the same frame)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import botsort_b200 as bs
from botsort_b200.synthetic import SceneConfig, SyntheticScene
n = 2000
scene = SyntheticScene(SceneConfig(n_ids=n, feat_dim=2048, seed=1))
frames = [scene.next_frame() for _ in range(5)]
ctx = bs.Context(max_tracks=n + 256, max_dets=n + 256, feat_dim=2048)
ctx.tracker_reset()
for i, f in enumerate(frames):
    if i == 4:
        os.environ["BT_ASSOC_DEBUG"] = sys.argv[1] if len(sys.argv) > 1 else "1"
        os.environ["BT_LAP_DEBUG"] = "1"
    ctx.update_arrays(f["boxes"], f["scores"], f["feats"].astype(np.float16))
ctx.sync()
