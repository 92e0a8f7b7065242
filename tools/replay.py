import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import botsort_b200 as bs
from botsort_b200.synthetic import SceneConfig, SyntheticScene
n = 2000
scene = SyntheticScene(SceneConfig(n_ids=n, feat_dim=2048, seed=1))
frames = [scene.next_frame() for _ in range(4)]
ctx = bs.Context(max_tracks=n + 256, max_dets=n + 256, feat_dim=2048)
ctx.tracker_reset()
import torch
from botsort_b200._lib import BT_DEVICE
import numpy as np
from botsort_b200._lib import BT_F16
for f in frames:
    ctx.update_arrays(f["boxes"], f["scores"], f["feats"].astype(np.float16))
if os.environ.get("BT_HOST_DEBUG"):
    fl = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for k in range(6):
        f = scene.next_frame()
        b_, s_ = (torch.from_numpy(f[k2]).cuda() for k2 in ("boxes", "scores"))
        f_ = torch.from_numpy(f["feats"].astype(np.float16)).cuda()
        fl.zero_(); torch.cuda.synchronize()
        ctx.update_arrays_raw(b_.data_ptr(), s_.data_ptr(), f_.data_ptr(), b_.shape[0], BT_DEVICE, None, BT_F16)
if len(sys.argv) > 1:
    os.environ["BT_ASSOC_DEBUG"] = sys.argv[1]
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 50
for k in range(3):
    print("replay avg us:", 1e3 * ctx.profile_replay_assoc(iters), flush=True)
