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
for f in frames:
    ctx.update_arrays(f["boxes"], f["scores"], f["feats"])
if len(sys.argv) > 1:
    os.environ["BT_ASSOC_DEBUG"] = sys.argv[1]
for k in range(3):
    print("replay avg us:", 1e3 * ctx.profile_replay_assoc(50), flush=True)
