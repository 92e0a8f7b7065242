"""Host phase times of the pipelined end-to-end loop (bt_submit_streams(frame k+1) then bt_step_streams(frame k),
pinned fp16 inputs) next to the same steps on device-resident inputs: run with BT_HOST_DEBUG=1 and compare the
phases -- does the next frame's 8.2 MB H2D copy slow the current step's PCIe round trips?"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import botsort_b200 as bs
from botsort_b200._lib import BT_DEVICE, BT_HOST, BT_F16
from botsort_b200.synthetic import SceneConfig, SyntheticScene

n = 2000
scene = SyntheticScene(SceneConfig(n_ids=n, feat_dim=2048, seed=1))
frames = [scene.next_frame() for _ in range(16)]
ctx = bs.Context(max_tracks=n + 256, max_dets=n + 256, feat_dim=2048)
ctx.tracker_reset()
pin = [(torch.from_numpy(f["boxes"]).pin_memory(), torch.from_numpy(f["scores"]).pin_memory(),
        torch.from_numpy(f["feats"].astype(np.float16)).pin_memory()) for f in frames]
def P(i): b, s, f = pin[i]; return [b.data_ptr()], [s.data_ptr()], [f.data_ptr()], [n]
b, s, f, m = P(0)
ctx.update_streams_raw([0], b, s, f, m, BT_HOST, BT_F16)
b, s, f, m = P(1)
ctx.submit_streams_raw([0], b, s, f, m, BT_HOST, BT_F16)
ts = []
for k in range(1, 15):
    t0 = time.perf_counter()
    b, s, f, m = P(k + 1)
    ctx.submit_streams_raw([0], b, s, f, m, BT_HOST, BT_F16)
    t1 = time.perf_counter()
    ctx.step_streams_raw([0])
    t2 = time.perf_counter()
    ctx.get_tracks(0)
    t3 = time.perf_counter()
    ts.append((t1 - t0, t2 - t1, t3 - t2))
ctx.step_streams_raw([0])
a = np.array(ts[4:]) * 1e6
print("pipelined e2e loop, us per iteration: submit %.1f  step %.1f  get_tracks %.1f  total %.1f" % (*a.mean(0), a.sum(1).mean()), flush=True)
