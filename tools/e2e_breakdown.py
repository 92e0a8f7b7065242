"""Where an end-to-end frame (host buffers in, tracks out) spends its wall time: update_arrays with pinned host
inputs vs device-resident inputs, and the get_tracks read-back (profiling aid for DESIGN section 6)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import botsort_b200 as bs
from botsort_b200._lib import BT_DEVICE, BT_HOST
from botsort_b200.synthetic import SceneConfig, SyntheticScene

n = 2000
scene = SyntheticScene(SceneConfig(n_ids=n, feat_dim=2048, seed=1))
frames = [scene.next_frame() for _ in range(24)]
ctx = bs.Context(max_tracks=n + 256, max_dets=n + 256, feat_dim=2048)
pinned = [{k: torch.from_numpy(np.ascontiguousarray(f[k])).pin_memory() for k in ("boxes", "scores", "feats")} for f in frames]
dev = [{k: v.cuda() for k, v in p.items()} for p in pinned]
for mode in ("host", "device"):
    ctx.tracker_reset()
    t_upd, t_get = [], []
    for i, (p, d) in enumerate(zip(pinned, dev)):
        src = p if mode == "host" else d
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ctx.update_arrays_raw(src["boxes"].data_ptr(), src["scores"].data_ptr(), src["feats"].data_ptr(), n,
                              BT_HOST if mode == "host" else BT_DEVICE)
        t1 = time.perf_counter()
        tr = ctx.get_tracks(0)
        t2 = time.perf_counter()
        if i >= 4:
            t_upd.append(t1 - t0); t_get.append(t2 - t1)
    print(f"{mode:6s} inputs: update_arrays {1e6 * np.mean(t_upd):7.1f} us   get_tracks {1e6 * np.mean(t_get):6.1f} us   ({len(tr['ids'])} tracks)")
# raw H2D of the same 16.4 MB for comparison
x = torch.empty_like(dev[0]["feats"])
for _ in range(3):
    x.copy_(pinned[0]["feats"], non_blocking=True)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    x.copy_(pinned[0]["feats"], non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 20
print(f"plain pinned H2D of the {x.numel() * 4 / 1e6:.1f} MB feature block: {1e6 * dt:.1f} us ({x.numel() * 4 / dt / 1e9:.1f} GB/s)")
