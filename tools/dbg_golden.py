import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import botsort_b200 as bs
from botsort_b200.synthetic import SceneConfig, SyntheticScene
from oracle import oracle_np as O
from test_gpu_tracker import _compare_frame
from test_gpu_soak import _scene_for
seed = int(sys.argv[1]) if len(sys.argv) > 1 else 1001
flags = int(sys.argv[2]) if len(sys.argv) > 2 else 0
ctx = bs.Context(max_tracks=1024, max_dets=1024, feat_dim=256, flags=flags)
sc, with_reid, frames = _scene_for(seed)
cfg = ctx.default_config(); cfg.with_reid = 1 if with_reid else 0
ctx.tracker_reset(cfg)
scene = SyntheticScene(sc)
oracle = O.OracleBoTSORT(mode="vectorized", lap_solver="jv", use_features=with_reid)
print("seed", seed, "reid", with_reid, "n_ids", sc.n_ids, "frames", frames)
for k in range(frames):
    fr = scene.next_frame()
    feats = fr["feats"] if with_reid else None
    oracle.update_arrays(fr["boxes"], fr["scores"], feats)
    info = ctx.update_arrays(fr["boxes"], fr["scores"], feats)
    try:
        _compare_frame(ctx, oracle, k + 1)
    except AssertionError as e:
        print("frame", k + 1, info)
        print(str(e)[:1500])
        for stage in (1, 2, 3):
            g = ctx.get_matches(stage); r = oracle.last[f"matches{stage}"].astype(np.int32)
            gs = set(map(tuple, g.tolist())); rs = set(map(tuple, r.tolist()))
            print(f"stage {stage}: gpu {len(gs)} ref {len(rs)} only GPU {sorted(gs - rs)[:10]} only oracle {sorted(rs - gs)[:10]}")
            d = oracle.last.get(f"dists{stage}")
            if d is not None:
                for (a, b) in sorted((gs - rs) | (rs - gs))[:10]:
                    print(f"   cost[{a},{b}] = {d[a, b]!r}")
        break
else:
    print("no mismatch")
