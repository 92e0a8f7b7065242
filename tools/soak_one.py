"""Re-run one soak seed and print the first mismatching frame in full (debug aid for tools/soak.py)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import botsort_b200 as bs
from botsort_b200.synthetic import SceneConfig, SyntheticScene
from oracle import oracle_np as O
from test_gpu_tracker import _compare_frame

seed = int(sys.argv[1])
ctx = bs.Context(max_tracks=1024, max_dets=1024, feat_dim=256,
                 flags=1 if os.environ.get("SOAK_SIMT") else 0)   # SOAK_SIMT=1: fp32 CUDA-core similarity (BT_FLAG_SIMT_SIM)
rng = np.random.default_rng(seed)
with_reid = bool(rng.random() < 0.7)
n_ids = int(rng.integers(8, 400))
pitch = float(rng.uniform(22, 80))
sc = SceneConfig(n_ids=n_ids, feat_dim=256, seed=seed, pitch_x=pitch, pitch_y=pitch * 1.7,
                 low_frac=float(rng.uniform(0, 0.3)), drop_frac=float(rng.uniform(0, 0.25)),
                 mid_frac=float(rng.uniform(0, 0.1)), walk=float(rng.uniform(1, 8)),
                 newcomer_every=int(rng.integers(2, 9)), with_features=with_reid)
cfg = ctx.default_config(); cfg.with_reid = 1 if with_reid else 0
ctx.tracker_reset(cfg)
scene = SyntheticScene(sc)
oracle = O.OracleBoTSORT(mode="vectorized", lap_solver="jv", use_features=with_reid)
frames = int(rng.integers(15, 45))
np.set_printoptions(linewidth=200, precision=6, suppress=True)
for k in range(frames):
    fr = scene.next_frame()
    feats = fr["feats"] if with_reid else None
    oracle.update_arrays(fr["boxes"], fr["scores"], feats)
    ctx.update_arrays(fr["boxes"], fr["scores"], feats)
    try:
        _compare_frame(ctx, oracle, k + 1)
    except AssertionError as e:
        print("frame", k + 1, "mismatch:\n", str(e)[:1500])
        for stage in (1, 2, 3):
            g = ctx.get_matches(stage); r = oracle.last[f"matches{stage}"].astype(np.int32)
            if g.shape != r.shape or not np.array_equal(g, r):
                gs = set(map(tuple, g.tolist())); rs = set(map(tuple, r.tolist()))
                print(f"stage {stage}: only GPU {sorted(gs - rs)}  only oracle {sorted(rs - gs)}")
                d = oracle.last.get(f"dists{stage}")
                if d is not None:
                    for (a, b) in sorted((gs - rs) | (rs - gs)):
                        print(f"   cost[{a},{b}] = {d[a, b]!r}")
                    rows = sorted({a for a, _ in (gs ^ rs)}); cols = sorted({b for _, b in (gs ^ rs)})
                    print("   sub-matrix rows", rows, "cols", cols); print(d[np.ix_(rows, cols)])
        break
else:
    print("no mismatch")
