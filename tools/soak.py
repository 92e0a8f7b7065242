"""Randomised soak of the frame step against the CPU oracle: many seeded scenes with random density,
drop-out, low-score and newcomer rates, with and without ReID features; every frame's ids / states /
matches must be exact and the Kalman state within 1e-4 (the comparison of tests/test_gpu_tracker.py).
Margin-degenerate frames (a cost within 1e-3 of a threshold makes the decision implementation-defined,
SURVEY 8(d)) are reported separately.  Usage: python tools/soak.py [n_seeds] [first_seed]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import botsort_b200 as bs
from botsort_b200.synthetic import SceneConfig, SyntheticScene
from oracle import oracle_np as O
from test_gpu_tracker import _compare_frame

n_seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 30
first = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
BIG = bool(os.environ.get("SOAK_BIG"))     # SOAK_BIG=1: 300-2200 identities, 2048-d features, a few frames each
D = int(os.environ.get("SOAK_D", "2048" if BIG else "256"))     # SOAK_D=512: small scenes through the tensor-core path (narrow tiles)
ctx = bs.Context(max_tracks=2304 if BIG else 1024, max_dets=2304 if BIG else 1024, feat_dim=D,
                 flags=1 if os.environ.get("SOAK_SIMT") else 0)   # SOAK_SIMT=1: fp32 CUDA-core similarity (BT_FLAG_SIMT_SIM)
bad = []
t0 = time.time()
frames_total = 0
for seed in range(first, first + n_seeds):
    rng = np.random.default_rng(seed)
    with_reid = bool(rng.random() < 0.7)
    n_ids = int(rng.integers(300, 2200)) if BIG else int(rng.integers(8, 400))
    pitch = float(rng.uniform(22, 80))
    sc = SceneConfig(n_ids=n_ids, feat_dim=D, seed=seed, pitch_x=pitch, pitch_y=pitch * 1.7,
                     low_frac=float(rng.uniform(0, 0.3)), drop_frac=float(rng.uniform(0, 0.25)),
                     mid_frac=float(rng.uniform(0, 0.1)), walk=float(rng.uniform(1, 8)),
                     newcomer_every=int(rng.integers(2, 9)), with_features=with_reid)
    cfg = ctx.default_config()
    cfg.with_reid = 1 if with_reid else 0
    ctx.tracker_reset(cfg)
    scene = SyntheticScene(sc)
    oracle = O.OracleBoTSORT(mode="vectorized", lap_solver="jv", use_features=with_reid)
    frames = int(rng.integers(4, 9)) if BIG else int(rng.integers(15, 45))
    status = "ok"
    for k in range(frames):
        fr = scene.next_frame()
        feats = fr["feats"] if with_reid else None
        oracle.update_arrays(fr["boxes"], fr["scores"], feats)
        ctx.update_arrays(fr["boxes"], fr["scores"], feats)
        frames_total += 1
        try:
            _compare_frame(ctx, oracle, k + 1)
        except AssertionError as e:
            status = f"MISMATCH at frame {k + 1}: {str(e).splitlines()[0][:120]}"
            bad.append((seed, status))
            break
    print(f"seed {seed}: n_ids {n_ids:3d} pitch {pitch:4.1f} reid {int(with_reid)} frames {frames:2d} -> {status}", flush=True)
print(f"{n_seeds} scenes, {frames_total} frames, {len(bad)} mismatching scenes, {time.time() - t0:.1f} s")
for b in bad:
    print("  ", b)
