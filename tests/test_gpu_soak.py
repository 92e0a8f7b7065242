"""Randomised scenes against the CPU oracle (a small slice of tools/soak.py, which ran 500 scenes / 14.5 k
frames with one mismatch: a 2.5e-6 near-tie between two tracks' ReID costs for one detection, inside the fp16
similarity's 3e-5 error -- exact again with the fp32 similarity flag).  Every frame: ids, states, list
order and matches exact, Kalman state 1e-4."""
import numpy as np
import pytest

import botsort_b200 as bs
from botsort_b200.synthetic import SceneConfig, SyntheticScene
from oracle import oracle_np as O
from test_gpu_tracker import _compare_frame

pytestmark = pytest.mark.gpu


def _scene_for(seed):
    rng = np.random.default_rng(seed)
    with_reid = bool(rng.random() < 0.7)
    n_ids = int(rng.integers(8, 400))
    pitch = float(rng.uniform(22, 80))
    sc = SceneConfig(n_ids=n_ids, feat_dim=256, seed=seed, pitch_x=pitch, pitch_y=pitch * 1.7,
                     low_frac=float(rng.uniform(0, 0.3)), drop_frac=float(rng.uniform(0, 0.25)),
                     mid_frac=float(rng.uniform(0, 0.1)), walk=float(rng.uniform(1, 8)),
                     newcomer_every=int(rng.integers(2, 9)), with_features=with_reid)
    return sc, with_reid, int(rng.integers(15, 45))


@pytest.mark.parametrize("flags", [0, bs._lib.BT_FLAG_SIMT_SIM])
def test_random_scenes_match_the_oracle(flags):
    ctx = bs.Context(max_tracks=1024, max_dets=1024, feat_dim=256, flags=flags)
    try:
        for seed in (1001, 1009, 1020, 1029, 1034, 1036, 2226):
            sc, with_reid, frames = _scene_for(seed)
            cfg = ctx.default_config()
            cfg.with_reid = 1 if with_reid else 0
            ctx.tracker_reset(cfg)
            scene = SyntheticScene(sc)
            oracle = O.OracleBoTSORT(mode="vectorized", lap_solver="jv", use_features=with_reid)
            for k in range(frames):
                fr = scene.next_frame()
                feats = fr["feats"] if with_reid else None
                oracle.update_arrays(fr["boxes"], fr["scores"], feats)
                ctx.update_arrays(fr["boxes"], fr["scores"], feats)
                _compare_frame(ctx, oracle, k + 1)
    finally:
        ctx.close()


def test_wide_tiles_of_the_association_kernel(monkeypatch):
    """The tile width of the association GEMM is picked per launch (224 or 256 columns; 256 only wins for
    2017-2304 detections): force the 256-column variant of the candidate path through a few scenes."""
    monkeypatch.setenv("BT_ASSOC_BN", "256")
    ctx = bs.Context(max_tracks=1024, max_dets=1024, feat_dim=256)
    try:
        for seed in (1001, 1020, 1034):
            sc, with_reid, frames = _scene_for(seed)
            cfg = ctx.default_config()
            cfg.with_reid = 1 if with_reid else 0
            ctx.tracker_reset(cfg)
            scene = SyntheticScene(sc)
            oracle = O.OracleBoTSORT(mode="vectorized", lap_solver="jv", use_features=with_reid)
            for k in range(frames):
                fr = scene.next_frame()
                feats = fr["feats"] if with_reid else None
                oracle.update_arrays(fr["boxes"], fr["scores"], feats)
                ctx.update_arrays(fr["boxes"], fr["scores"], feats)
                _compare_frame(ctx, oracle, k + 1)
    finally:
        ctx.close()
