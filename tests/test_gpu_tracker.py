"""GPU parity of the whole frame step (bt_update_arrays through ctypes) against the CPU oracle
restatement of BoTSORT.update on identical seeded synthetic streams.

ids / states / list order / matches: exact.  Kalman means, covariances, boxes: 1e-4.
"""
import numpy as np
import pytest

from oracle import oracle_np as O
from botsort_b200.synthetic import SceneConfig, SyntheticScene

pytestmark = pytest.mark.gpu

TOL = 1e-4
INT_FIELDS = ("ids", "state", "activated", "frame_id", "start_frame", "tracklet_len", "det_index")


def _compare_frame(ctx, oracle, frame_no, stream=0):
    snap = oracle.snapshot()
    for which, name in ((0, "tracked"), (1, "lost")):
        got = ctx.get_tracks(which, with_state=True, stream=stream)
        ref = snap[name]
        for f in INT_FIELDS:
            np.testing.assert_array_equal(got[f], ref[f].astype(np.int32), err_msg=f"frame {frame_no} {name}.{f}")
        np.testing.assert_allclose(got["score"], ref["score"], atol=1e-7, err_msg=f"frame {frame_no} {name}.score")
        if len(ref["ids"]):
            assert np.max(np.abs(got["tlbr"] - ref["tlbr"])) <= TOL, f"frame {frame_no} {name}.tlbr"
            assert np.max(np.abs(got["mean"] - ref["mean"])) <= TOL, f"frame {frame_no} {name}.mean"
            assert np.max(np.abs(got["cov"] - ref["cov"])) <= TOL, f"frame {frame_no} {name}.cov"
    for stage in (1, 2, 3):
        np.testing.assert_array_equal(ctx.get_matches(stage, stream=stream), oracle.last[f"matches{stage}"].astype(np.int32),
                                      err_msg=f"frame {frame_no} stream {stream} matches{stage}")


def _run(ctx, scene_cfg, frames, with_reid=True):
    cfg = ctx.default_config()
    cfg.with_reid = 1 if with_reid else 0
    ctx.tracker_reset(cfg)
    scene = SyntheticScene(scene_cfg)
    oracle = O.OracleBoTSORT(mode="vectorized", lap_solver="jv", use_features=with_reid)
    infos = []
    for k in range(frames):
        fr = scene.next_frame()
        feats = fr["feats"] if with_reid else None
        oracle.update_arrays(fr["boxes"], fr["scores"], feats)
        infos.append(ctx.update_arrays(fr["boxes"], fr["scores"], feats))
        _compare_frame(ctx, oracle, k + 1)
    return infos


def test_c1_64x64_reid(ctx):
    """BASELINE config 1: 64 tracks x 64 dets, 2048-d features, births/lost/refind/unconfirmed paths."""
    infos = _run(ctx, SceneConfig(n_ids=64, feat_dim=2048, seed=3, low_frac=0.15, drop_frac=0.1, mid_frac=0.05,
                                  newcomer_every=3), frames=40)
    assert sum(i["n_matches2"] for i in infos) > 0
    assert sum(i["n_matches3"] for i in infos) > 0
    assert any(i["n_lost"] > 0 for i in infos)


def test_c1_crowded(ctx):
    """Dense scene: boxes overlap their neighbours, the candidate graph has multi-row components."""
    _run(ctx, SceneConfig(n_ids=96, feat_dim=2048, seed=5, pitch_x=30.0, pitch_y=50.0, low_frac=0.2,
                          drop_frac=0.15, walk=5.0), frames=30)


def test_c2_512_iou_only(ctx):
    """BASELINE config 2: 512 x 512, IoU-only association (no ReID)."""
    infos = _run(ctx, SceneConfig(n_ids=512, feat_dim=2048, seed=7, low_frac=0.1, drop_frac=0.05,
                                  with_features=False, pitch_x=45.0, pitch_y=80.0), frames=12, with_reid=False)
    assert infos[-1]["n_pool"] >= 480


def test_c3_2000x2000(ctx):
    """BASELINE config 3: 2000 x 2000 x 2048-d."""
    infos = _run(ctx, SceneConfig(n_ids=2000, feat_dim=2048, seed=11, low_frac=0.05, drop_frac=0.02), frames=4)
    assert infos[-1]["n_pool"] >= 1900


def test_empty_and_ragged_frames(ctx):
    ctx.tracker_reset()
    oracle = O.OracleBoTSORT()
    scene = SyntheticScene(SceneConfig(n_ids=20, feat_dim=2048, seed=1))
    empty = (np.zeros((0, 4), np.int32), np.zeros(0, np.float32), np.zeros((0, 2048), np.float32))
    seq = []
    seq.append(empty)                                   # frame 1 without detections
    f = scene.next_frame(); seq.append((f["boxes"], f["scores"], f["feats"]))
    f = scene.next_frame(); seq.append((f["boxes"][:7], f["scores"][:7], f["feats"][:7]))
    seq.append(empty)                                   # everything goes lost
    f = scene.next_frame(); seq.append((f["boxes"], f["scores"], f["feats"]))
    f = scene.next_frame(); seq.append((f["boxes"], np.full(len(f["boxes"]), 0.3, np.float32), f["feats"]))
    for k, (b, s, ft) in enumerate(seq):
        oracle.update_arrays(b, s, ft)
        ctx.update_arrays(b, s, ft)
        _compare_frame(ctx, oracle, k + 1)


def test_track_features_state(ctx):
    """A10 exposed state: curr / smooth feature banks follow STrack.update_body_features."""
    ctx.tracker_reset()
    oracle = O.OracleBoTSORT()
    scene = SyntheticScene(SceneConfig(n_ids=32, feat_dim=2048, seed=9))
    for _ in range(5):
        f = scene.next_frame()
        oracle.update_arrays(f["boxes"], f["scores"], f["feats"])
        ctx.update_arrays(f["boxes"], f["scores"], f["feats"])
    curr, smooth = ctx.get_track_features(0)
    ref_curr = np.array([t.curr_feat for t in oracle.tracked])
    ref_smooth = np.array([t.smooth_feat for t in oracle.tracked])
    assert np.max(np.abs(curr - ref_curr)) <= 1e-6
    assert np.max(np.abs(smooth - ref_smooth)) <= 1e-6


def test_capacity_error(ctx):
    ctx.tracker_reset()
    with pytest.raises(ValueError):
        ctx.update_arrays(np.zeros((ctx.max_dets + 1, 4), np.int32), np.zeros(ctx.max_dets + 1, np.float32),
                          np.zeros((ctx.max_dets + 1, 2048), np.float32))


def test_device_resident_inputs_match_host_inputs(ctx):
    """loc = BT_DEVICE (inputs already in HBM, what bench.py's `value` leg times) gives the same tracks
    as host buffers."""
    import torch
    from botsort_b200._lib import BT_DEVICE
    scene_cfg = SceneConfig(n_ids=150, feat_dim=2048, seed=13, low_frac=0.1, drop_frac=0.1, newcomer_every=4)
    frames = [SyntheticScene(scene_cfg).next_frame()]
    sc = SyntheticScene(scene_cfg)
    frames = [sc.next_frame() for _ in range(8)]
    ctx.tracker_reset()
    ref = []
    for fr in frames:
        ctx.update_arrays(fr["boxes"], fr["scores"], fr["feats"])
        ref.append((ctx.get_tracks(0, with_state=True), ctx.get_tracks(1)))
    ctx.tracker_reset()
    for fr, (rt, rl) in zip(frames, ref):
        b = torch.from_numpy(fr["boxes"]).cuda()
        s = torch.from_numpy(fr["scores"]).cuda()
        f = torch.from_numpy(fr["feats"]).cuda()
        torch.cuda.synchronize()
        ctx.update_arrays_raw(b.data_ptr(), s.data_ptr(), f.data_ptr(), b.shape[0], BT_DEVICE)
        gt, gl = ctx.get_tracks(0, with_state=True), ctx.get_tracks(1)
        for key in INT_FIELDS:
            np.testing.assert_array_equal(gt[key], rt[key])
            np.testing.assert_array_equal(gl[key], rl[key])
        np.testing.assert_array_equal(gt["tlbr"], rt["tlbr"])
        np.testing.assert_array_equal(gt["mean"], rt["mean"])


def test_two_contexts_interleaved(ctx):
    """Independent video streams = independent bt_ctx objects (own CUDA stream, own id counter):
    interleaving two trackers frame by frame gives the same tracks as running each alone."""
    import botsort_b200 as bs
    cfgs = [SceneConfig(n_ids=60, feat_dim=2048, seed=21, low_frac=0.1, drop_frac=0.1),
            SceneConfig(n_ids=45, feat_dim=2048, seed=22, low_frac=0.2, drop_frac=0.05, newcomer_every=2)]
    seqs = []
    for c in cfgs:
        sc = SyntheticScene(c)
        seqs.append([sc.next_frame() for _ in range(10)])
    alone = []
    for seq in seqs:
        ctx.tracker_reset()
        out = []
        for fr in seq:
            ctx.update_arrays(fr["boxes"], fr["scores"], fr["feats"])
            out.append(ctx.get_tracks(0))
        alone.append(out)
    a = bs.Context(max_tracks=256, max_dets=256, feat_dim=2048)
    b = bs.Context(max_tracks=256, max_dets=256, feat_dim=2048)
    try:
        for k in range(10):
            for c, seq, ref in ((a, seqs[0], alone[0]), (b, seqs[1], alone[1])):
                fr = seq[k]
                c.update_arrays(fr["boxes"], fr["scores"], fr["feats"])
                got = c.get_tracks(0)
                np.testing.assert_array_equal(got["ids"], ref[k]["ids"])
                np.testing.assert_array_equal(got["tlbr"], ref[k]["tlbr"])
    finally:
        a.close()
        b.close()


def test_replay_hook_leaves_state_untouched(ctx):
    """bt_profile_replay_assoc (bench.py's roofline measurement) must not disturb the tracker."""
    scene = SyntheticScene(SceneConfig(n_ids=300, feat_dim=2048, seed=31))
    frames = [scene.next_frame() for _ in range(6)]
    outs = []
    for use_replay in (False, True):
        ctx.tracker_reset()
        res = []
        for fr in frames:
            ctx.update_arrays(fr["boxes"], fr["scores"], fr["feats"])
            if use_replay:
                assert ctx.profile_replay_assoc(3) > 0
            res.append(ctx.get_tracks(0, with_state=True))
        outs.append(res)
    for r0, r1 in zip(*outs):
        np.testing.assert_array_equal(r0["ids"], r1["ids"])
        np.testing.assert_array_equal(r0["mean"], r1["mean"])


def test_very_crowded_scene_large_components(ctx):
    """Boxes packed so tightly that the candidate graph is a few huge components (the LAP's global-memory
    path: > 512 complex rows / > 2048 edges), IoU-only and with features."""
    for with_reid in (False, True):
        infos = _run(ctx, SceneConfig(n_ids=700, feat_dim=2048, seed=17, pitch_x=12.0, pitch_y=16.0, walk=1.0,
                                      size_jitter=0.5, low_frac=0.1, drop_frac=0.05, with_features=with_reid),
                     frames=4, with_reid=with_reid)
        assert infos[-1]["n_pool"] >= 600


def test_out_of_range_coordinates(ctx):
    """Coordinates beyond the 15-bit integer screen of the association epilogue (and negative track boxes)
    must fall back to the exact path, never lose a candidate."""
    ctx.tracker_reset()
    oracle = O.OracleBoTSORT()
    rng = np.random.default_rng(3)
    n = 40
    base = np.stack([rng.uniform(0, 200, n), rng.uniform(0, 200, n)], axis=1)
    base[: n // 2] += 40000.0                      # half of the scene far beyond 32767 px
    wh = np.stack([rng.uniform(30, 60, n), rng.uniform(50, 90, n)], axis=1)
    feats_id = rng.standard_normal((n, 2048)).astype(np.float32)
    feats_id /= np.linalg.norm(feats_id, axis=1, keepdims=True)
    vel = np.zeros_like(base)
    vel[n // 2:] = -25.0                            # the near half drifts out of the frame: negative track boxes
    for k in range(8):
        pos = base + vel * k + rng.uniform(-2, 2, base.shape)
        boxes = np.hstack([np.maximum(pos, 0), np.maximum(pos + wh, 1)]).astype(np.int32)
        keep = (boxes[:, 2] > boxes[:, 0] + 4) & (boxes[:, 3] > boxes[:, 1] + 4)
        perm = rng.permutation(np.nonzero(keep)[0])
        f = feats_id[perm] + 0.005 * rng.standard_normal((len(perm), 2048)).astype(np.float32)
        f /= np.linalg.norm(f, axis=1, keepdims=True)
        scores = np.full(len(perm), 0.95, np.float32)
        oracle.update_arrays(boxes[perm], scores, f.astype(np.float32))
        ctx.update_arrays(boxes[perm], scores, f.astype(np.float32))
        _compare_frame(ctx, oracle, k + 1)
