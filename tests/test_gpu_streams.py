"""GPU parity of the round-2 entry points against the CPU oracle: fp16 feature ingest, several video streams
per launch (bt_update_streams), the submit / step split, concurrent contexts on host threads, the face
similarity term, non-unit-norm features, scores exactly on the thresholds, a full track store, and the exact
re-costing of near-tied appearance costs."""
import threading

import numpy as np
import pytest

import botsort_b200 as bs
from botsort_b200.synthetic import SceneConfig, SyntheticScene
from oracle import oracle_np as O
from test_gpu_tracker import INT_FIELDS, _compare_frame

pytestmark = pytest.mark.gpu


def _frames(cfg, count):
    sc = SyntheticScene(cfg)
    return [sc.next_frame() for _ in range(count)]


@pytest.mark.parametrize("feat_dim,flags", [(2048, 0), (256, 0), (2048, bs._lib.BT_FLAG_SIMT_SIM)])
def test_fp16_feature_ingest(feat_dim, flags):
    """BT_F16 rows = the reference fed with feats.astype(float32): the tensor cores multiply those values exactly."""
    ctx = bs.Context(max_tracks=512, max_dets=512, feat_dim=feat_dim, flags=flags)
    try:
        frames = _frames(SceneConfig(n_ids=150, feat_dim=feat_dim, seed=41, low_frac=0.15, drop_frac=0.1, mid_frac=0.05,
                                     newcomer_every=3, pitch_x=45.0, pitch_y=80.0), 25)
        oracle = O.OracleBoTSORT()
        for k, fr in enumerate(frames):
            f16 = fr["feats"].astype(np.float16)
            oracle.update_arrays(fr["boxes"], fr["scores"], f16.astype(np.float32))
            ctx.update_arrays(fr["boxes"], fr["scores"], f16)
            _compare_frame(ctx, oracle, k + 1)
        curr, smooth = ctx.get_track_features(0)
        assert np.max(np.abs(curr - np.array([t.curr_feat for t in oracle.tracked]))) <= 1e-6
        assert np.max(np.abs(smooth - np.array([t.smooth_feat for t in oracle.tracked]))) <= 1e-6
    finally:
        ctx.close()


@pytest.mark.parametrize("dtype", [np.float16, np.float32])
def test_four_streams_one_launch(dtype):
    """Video streams are a leading batch dimension: one bt_update_streams call steps four independent
    trackers of different sizes; each must equal its own oracle (ids are per stream)."""
    S = 4
    ctx = bs.Context(max_tracks=640, max_dets=640, feat_dim=2048, n_streams=S)
    try:
        cfgs = [SceneConfig(n_ids=n, feat_dim=2048, seed=60 + i, low_frac=0.1 + 0.05 * i, drop_frac=0.05 * (i + 1),
                            mid_frac=0.04, newcomer_every=2 + i, pitch_x=40.0 + 10 * i, pitch_y=75.0 + 12 * i)
                for i, n in enumerate((500, 37, 260, 129))]
        seqs = [_frames(c, 14) for c in cfgs]
        oracles = [O.OracleBoTSORT() for _ in range(S)]
        for k in range(14):
            frs = [seqs[s][k] for s in range(S)]
            feats = [fr["feats"].astype(dtype) for fr in frs]
            if k == 5:   # a frame without detections in one stream, a nearly empty one in another
                frs[1] = {"boxes": np.zeros((0, 4), np.int32), "scores": np.zeros(0, np.float32)}
                feats[1] = np.zeros((0, 2048), dtype)
                frs[3] = {"boxes": frs[3]["boxes"][:3], "scores": frs[3]["scores"][:3]}
                feats[3] = feats[3][:3]
            for s in range(S):
                oracles[s].update_arrays(frs[s]["boxes"], frs[s]["scores"], feats[s].astype(np.float32))
            infos = ctx.update_streams(list(range(S)), [f["boxes"] for f in frs], [f["scores"] for f in frs], feats)
            assert [i["frame_id"] for i in infos] == [k + 1] * S
            for s in range(S):
                _compare_frame(ctx, oracles[s], k + 1, stream=s)
        # a subset of the streams, in another order
        for k in range(3):
            frs = {s: _frames(cfgs[s], 1)[0] for s in (2, 0)}
            for s in (2, 0):
                oracles[s].update_arrays(frs[s]["boxes"], frs[s]["scores"], frs[s]["feats"].astype(dtype).astype(np.float32))
            ctx.update_streams([2, 0], [frs[2]["boxes"], frs[0]["boxes"]], [frs[2]["scores"], frs[0]["scores"]],
                               [frs[2]["feats"].astype(dtype), frs[0]["feats"].astype(dtype)])
            for s in (2, 0):
                _compare_frame(ctx, oracles[s], 15 + k, stream=s)
    finally:
        ctx.close()


def test_submit_step_pipeline_matches_plain_updates():
    """bt_submit_streams(frame k+1) before bt_step_streams(frame k): same tracks as bt_update_streams."""
    cfg = SceneConfig(n_ids=300, feat_dim=2048, seed=77, low_frac=0.1, drop_frac=0.08, newcomer_every=3)
    frames = _frames(cfg, 12)
    f16 = [fr["feats"].astype(np.float16) for fr in frames]
    ctx = bs.Context(max_tracks=512, max_dets=512, feat_dim=2048)
    try:
        ref = []
        for fr, f in zip(frames, f16):
            ctx.update_arrays(fr["boxes"], fr["scores"], f)
            ref.append((ctx.get_tracks(0, with_state=True), ctx.get_tracks(1)))
        ctx.tracker_reset()
        keep = [ctx.submit_streams([0], [frames[0]["boxes"]], [frames[0]["scores"]], [f16[0]])]
        with pytest.raises(Exception):      # at most two frames in flight per stream
            ctx.submit_streams([0], [frames[1]["boxes"]], [frames[1]["scores"]], [f16[1]])
            ctx.submit_streams([0], [frames[2]["boxes"]], [frames[2]["scores"]], [f16[2]])
        # the failed third submit changed nothing: frames 0 and 1 are queued
        for k in range(len(frames)):
            if k >= 1 and k + 1 < len(frames):
                keep.append(ctx.submit_streams([0], [frames[k + 1]["boxes"]], [frames[k + 1]["scores"]], [f16[k + 1]]))
            info = ctx.step_streams([0])[0]
            assert info["frame_id"] == k + 1
            gt, gl = ctx.get_tracks(0, with_state=True), ctx.get_tracks(1)
            for key in INT_FIELDS:
                np.testing.assert_array_equal(gt[key], ref[k][0][key])
                np.testing.assert_array_equal(gl[key], ref[k][1][key])
            np.testing.assert_array_equal(gt["mean"], ref[k][0]["mean"])
        with pytest.raises(Exception):
            ctx.step_streams([0])           # nothing submitted
    finally:
        ctx.close()


def test_four_contexts_on_four_host_threads():
    """One ctx per host thread, all on the same GPU, stepping concurrently (ctypes releases the GIL)."""
    n_threads, n_frames = 4, 15
    seqs = [_frames(SceneConfig(n_ids=120 + 40 * i, feat_dim=2048, seed=90 + i, low_frac=0.15, drop_frac=0.1,
                                newcomer_every=3), n_frames) for i in range(n_threads)]
    errors = []
    start = threading.Barrier(n_threads)

    def work(i):
        try:
            ctx = bs.Context(max_tracks=512, max_dets=512, feat_dim=2048)
            oracle = O.OracleBoTSORT()
            start.wait()
            for k, fr in enumerate(seqs[i]):
                oracle.update_arrays(fr["boxes"], fr["scores"], fr["feats"])
                ctx.update_arrays(fr["boxes"], fr["scores"], fr["feats"])
                _compare_frame(ctx, oracle, k + 1)
            ctx.close()
        except BaseException as e:     # noqa: BLE001
            errors.append((i, repr(e)[:400]))

    threads = [threading.Thread(target=work, args=(i,)) for i in range(n_threads)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors


def test_face_similarity_term():
    """demo:1541-1546: emb_dists_comp = min(body, face) decides the appearance gate.  Random face similarities
    (mostly 0 = no face seen, some high) against the oracle's fuse_stage1 with the same matrix."""
    cfg = SceneConfig(n_ids=90, feat_dim=2048, seed=101, low_frac=0.1, drop_frac=0.1, pitch_x=40.0, pitch_y=70.0,
                      feat_noise=0.03)
    frames = _frames(cfg, 14)
    rng = np.random.default_rng(5)
    ctx = bs.Context(max_tracks=256, max_dets=256, feat_dim=2048)
    try:
        oracle = O.OracleBoTSORT()
        for k, fr in enumerate(frames):
            n_pool = len([t for t in oracle.tracked if t.is_activated]) + len(oracle.lost)
            m = len(fr["boxes"])
            face = np.zeros((n_pool, m), np.float32)
            hot = rng.random((n_pool, m)) < 0.02
            face[hot] = rng.uniform(0.6, 0.99, int(hot.sum())).astype(np.float32)
            oracle.update_arrays(fr["boxes"], fr["scores"], fr["feats"], face_sims=face if n_pool else None)
            ctx.update_arrays(fr["boxes"], fr["scores"], fr["feats"], face_sim=face if n_pool else None)
            _compare_frame(ctx, oracle, k + 1)
    finally:
        ctx.close()


def test_non_unit_norm_features():
    """The first association sees the RAW encoder rows (demo:1453-1460), tracks keep row / ||row|| (demo:497-502),
    the unconfirmed association compares normalised rows (demo:1593-1599): feed rows of norm 0.8 .. 1.2."""
    cfg = SceneConfig(n_ids=120, feat_dim=2048, seed=111, low_frac=0.1, drop_frac=0.1, newcomer_every=2)
    frames = _frames(cfg, 16)
    rng = np.random.default_rng(7)
    for dtype in (np.float32, np.float16):
        ctx = bs.Context(max_tracks=256, max_dets=256, feat_dim=2048)
        try:
            oracle = O.OracleBoTSORT()
            for k, fr in enumerate(frames):
                scale = rng.uniform(0.8, 1.2, (len(fr["boxes"]), 1)).astype(np.float32)
                f = (fr["feats"] * scale).astype(dtype)
                oracle.update_arrays(fr["boxes"], fr["scores"], f.astype(np.float32))
                ctx.update_arrays(fr["boxes"], fr["scores"], f)
                _compare_frame(ctx, oracle, k + 1)
        finally:
            ctx.close()


def test_scores_exactly_on_the_thresholds():
    """float(score) is compared with Python doubles (demo:1022, 1501, 1531, 1617): float32(0.4) > 0.4 is True,
    float32(0.1) >= 0.1 is True, float32(0.9) < 0.9 is False."""
    cfg = SceneConfig(n_ids=40, feat_dim=2048, seed=121)
    frames = _frames(cfg, 6)
    ctx = bs.Context(max_tracks=128, max_dets=128, feat_dim=2048)
    try:
        oracle = O.OracleBoTSORT()
        edge = np.array([0.1, 0.4, 0.9, np.nextafter(np.float32(0.1), np.float32(0)), np.nextafter(np.float32(0.4), np.float32(0)),
                         np.nextafter(np.float32(0.9), np.float32(0))], dtype=np.float32)
        for k, fr in enumerate(frames):
            sc = fr["scores"].copy()
            if k >= 1:
                sc[: len(edge)] = np.roll(edge, k)
            if k == 3:      # newcomers exactly on the birth threshold
                extra_b = fr["boxes"][:3].copy(); extra_b[:, [0, 2]] += 3000
                fr = dict(fr, boxes=np.vstack([fr["boxes"], extra_b]), feats=np.vstack([fr["feats"], fr["feats"][:3][:, ::-1]]))
                sc = np.concatenate([sc, np.array([0.9, np.nextafter(np.float32(0.9), np.float32(0)), 0.95], np.float32)])
            oracle.update_arrays(fr["boxes"], sc, np.ascontiguousarray(fr["feats"]))
            ctx.update_arrays(fr["boxes"], sc, np.ascontiguousarray(fr["feats"]))
            _compare_frame(ctx, oracle, k + 1)
    finally:
        ctx.close()


def test_full_track_store_skips_births_and_keeps_tracking():
    """The reference's lists are unbounded; a full store drops births (reported) and the frame stays consistent."""
    cfg = SceneConfig(n_ids=200, feat_dim=256, seed=131, drop_frac=0.05)
    frames = _frames(cfg, 8)
    ctx = bs.Context(max_tracks=128, max_dets=256, feat_dim=256)
    try:
        info = ctx.update_arrays(frames[0]["boxes"], frames[0]["scores"], frames[0]["feats"])
        assert info["n_births"] == 128 and info["n_births_skipped"] == len(frames[0]["boxes"]) - 128
        ids0 = ctx.get_tracks(0)["ids"].copy()
        for fr in frames[1:]:
            info = ctx.update_arrays(fr["boxes"], fr["scores"], fr["feats"])
            tr = ctx.get_tracks(0)
            assert len(tr["ids"]) + info["n_lost"] <= 128
            assert info["n_matches1"] >= 100            # the stored tracks keep being followed
            assert set(tr["ids"]) <= set(range(1, 1 + 128 + 8 * 200))
        assert len(set(ids0) & set(ctx.get_tracks(0)["ids"])) >= 100
    finally:
        ctx.close()


def _near_tie_scene(n_pairs, d, rng, delta):
    """Pairs of identities with almost the same appearance (twins) standing next to each other: for each pair the
    2 x 2 block of appearance costs decides who is who, and the two candidate assignments differ by ~delta --
    far below the 3e-5 error of fp16 operands, far above float32 round-off."""
    g = rng.standard_normal((n_pairs, d)).astype(np.float32)
    g /= np.linalg.norm(g, axis=1, keepdims=True)
    twin = g + delta * rng.standard_normal((n_pairs, d)).astype(np.float32)
    twin /= np.linalg.norm(twin, axis=1, keepdims=True)
    ident = np.empty((2 * n_pairs, d), np.float32)
    ident[0::2], ident[1::2] = g, twin
    cx = 200.0 + 260.0 * np.repeat(np.arange(n_pairs), 2) + np.tile([0.0, 38.0], n_pairs)
    cy = np.full(2 * n_pairs, 400.0)
    return ident, cx, cy


@pytest.mark.parametrize("dtype", [np.float32, np.float16])
def test_near_tied_appearance_costs_are_recosted_exactly(dtype, monkeypatch):
    """SURVEY hard part 2: entries that decide an assignment are re-costed from the fp32 operands (float64
    accumulation) inside the LAP, so twins whose costs differ by ~1e-6 are told apart exactly like the oracle does."""
    rng = np.random.default_rng(2024)
    n_pairs, d = 48, 2048
    ident, cx, cy = _near_tie_scene(n_pairs, d, rng, delta=3e-4)
    n = 2 * n_pairs

    def frame(k):
        r = np.random.default_rng(1000 + k)
        x = cx + r.uniform(-2, 2, n); y = cy + r.uniform(-2, 2, n)
        boxes = np.stack([x - 30, y - 60, x + 30, y + 60], axis=1).astype(np.int32)
        f = ident + 0.002 * r.standard_normal((n, d)).astype(np.float32)
        f /= np.linalg.norm(f, axis=1, keepdims=True)
        perm = r.permutation(n)
        return boxes[perm], np.full(n, 0.95, np.float32), np.ascontiguousarray(f[perm]).astype(dtype)

    def run(ctx):
        oracle = O.OracleBoTSORT()
        bad = 0
        for k in range(12):
            b, s, f = frame(k)
            oracle.update_arrays(b, s, f.astype(np.float32))
            ctx.update_arrays(b, s, f)
            got = ctx.get_matches(1)
            ref = oracle.last["matches1"].astype(np.int32)
            if got.shape != ref.shape or not np.array_equal(got, ref):
                bad += 1
                ctx.tracker_reset(); oracle = O.OracleBoTSORT()    # resynchronise
        return bad

    ctx = bs.Context(max_tracks=256, max_dets=256, feat_dim=d)
    try:
        assert run(ctx) == 0
    finally:
        ctx.close()
    if dtype == np.float32:
        # the test has teeth: without the exact re-costing the fp16-rounded operands DO flip some of these
        monkeypatch.setenv("BT_NO_REFINE", "1")
        ctx = bs.Context(max_tracks=256, max_dets=256, feat_dim=d)
        try:
            assert run(ctx) > 0
        finally:
            ctx.close()


_SWITCHES = ["BT_GRAPH", "BT_DIRECT_RESULT", "BT_CTRL_COPY", "BT_NO_PREBUILD", "BT_COPY_INLINE", "BT_NO_L2_PREFETCH", "BT_NO_PDL"]


@pytest.mark.gpu
@pytest.mark.parametrize("switch", _SWITCHES)
def test_runtime_switches_do_not_change_the_tracking(monkeypatch, switch):
    """Every A/B switch of the frame step (captured frame graph, LAP kernel publishing the
    assignments itself, control block by memcpy, pool lists built at the start of the step, read-back on the main
    stream, no operand prefetch, no programmatic dependent launch) tracks exactly like the default path: ids,
    states and boxes over a sequence with births, lost and re-found tracks."""
    scene = SyntheticScene(SceneConfig(n_ids=150, feat_dim=512, seed=31, drop_frac=0.1, newcomer_every=3))
    frames = [scene.next_frame() for _ in range(12)]
    out = []
    for on in (False, True):
        for name in _SWITCHES:
            monkeypatch.delenv(name, raising=False)
        if on:
            monkeypatch.setenv(switch, "1")
        ctx = bs.Context(max_tracks=512, max_dets=512, feat_dim=512)
        ctx.tracker_reset()
        seq = []
        for f in frames:
            ctx.update_arrays(f["boxes"], f["scores"], f["feats"].astype(np.float16))
            tr = ctx.get_tracks(0)
            seq.append((tr["ids"].copy(), tr["state"].copy(), tr["tlbr"].copy()))
        out.append(seq)
        ctx.close()
    for (i0, s0, b0), (i1, s1, b1) in zip(*out):
        assert np.array_equal(i0, i1) and np.array_equal(s0, s1)
        np.testing.assert_array_equal(b0, b1)
