"""BASELINE config 4 end to end on the GPU: synthetic 480x640 frame + raw YOLOX head ->
bt_yolox_postprocess (decode + NMS + _postprocess) -> body boxes -> bt_reid_crop_gather ->
stub FastReID (fixed seeded random projection of the 3x256x128 crop tensor, L2-normalised) ->
BoTSORT.update_arrays; the same pipeline runs on the CPU oracle.  ids / boxes exact, states 1e-4."""
import numpy as np
import pytest

from oracle import detector_np as Dn
from oracle import oracle_np as O

pytestmark = pytest.mark.gpu

D = 256   # stub ReID feature size (multiple of 64: tensor-core path)


def _scene(rng, k):
    x1 = rng.uniform(10, 500, k); y1 = rng.uniform(10, 250, k)
    w = rng.uniform(30, 90, k); h = rng.uniform(80, 200, k)
    return np.stack([x1, y1, np.minimum(x1 + w, 630), np.minimum(y1 + h, 470)], axis=1)


def _frame_and_head(rng, boxes, scores):
    frame = rng.integers(0, 256, (480, 640, 3), dtype=np.uint8)
    for i, b in enumerate(boxes.astype(int)):            # paste a distinct texture per identity
        patch_rng = np.random.default_rng(1000 + i)
        frame[b[1]:b[3], b[0]:b[2]] = patch_rng.integers(0, 256, (b[3] - b[1], b[2] - b[0], 3), dtype=np.uint8)
    classes = np.zeros(len(boxes), dtype=np.int64)
    raw = Dn.synth_yolox_head(rng, boxes, classes, scores, clutter=200)
    return frame, raw


def test_c4_detect_crop_reid_track(ctx):
    import botsort_b200 as bs
    rng = np.random.default_rng(0)
    proj = np.random.default_rng(7).standard_normal((3 * 256 * 128, D)).astype(np.float32) / 300.0
    k = 12
    base = _scene(rng, k)
    trk = bs.Context(max_tracks=256, max_dets=256, feat_dim=D)
    oracle = O.OracleBoTSORT()
    trk.tracker_reset()
    try:
        for f in range(8):
            boxes = base + rng.uniform(-3, 3, base.shape)
            scores = np.full(k, 0.96)
            if f > 2:
                scores[f % k] = 0.3                      # one low-score detection per frame (second association)
            frame, raw = _frame_and_head(rng, boxes, scores)
            # --- GPU pipeline ---
            det = ctx.yolox_postprocess(raw)             # rows: class, score, x1, y1, x2, y2
            det_o = Dn.yolox_postprocess(raw, img_h=480, img_w=640)
            assert det.shape == det_o.shape and np.array_equal(det[:, [0, 2, 3, 4, 5]], det_o[:, [0, 2, 3, 4, 5]])
            body = det[det[:, 0] == 0]
            b_int = body[:, 2:6].astype(np.int32)
            sc = body[:, 1].astype(np.float32)
            crops = ctx.reid_crop_gather(frame, b_int)
            crops_o = Dn.crop_preprocess(frame, b_int)
            np.testing.assert_array_equal(crops, crops_o)
            feats = crops.reshape(len(b_int), -1) @ proj          # stub encoder (host; inference is out of scope)
            feats /= np.linalg.norm(feats, axis=1, keepdims=True)
            feats = feats.astype(np.float32)
            trk.update_arrays(b_int, sc, feats)
            oracle.update_arrays(b_int, det_o[det_o[:, 0] == 0][:, 1].astype(np.float32), feats.copy())
            got = trk.get_tracks(0, with_state=True)
            ref = oracle.snapshot()["tracked"]
            np.testing.assert_array_equal(got["ids"], ref["ids"].astype(np.int32), err_msg=f"frame {f + 1}")
            np.testing.assert_array_equal(got["state"], ref["state"].astype(np.int32))
            if len(ref["ids"]):
                assert np.max(np.abs(got["mean"] - ref["mean"])) <= 1e-4
                assert np.max(np.abs(got["tlbr"] - ref["tlbr"])) <= 1e-4
        assert len(got["ids"]) >= k - 2
    finally:
        trk.close()


def test_c4_device_chained(ctx):
    """raw head -> bt_detect_stage (decode + NMS + _postprocess, body boxes / scores staged straight
    into the tracker's input buffers, crops into the encoder's input buffer) -> stub encoder ON THE DEVICE writing fp16
    rows into the tracker's feature buffer -> bt_update_streams(BT_DEVICE, m = max_per_class).  No host hop between
    the stages; the CPU oracle runs the same frames (it is fed the encoder's own fp16 output)."""
    import torch
    import botsort_b200 as bs
    from botsort_b200._lib import BT_DEVICE, BT_F16
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(1)
    proj = torch.from_numpy(np.random.default_rng(7).standard_normal((3 * 256 * 128, D)).astype(np.float32) / 300.0).to(dev)
    k = 14
    base = _scene(rng, k)
    trk = bs.Context(max_tracks=256, max_dets=256, feat_dim=D)
    oracle = O.OracleBoTSORT()
    trk.tracker_reset()
    ycfg = bs.BtYoloxConfig()
    trk.lib.bt_default_yolox_config(__import__("ctypes").byref(ycfg))
    mb = ycfg.max_per_class
    crops = torch.empty((mb, 3, 256, 128), dtype=torch.float32, device=dev)
    det_out = torch.zeros((256, 6), dtype=torch.float64, device=dev)
    det_cnt = torch.zeros(1, dtype=torch.int32, device=dev)

    class _Raw:
        def __init__(self, ptr, shape, typestr):
            self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (ptr, False), "version": 3}
    try:
        for f in range(8):
            boxes = base + rng.uniform(-3, 3, base.shape)
            scores = np.full(k, 0.96)
            if f > 2:
                scores[f % k] = 0.3
            frame, raw = _frame_and_head(rng, boxes, scores)
            d_frame = torch.from_numpy(frame).to(dev)
            d_raw = torch.from_numpy(raw).to(dev)
            pb, ps, pf = trk.input_buffers(0)
            torch.cuda.synchronize()
            trk.detect_stage(d_raw.data_ptr(), d_frame.data_ptr(), 480, 640, crops.data_ptr(), ycfg,
                             det_out_ptr=det_out.data_ptr(), max_out=256, det_count_ptr=det_cnt.data_ptr())
            trk.sync()                                          # (torch's stream below is not the ctx stream)
            feats16 = torch.as_tensor(_Raw(pf, (mb, D), "<f2"), device=dev)
            f32 = crops.reshape(mb, -1) @ proj                  # the stub encoder, on the device
            f32 = f32 / f32.norm(dim=1, keepdim=True).clamp_min(1e-12)
            feats16.copy_(f32.half())
            torch.cuda.synchronize()
            trk.update_streams_raw([0], [pb], [ps], [pf], [mb], BT_DEVICE, BT_F16)
            # --- the oracle: same decoded list, same crops, the encoder's fp16 rows ---
            det_o = Dn.yolox_postprocess(raw, img_h=480, img_w=640)
            n_det = int(det_cnt.cpu()[0])
            got_det = det_out.cpu().numpy()[:n_det]
            assert got_det.shape == det_o.shape and np.array_equal(got_det[:, [0, 2, 3, 4, 5]], det_o[:, [0, 2, 3, 4, 5]])
            body = det_o[det_o[:, 0] == 0]
            nb = len(body)
            b_int = body[:, 2:6].astype(np.int32)
            staged_boxes = torch.as_tensor(_Raw(pb, (mb, 4), "<i4"), device=dev).cpu().numpy()
            staged_scores = torch.as_tensor(_Raw(ps, (mb,), "<f4"), device=dev).cpu().numpy()
            np.testing.assert_array_equal(staged_boxes[:nb], b_int)
            assert np.all(staged_scores[nb:] == 0)
            np.testing.assert_array_equal(crops.cpu().numpy()[:nb], Dn.crop_preprocess(frame, b_int))
            oracle.update_arrays(b_int, staged_scores[:nb], feats16.cpu().numpy()[:nb].astype(np.float32))
            got = trk.get_tracks(0, with_state=True)
            ref = oracle.snapshot()["tracked"]
            np.testing.assert_array_equal(got["ids"], ref["ids"].astype(np.int32), err_msg=f"frame {f + 1}")
            np.testing.assert_array_equal(got["state"], ref["state"].astype(np.int32))
            np.testing.assert_array_equal(got["det_index"], ref["det_index"].astype(np.int32))
            if len(ref["ids"]):
                assert np.max(np.abs(got["mean"] - ref["mean"])) <= 1e-4
                assert np.max(np.abs(got["tlbr"] - ref["tlbr"])) <= 1e-4
        assert len(got["ids"]) >= k - 2
    finally:
        trk.close()
