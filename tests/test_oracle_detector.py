"""Pins the detector-side oracle: the cv2.resize restatement bit-exactly against cv2 itself
(cv2 is part of the image), and the YOLOX decode+NMS restatement through its properties
(the in-graph arithmetic is not in the reference tree: parity unpinned by the reference)."""
import numpy as np
import pytest

from oracle import detector_np as Dn


def test_resize_restatement_is_bit_exact_against_cv2():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(0)
    shapes = [(512, 256), (1024, 512), (256, 128), (128, 64), (2, 2), (1, 300), (300, 1)]
    shapes += [(int(rng.integers(2, 400)), int(rng.integers(2, 300))) for _ in range(60)]
    for h, w in shapes:
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        np.testing.assert_array_equal(Dn.resize_linear_u8(img, 256, 128), cv2.resize(img, (128, 256)), err_msg=f"{h}x{w}")


def test_yolox_restatement_properties():
    rng = np.random.default_rng(1)
    boxes = [(50, 60, 120, 260), (300, 100, 380, 300), (400, 50, 420, 80), (10, 300, 200, 470)]
    raw = Dn.synth_yolox_head(rng, boxes, classes=[0, 0, 3, 1], scores=[0.95, 0.6, 0.5, 0.9])
    det = Dn.yolox_decode_nms(raw)
    assert len(det) > 0
    for c in np.unique(det[:, 0]):
        rows = det[det[:, 0] == c]
        assert len(rows) <= 50 and np.all(np.diff(rows[:, 1]) <= 0) and np.all(rows[:, 1] > 0.15)
        for i in range(len(rows)):                      # survivors do not overlap above the NMS threshold
            iou = Dn._iou32(rows[i, 2:6], rows[:, 2:6])
            iou[i] = 0
            assert np.all(iou <= np.float32(0.80))
    assert np.all(np.diff(det[:, 0]) >= 0)
    post = Dn.yolox_postprocess(raw, img_h=720, img_w=1280)
    assert np.all(post[:, 1] > 0.35) and np.all(post[:, 2:] == np.floor(post[:, 2:]))
    planted = {(0, 0.95), (0, 0.6), (3, 0.5), (1, 0.9)}
    for c, s in planted:                                # each planted box survives once, its duplicate is suppressed
        hit = [r for r in post if int(r[0]) == c and abs(r[1] - s) < 2e-3]
        assert len(hit) == 1, (c, s, post)


def test_crop_preprocess_layout():
    rng = np.random.default_rng(2)
    frame = rng.integers(0, 256, (120, 160, 3), dtype=np.uint8)
    out = Dn.crop_preprocess(frame, np.array([[0, 0, 128, 120], [10, 10, 10, 50]]), 256, 128)
    assert out.shape == (2, 3, 256, 128) and out.dtype == np.float32
    assert np.all(out[1] == 0)                          # empty crop
    # exact-size crop: resize is the identity, channel 0 of the output is R = source channel 2
    frame2 = rng.integers(0, 256, (256, 128, 3), dtype=np.uint8)
    o = Dn.crop_preprocess(frame2, np.array([[0, 0, 128, 256]]))
    ref = ((frame2[..., 2] / 255.0 - np.float32(0.485)) / np.float32(0.229)).astype(np.float32)
    np.testing.assert_array_equal(o[0, 0], ref)
