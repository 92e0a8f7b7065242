"""Detector-side kernels (north_star (d)) against oracle/detector_np.py."""
import numpy as np
import pytest

from oracle import detector_np as Dn

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", range(6))
def test_yolox_postprocess(ctx, seed):
    import botsort_b200 as bs
    rng = np.random.default_rng(seed)
    k = int(rng.integers(3, 40))
    x1 = rng.uniform(0, 560, k); y1 = rng.uniform(0, 380, k)
    boxes = np.stack([x1, y1, x1 + rng.uniform(16, 200, k), y1 + rng.uniform(24, 300, k)], axis=1)
    boxes[:, 2] = np.minimum(boxes[:, 2], 639); boxes[:, 3] = np.minimum(boxes[:, 3], 479)
    raw = Dn.synth_yolox_head(rng, boxes, classes=rng.integers(0, 4, k), scores=rng.uniform(0.2, 0.97, k),
                              clutter=400)
    cfg = bs.BtYoloxConfig()
    ctx.lib.bt_default_yolox_config(cfg)
    cfg.img_h, cfg.img_w = 720, 1280
    got = ctx.yolox_postprocess(raw, cfg)
    ref = Dn.yolox_postprocess(raw, img_h=720, img_w=1280)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    np.testing.assert_array_equal(got[:, 0], ref[:, 0])                 # classes
    np.testing.assert_array_equal(got[:, 2:], ref[:, 2:])               # integer boxes: exact
    assert np.max(np.abs(got[:, 1] - ref[:, 1])) <= 1e-5                # scores


def test_yolox_postprocess_empty_and_cap(ctx):
    rng = np.random.default_rng(0)
    raw = Dn.synth_yolox_head(rng, [], [], [], clutter=0)
    assert ctx.yolox_postprocess(raw).shape == (0, 6)
    # > 50 confident non-overlapping boxes of one class: capped at 50 per class
    boxes = [(8 + 24 * (i % 26), 8 + 40 * (i // 26), 8 + 24 * (i % 26) + 16, 8 + 40 * (i // 26) + 30) for i in range(70)]
    raw = Dn.synth_yolox_head(rng, boxes, [0] * 70, list(np.linspace(0.5, 0.95, 70)), clutter=0)
    got = ctx.yolox_postprocess(raw)
    ref = Dn.yolox_postprocess(raw, img_h=480, img_w=640)
    assert (got[:, 0] == 0).sum() == (ref[:, 0] == 0).sum() <= 50
    np.testing.assert_array_equal(got[:, 2:], ref[:, 2:])


@pytest.mark.parametrize("seed", range(3))
def test_reid_crop_gather(ctx, seed):
    rng = np.random.default_rng(seed)
    h, w = int(rng.integers(200, 720)), int(rng.integers(300, 1280))
    frame = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    n = 24
    x1 = rng.integers(0, w - 4, n); y1 = rng.integers(0, h - 4, n)
    boxes = np.stack([x1, y1, x1 + rng.integers(1, 400, n), y1 + rng.integers(1, 600, n)], axis=1).astype(np.int32)
    boxes[0] = [0, 0, w, h]
    boxes[1] = [10, 10, 10 + 256, 10 + 512] if w > 300 and h > 530 else boxes[1]   # exact 2x downscale
    boxes[2] = [5, 5, 5, 90]                                                         # empty crop -> zeros
    got = ctx.reid_crop_gather(frame, boxes)
    ref = Dn.crop_preprocess(frame, boxes)
    np.testing.assert_array_equal(got, ref)


def test_yolox_postprocess_beyond_the_on_chip_window(ctx):
    """More candidates than the kernel's shared-memory window (1024 boxes per class), and the head of the list is a pile of
    near-identical boxes: one is kept, the rest must be suppressed, and the greedy selection has to
    continue over the remaining candidates exactly like ONNX NonMaxSuppression."""
    rng = np.random.default_rng(11)
    raw = Dn.synth_yolox_head(rng, [(100.4, 80.3, 220.6, 300.7), (400.2, 120.6, 520.3, 400.4)], [0, 1], [0.95, 0.9], clutter=2500)
    gx, gy, st = Dn.yolox_grid(480, 640)
    dup = np.nonzero(st == 8)[0][100:800]                      # 700 anchors that all decode to (almost) one box
    cx, cy, w, h = 320.0, 240.0, 150.0, 260.0
    raw[dup, 0] = (cx + rng.uniform(-1, 1, len(dup))) / 8 - gx[dup]
    raw[dup, 1] = (cy + rng.uniform(-1, 1, len(dup))) / 8 - gy[dup]
    raw[dup, 2] = np.log(w / 8)
    raw[dup, 3] = np.log(h / 8)
    p = np.linspace(0.999, 0.97, len(dup))                     # the highest scores of class 2
    raw[dup, 4] = np.log(p / (1 - p))
    raw[dup, 5:] = -6.0
    raw[dup, 7] = np.log(p / (1 - p))
    raw = raw.astype(np.float32)
    got = ctx.yolox_postprocess(raw)
    ref = Dn.yolox_postprocess(raw, img_h=480, img_w=640)
    assert got.shape == ref.shape and len(ref) > 60
    assert np.array_equal(got[:, [0, 2, 3, 4, 5]], ref[:, [0, 2, 3, 4, 5]])
    np.testing.assert_allclose(got[:, 1], ref[:, 1], atol=2e-6)
    assert (ref[:, 0] == 2).sum() >= 2                         # the duplicate pile collapsed, clutter of class 2 followed
