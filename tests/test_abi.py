"""CPU checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/botsort_b200.h declares, and fails loudly (no CPU fallback) when there is no GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "botsort_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bt_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from botsort_b200._lib import EXPORTS, LIB_PATH, load_library
    assert os.path.exists(LIB_PATH), "build the library first: python __graft_entry__.py"
    lib = load_library()
    declared = _declared_symbols()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/botsort_b200.h but not exported"
    assert sorted(EXPORTS) == declared, "binding list and header disagree"
    assert lib.bt_version() == 200


def test_defaults_match_reference_constants():
    """demo:1268-1277, demo:1571, demo:1604, demo:1667, demo:473, demo:862 / model name demo:34."""
    import botsort_b200 as bs
    lib = bs.load_library()
    cfg = bs.BtConfig()
    lib.bt_default_config(ctypes.byref(cfg))
    # the thresholds are Python floats in the reference and doubles in bt_config: exact equality, no float32 detour
    assert (cfg.track_high_thresh, cfg.track_low_thresh, cfg.new_track_thresh) == (0.40, 0.1, 0.9)
    assert (cfg.match_thresh, cfg.second_thresh, cfg.unconfirmed_thresh) == (0.8, 0.5, 0.7)
    assert (cfg.proximity_thresh, cfg.appearance_thresh) == (0.5, 0.25)
    assert cfg.duplicate_iou_dist == 0.15 and cfg.track_buffer == 300 and cfg.frame_rate == 30
    assert cfg.ema_alpha == 0.9 and cfg.with_reid == 1
    y = bs.BtYoloxConfig()
    lib.bt_default_yolox_config(ctypes.byref(y))
    assert (y.in_h, y.in_w, y.num_classes, y.max_per_class) == (480, 640, 4, 50)
    assert np.float32(y.nms_score_thresh) == np.float32(0.15) and np.float32(y.nms_iou_thresh) == np.float32(0.80)
    assert np.float32(y.post_score_thresh) == np.float32(0.35)


def test_no_cpu_fallback():
    import botsort_b200 as bs
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present")
    with pytest.raises(bs.BotsortError, match="no CUDA device"):
        bs.Context(max_tracks=128, max_dets=128, feat_dim=64)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "bot-sort-onnx-tensorrt_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f"{f} imports the oracle"
                assert "oracle_np" not in src and "liboracle" not in src, f"{f} references the oracle"
