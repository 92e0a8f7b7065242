"""botsort_b200 -- B200-native BoT-SORT per-frame tracking hot path.

Drop-in mirror of the tracker API of PINTO0309/BoT-SORT-ONNX-TensorRT's
`demo_bottrack_onnx_tflite.py` (KalmanFilter, STrack, BoTSORT, iou_distance, linear_assignment, ...)
over hand-written sm_100a CUDA kernels reached through the C ABI in `include/botsort_b200.h`.
Nothing here computes on the CPU: importing works anywhere, but creating a `Context` (or any
object that needs one) requires the built `libbotsort_b200.so` and a B200.
"""
__version__ = "0.1.0"

from ._lib import (BT_DEVICE, BT_HOST, BT_FLAG_NO_F32_FEATURES, BT_FLAG_SIMT_SIM, BotsortError, BtConfig,  # noqa: F401
                   BtFrameInfo, BtYoloxConfig, Context, load_library)
