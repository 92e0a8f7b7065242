"""BT_YOLOX_DEBUG=1 python tools/dbg_yolox.py -- phase timestamps of the YOLOX post-process cluster kernel."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import botsort_b200 as bs
from botsort_b200._lib import BT_DEVICE
from botsort_b200.synthetic import DetectorScene
ctx = bs.Context(max_tracks=256, max_dets=256, feat_dim=256)
sc = DetectorScene(k=40, seed=1234)
ycfg = bs.BtYoloxConfig(); ctx.lib.bt_default_yolox_config(C.byref(ycfg))
dev = torch.device("cuda", 0)
out = torch.zeros((256, 6), dtype=torch.float64, device=dev); cnt = torch.zeros(1, dtype=torch.int32, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for i in range(4):
    frame, raw = sc.next_frame()
    d = torch.from_numpy(raw).to(dev)
    flush.zero_(); torch.cuda.synchronize()
    ctx._check(ctx.lib.bt_yolox_postprocess(ctx.h, C.c_void_p(d.data_ptr()), C.byref(ycfg), C.c_void_p(out.data_ptr()), 256,
                                            C.c_void_p(cnt.data_ptr()), BT_DEVICE))
    ctx.sync(); print("--", int(cnt.cpu()[0]))
ctx.close()
