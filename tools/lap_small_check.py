"""Small dense / sparse assignments through the CTA-wide component solver against the JV port -- sized for
compute-sanitizer runs (racecheck: 0 hazards, memcheck: 0 errors on the final round-2 build)."""
import sys, numpy as np
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import botsort_b200 as bs
from oracle import oracle_np as O
ctx = bs.Context(max_tracks=512, max_dets=512, feat_dim=64)
rng = np.random.default_rng(5)
for n, dens in ((200, 1.0), (300, 0.2)):
    c = rng.uniform(0, 1, (n, n)); c[rng.uniform(size=(n, n)) > dens] = 1.0
    x, y = ctx.lapjv(c, 0.8)
    rx, ry = O.lapjv_extended(c, 0.8, "jv")
    print(n, dens, "exact", bool(np.array_equal(x, rx) and np.array_equal(y, ry)), int((x >= 0).sum()))
