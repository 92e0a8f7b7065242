"""Dense linear assignment through the C-ABI (bt_linear_assignment): wall time per call and exactness against the
JV port, for the sizes VERDICT r01 item 8 names.  python tools/lap_bench.py [out.json]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import botsort_b200 as bs
from oracle import oracle_np as O

ctx = bs.Context(max_tracks=2048, max_dets=2048, feat_dim=64)
out = []
for n, density, thresh in [(64, 1.0, 0.8), (512, 1.0, 0.8), (512, 0.05, 0.8), (1000, 1.0, 0.8), (2000, 1.0, 0.8), (2000, 0.01, 0.8)]:
    rng = np.random.default_rng(n)
    cost = rng.uniform(0.0, 1.0, (n, n))
    cost[rng.uniform(size=(n, n)) > density] = 1.0
    x, y = ctx.lapjv(cost, thresh)                      # warm-up
    t = []
    for _ in range(5):
        t0 = time.perf_counter(); x, y = ctx.lapjv(cost, thresh); t.append(time.perf_counter() - t0)
    t0 = time.perf_counter(); rx, ry = O.lapjv_extended(cost, thresh, "jv"); t_cpu = time.perf_counter() - t0
    rec = {"n": n, "density": density, "thresh": thresh, "gpu_ms_median": 1e3 * sorted(t)[2], "gpu_ms_min": 1e3 * min(t),
           "cpu_jv_port_ms": 1e3 * t_cpu, "matched": int((x >= 0).sum()), "exact": bool(np.array_equal(x, rx) and np.array_equal(y, ry)),
           "note": "wall clock of ctx.lapjv: H2D of the float64 cost (8 n^2 bytes), compaction, solve, D2H of x / y"}
    print(rec, flush=True)
    out.append(rec)
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
