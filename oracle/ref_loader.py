"""TEST INFRASTRUCTURE ONLY -- loads the *real* reference module in place.

Only `tests/`, `oracle/gen_golden.py` and ad-hoc validation scripts may import this
file.  It works only where `/root/reference` exists (the build container); it is never
reachable from the product package and never runs on the GPU box.

What it does (SURVEY.md section 7 "Oracle recipe"):
  * injects a `lap` module into `sys.modules` before the import, because the reference
    does `import lap` at module scope (`/root/reference/demo_bottrack_onnx_tflite.py:16`)
    and `lap==0.4.0` (`/root/reference/Dockerfile:26`) is not installable here;
  * the shim's `lapjv(cost, extend_cost=True, cost_limit=thresh)` follows the published
    lap 0.4.0 `_lapjv.pyx` construction: an (N+M)x(N+M) matrix filled with
    `cost_limit / 2`, lower-right block 0, upper-left block = cost, solved exactly
    (here with `scipy.optimize.linear_sum_assignment`), then indices >= M (resp. N)
    are mapped to -1 and the vectors truncated.  Source of lap is NOT under
    /root/reference -> "parity unpinned by the reference" for tie-breaking; the optimum
    itself is pinned by the brute-force test in tests/test_oracle_lap.py.
  * imports `/root/reference/demo_bottrack_onnx_tflite.py` via sys.path (never copied).
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np

REFERENCE_DIR = os.environ.get("BOTSORT_REFERENCE_DIR", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_DIR, "demo_bottrack_onnx_tflite.py"))


def _make_lap_shim() -> types.ModuleType:
    from scipy.optimize import linear_sum_assignment

    def lapjv(cost, extend_cost=False, cost_limit=np.inf, return_cost=True):
        c = np.ascontiguousarray(cost, dtype=np.double)
        if c.ndim != 2:
            raise ValueError("2-dimensional array expected")
        n_rows, n_cols = c.shape
        if n_rows != n_cols and not extend_cost:
            raise ValueError("Square cost array expected. Pass extend_cost=True.")
        if cost_limit < np.inf:
            n = n_rows + n_cols
            ext = np.empty((n, n), dtype=np.double)
            ext[:] = cost_limit / 2.0
            ext[n_rows:, n_cols:] = 0
            ext[:n_rows, :n_cols] = c
        elif n_rows != n_cols:
            n = max(n_rows, n_cols)
            ext = np.zeros((n, n), dtype=np.double)
            ext[:n_rows, :n_cols] = c
        else:
            n = n_rows
            ext = c
        ri, ci = linear_sum_assignment(ext)
        x = np.empty(n, dtype=np.int32)
        y = np.empty(n, dtype=np.int32)
        x[ri] = ci
        y[ci] = ri
        opt = float(ext[ri, ci].sum())
        if cost_limit < np.inf or n_rows != n_cols:
            x = x.copy()
            y = y.copy()
            x[x >= n_cols] = -1
            y[y >= n_rows] = -1
            x = x[:n_rows]
            y = y[:n_cols]
        return (opt, x, y) if return_cost else (x, y)

    mod = types.ModuleType("lap")
    mod.lapjv = lapjv
    mod.__version__ = "0.4.0-shim(scipy-LSA)"
    return mod


_REF = None


def load_reference():
    """Return the reference module object (imported in place, cached)."""
    global _REF
    if _REF is not None:
        return _REF
    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_DIR}")
    if "lap" not in sys.modules:
        sys.modules["lap"] = _make_lap_shim()
    if REFERENCE_DIR not in sys.path:
        sys.path.insert(0, REFERENCE_DIR)
    import demo_bottrack_onnx_tflite as ref  # noqa: E402

    _REF = ref
    return ref
