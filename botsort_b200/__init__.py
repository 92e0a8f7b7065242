"""Importable alias of the product package.

The package directory is named `bot-sort-onnx-tensorrt_b200/` (repo layout contract), which is
not a valid Python identifier; this stub makes it importable as `botsort_b200` by pointing the
package search path at that directory and executing its `__init__.py`.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "bot-sort-onnx-tensorrt_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _os, _f, _real
