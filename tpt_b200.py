"""Import shim: the package directory is named ``tiny-path-tracer_b200`` (hyphenated, as the
build contract requires), which the ``import`` statement cannot spell. ``import tpt_b200``
returns that package."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("tiny-path-tracer_b200")
sys.modules[__name__] = _pkg
