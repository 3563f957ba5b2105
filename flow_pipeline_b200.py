"""Import shim: the package directory is named 'flow-pipeline_b200' (hyphen, as the
reference's name); this module makes it importable as `flow_pipeline_b200`."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("flow-pipeline_b200")
sys.modules[__name__] = _pkg
