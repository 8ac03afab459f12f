"""Import shim: the package directory is named `tiny-object-detection_b200` (not a valid Python
identifier), so `import tod_b200` loads it from that path."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tiny-object-detection_b200")
_spec = importlib.util.spec_from_file_location("tod_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["tod_b200"] = _mod
_spec.loader.exec_module(_mod)
