"""Import shim: the product package lives in `road-vision-system_b200/` (a hyphen is not importable).

`import rvb200` gives that package under the name `rvb200`.
"""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "road-vision-system_b200")
_spec = importlib.util.spec_from_file_location("rvb200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["rvb200"] = _mod
_spec.loader.exec_module(_mod)
