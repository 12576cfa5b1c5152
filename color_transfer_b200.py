"""Import shim: makes ``import color_transfer_b200`` resolve to the ``color-transfer_b200/``
directory (a hyphen is not a legal Python identifier)."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "color-transfer_b200")
_spec = importlib.util.spec_from_file_location(
    "color_transfer_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["color_transfer_b200"] = _mod
_spec.loader.exec_module(_mod)
