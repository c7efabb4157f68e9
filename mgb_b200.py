"""Import shim: the product package lives in the directory ``multigridbarriermpi.jl_b200/`` (the
name the build contract fixes); a dot is not importable, so this module loads it under the
importable name ``mgb_b200``."""
import importlib.util
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
_pkg_dir = os.path.join(_here, "multigridbarriermpi.jl_b200")
_spec = importlib.util.spec_from_file_location(
    "mgb_b200", os.path.join(_pkg_dir, "__init__.py"), submodule_search_locations=[_pkg_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["mgb_b200"] = _mod
_spec.loader.exec_module(_mod)
