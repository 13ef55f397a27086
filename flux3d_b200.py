"""Import shim: the package directory is named ``flux3d.jl_b200`` (after the reference repo), which
is not a valid Python identifier, so it is registered here under the importable name ``flux3d_b200``.
``import flux3d_b200`` anywhere with the repo root on sys.path gives the real package."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "flux3d.jl_b200")
_spec = importlib.util.spec_from_file_location(
    "flux3d_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["flux3d_b200"] = _mod
_spec.loader.exec_module(_mod)
