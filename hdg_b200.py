"""Loader: imports the package directory `hdiscontinuousgalerkin.jl_b200/` under the importable
name `hdg_b200` (the directory name required by the project layout contains a dot)."""
import importlib.util as _u
import os as _os
import sys as _sys

_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "hdiscontinuousgalerkin.jl_b200")
_spec = _u.spec_from_file_location("hdg_b200", _os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = _u.module_from_spec(_spec)
_sys.modules["hdg_b200"] = _mod
_spec.loader.exec_module(_mod)
