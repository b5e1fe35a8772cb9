"""Loader for the ``approximategps.jl_b200`` package (its directory name contains a dot, so it cannot be
imported with a plain ``import`` statement).  ``import agp_b200`` returns that package."""
import importlib.util
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
_pkg_dir = os.path.join(_here, "approximategps.jl_b200")
_spec = importlib.util.spec_from_file_location(
    "agp_b200", os.path.join(_pkg_dir, "__init__.py"), submodule_search_locations=[_pkg_dir]
)
_mod = importlib.util.module_from_spec(_spec)
sys.modules["agp_b200"] = _mod
_spec.loader.exec_module(_mod)
