"""Loader: makes the package directory `differentialevolutionmcmc.jl_b200/` importable as
`demcmc_b200` (its own name contains a dot and cannot be an import name)."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "pkg")
_spec = importlib.util.spec_from_file_location("demcmc_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["demcmc_b200"] = _mod
_spec.loader.exec_module(_mod)
