"""Import shim: the package directory is named ``exponentialutilities.jl_b200`` (with a dot, as the
project layout prescribes), which ``import`` cannot spell.  ``import eu_b200`` loads that directory
as the package ``eu_b200``."""
import importlib.util as _ilu
import os as _os
import sys as _sys

_pkg_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "exponentialutilities.jl_b200")
_spec = _ilu.spec_from_file_location("eu_b200", _os.path.join(_pkg_dir, "__init__.py"),
                                     submodule_search_locations=[_pkg_dir])
_mod = _ilu.module_from_spec(_spec)
_sys.modules["eu_b200"] = _mod
_spec.loader.exec_module(_mod)
