"""Import alias for the package directory `pigeons.jl_b200/` (a dotted directory
name cannot be imported directly):  ``import pigeons_jl_b200 as pg``."""
import importlib.util as _ilu
import os as _os
import sys as _sys

_root = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "pigeons.jl_b200")
_spec = _ilu.spec_from_file_location(__name__, _os.path.join(_root, "__init__.py"),
                                     submodule_search_locations=[_root])
_mod = _ilu.module_from_spec(_spec)
_sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
