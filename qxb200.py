"""Import shim: ``import qxb200`` loads the package that lives in ``qxtools.jl_b200/``
(a directory name the task fixes but Python cannot import directly)."""
import importlib.util as _u
import os as _os
import sys as _sys

_here = _os.path.dirname(_os.path.abspath(__file__))
_pkg = _os.path.join(_here, "qxtools.jl_b200")
_spec = _u.spec_from_file_location("qxb200", _os.path.join(_pkg, "__init__.py"),
                                   submodule_search_locations=[_pkg])
_mod = _u.module_from_spec(_spec)
_sys.modules["qxb200"] = _mod
_spec.loader.exec_module(_mod)
