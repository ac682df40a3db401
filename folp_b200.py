"""Import alias for the package directory `firstorderlp.jl_b200/`.

The directory name required by the repo layout contains a dot, so it cannot be
imported by name. `import folp_b200` executes this shim, which loads
`firstorderlp.jl_b200/__init__.py` as the package `folp_b200` (sub-modules
resolve through its __path__) and replaces itself in sys.modules.
"""
import importlib.util as _u
import os as _os
import sys as _sys

_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "firstorderlp.jl_b200")
_spec = _u.spec_from_file_location(
    "folp_b200", _os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir]
)
_mod = _u.module_from_spec(_spec)
_sys.modules["folp_b200"] = _mod
_spec.loader.exec_module(_mod)
