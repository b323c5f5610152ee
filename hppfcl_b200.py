"""Import alias: `import hppfcl_b200` loads the package directory `hpp-fcl_b200/`
(whose name, fixed by the project layout, is not a valid Python identifier)."""
import importlib.util as _u
import os as _os
import sys as _sys

_d = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "hpp-fcl_b200")
_spec = _u.spec_from_file_location("hppfcl_b200", _os.path.join(_d, "__init__.py"),
                                   submodule_search_locations=[_d])
_mod = _u.module_from_spec(_spec)
_sys.modules["hppfcl_b200"] = _mod
_spec.loader.exec_module(_mod)
