"""Import shim: `import mcm_b200 as mcm` loads the package that lives in the directory
`montecarlomeasurements.jl_b200/` (a dot in a directory name cannot be spelled in an import statement)."""
import importlib.util as _u
import os as _os
import sys as _sys

_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "montecarlomeasurements.jl_b200")
_spec = _u.spec_from_file_location("mcm_b200", _os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = _u.module_from_spec(_spec)
_sys.modules["mcm_b200"] = _mod
_spec.loader.exec_module(_mod)
