"""Import alias: `import lowrankmodels_b200 as lrm` loads the package that lives in the directory
`lowrankmodels.jl_b200/` (a name Python's import statement cannot spell because of the dot)."""
import importlib.util as _u
import os as _os
import sys as _sys

_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "lowrankmodels.jl_b200")
_spec = _u.spec_from_file_location(__name__, _os.path.join(_dir, "__init__.py"),
                                   submodule_search_locations=[_dir])
_mod = _u.module_from_spec(_spec)
_sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
