"""Import shim: the package directory is `adtomo.jl_b200/` (a dot is not importable), so this
module loads it under the name `adtomo_jl_b200`."""
import importlib.util
import os
import sys

_p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "adtomo.jl_b200")
_spec = importlib.util.spec_from_file_location("adtomo_jl_b200", os.path.join(_p, "__init__.py"),
                                               submodule_search_locations=[_p])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["adtomo_jl_b200"] = _mod
_spec.loader.exec_module(_mod)
