"""Import shim: the product package lives in the directory `lgar-py_b200/` (the name the
project brief fixes), which is not a valid Python identifier.  `import lgar_b200` loads it."""
import importlib.util
import os
import sys

_pkg_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lgar-py_b200")
_spec = importlib.util.spec_from_file_location(
    "lgar_b200", os.path.join(_pkg_dir, "__init__.py"), submodule_search_locations=[_pkg_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["lgar_b200"] = _mod
_spec.loader.exec_module(_mod)
