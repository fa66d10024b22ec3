"""Import shim: `import worldb200` loads the package that lives in `world-class_b200/`
(the directory name required by the build layout is not a valid Python identifier)."""
import importlib.util
import os
import sys

_pkg_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "world-class_b200")
_spec = importlib.util.spec_from_file_location(
    "worldb200", os.path.join(_pkg_dir, "__init__.py"), submodule_search_locations=[_pkg_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["worldb200"] = _mod
_spec.loader.exec_module(_mod)
