"""Import shim: `import bvc_b200` loads the package in ./baby-vision-curriculum_b200/ (hyphenated directory)."""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "baby-vision-curriculum_b200")
_NAME = "bvc_b200"
_spec = importlib.util.spec_from_file_location(_NAME, os.path.join(_PKG_DIR, "__init__.py"),
                                               submodule_search_locations=[_PKG_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules[_NAME] = _mod
_spec.loader.exec_module(_mod)
