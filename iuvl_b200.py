"""Import alias: ``import iuvl_b200`` loads the package in ``interactable-unified-vision-language_b200/``."""
import importlib.util
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
_pkg_dir = os.path.join(_here, "interactable-unified-vision-language_b200")
_spec = importlib.util.spec_from_file_location(
    "iuvl_b200", os.path.join(_pkg_dir, "__init__.py"), submodule_search_locations=[_pkg_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["iuvl_b200"] = _mod
_spec.loader.exec_module(_mod)
