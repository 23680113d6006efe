"""Import shim: ``import mgicp_b200`` loads the package in
``point-cloud-registration-with-global-refinement_b200/`` (that name is not a Python identifier)."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "point-cloud-registration-with-global-refinement_b200")
_spec = importlib.util.spec_from_file_location("mgicp_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["mgicp_b200"] = _mod
_spec.loader.exec_module(_mod)
