"""Import alias: the package directory is ``mpc-code_b200/`` (not a valid identifier),
so ``import mpc_code_b200`` resolves here and loads that directory as the package."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "mpc-code_b200")
_spec = importlib.util.spec_from_file_location(
    "mpc_code_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["mpc_code_b200"] = _mod
_spec.loader.exec_module(_mod)
