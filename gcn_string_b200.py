"""Import shim: the package directory is ``gcn-string_b200/`` (the layout the task
prescribes); a hyphen is not importable, so this module replaces itself in
``sys.modules`` with that directory loaded as the package ``gcn_string_b200``."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "gcn-string_b200")
_spec = importlib.util.spec_from_file_location(
    __name__, os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
