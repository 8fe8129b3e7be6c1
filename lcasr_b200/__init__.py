"""Importable alias of the package directory ``long-context-asr_b200/`` (a hyphen cannot appear in
a Python module name).  ``import lcasr_b200`` executes long-context-asr_b200/__init__.py under this
name, so ``lcasr_b200.model``, ``lcasr_b200.ops`` ... resolve to the files in that directory."""
import importlib.util as _ilu
import os as _os
import sys as _sys

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "long-context-asr_b200")
_spec = _ilu.spec_from_file_location(__name__, _os.path.join(_real, "__init__.py"), submodule_search_locations=[_real])
_mod = _ilu.module_from_spec(_spec)
_sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
