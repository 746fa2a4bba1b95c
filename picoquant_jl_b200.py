"""Import shim: the package directory is literally ``picoquant.jl_b200/`` (the
name the build contract fixes), which is not a valid Python identifier.  This
module loads that directory as the package ``picoquant_jl_b200`` so that
``import picoquant_jl_b200`` / ``from picoquant_jl_b200.host import ...`` work.
"""
import importlib.util
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
_pkg_dir = os.path.join(_here, "picoquant.jl_b200")
_spec = importlib.util.spec_from_file_location(
    __name__, os.path.join(_pkg_dir, "__init__.py"),
    submodule_search_locations=[_pkg_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
