"""Import shim: the package directory name contains hyphens (it mirrors the reference repo's
name), so it is loaded by path and re-exported here as ``csfm_b200``."""
import importlib
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)),
                        "compressed-fm-index-implementation-with-learned-optimizations_b200")
_NAME = "compressed_fm_index_implementation_with_learned_optimizations_b200"

if _NAME not in sys.modules:
    _spec = importlib.util.spec_from_file_location(_NAME, os.path.join(_PKG_DIR, "__init__.py"),
                                                   submodule_search_locations=[_PKG_DIR])
    _mod = importlib.util.module_from_spec(_spec)
    sys.modules[_NAME] = _mod
    _spec.loader.exec_module(_mod)
_mod = sys.modules[_NAME]
globals().update({k: getattr(_mod, k) for k in dir(_mod) if not k.startswith("__")})
workloads = importlib.import_module(_NAME + ".workloads")
parallel = importlib.import_module(_NAME + ".parallel")
PACKAGE_DIR = _PKG_DIR
