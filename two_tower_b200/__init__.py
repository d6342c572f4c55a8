"""Import shim: the product package lives in ``two-tower-amazon-recommender_b200/`` (a directory
name Python cannot import); this module loads it under the name ``two_tower_b200``."""
import importlib.util as _u
import sys as _sys
from pathlib import Path as _Path

_pkg_dir = _Path(__file__).resolve().parent.parent / "two-tower-amazon-recommender_b200"
_spec = _u.spec_from_file_location(__name__, _pkg_dir / "__init__.py",
                                   submodule_search_locations=[str(_pkg_dir)])
_mod = _u.module_from_spec(_spec)
_sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
