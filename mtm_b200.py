"""Importable name for the package directory ``multimodal-trajectory-modeling_b200`` (hyphens
cannot be spelled in an ``import`` statement).  The package is loaded ONCE, under the name
``mtm_b200``; the hyphenated name is registered as an alias of the same module object."""

import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "multimodal-trajectory-modeling_b200")
_hyphen = "multimodal-trajectory-modeling_b200"
if _hyphen in sys.modules:
    sys.modules[__name__] = sys.modules[_hyphen]
else:
    _spec = importlib.util.spec_from_file_location(__name__, os.path.join(_dir, "__init__.py"),
                                                   submodule_search_locations=[_dir])
    _mod = importlib.util.module_from_spec(_spec)
    sys.modules[__name__] = _mod
    sys.modules[_hyphen] = _mod
    _spec.loader.exec_module(_mod)
