"""Importable alias of the package directory ``image-search-engine-for-historical-research_b200/``
(a hyphenated name cannot appear in an ``import`` statement):

    from xs_b200 import matching_L2, KNN, rank_ip, ExactIndex
"""
import importlib as _importlib
import os as _os
import sys as _sys

_root = _os.path.dirname(_os.path.abspath(__file__))
if _root not in _sys.path:
    _sys.path.insert(0, _root)
_pkg = _importlib.import_module("image-search-engine-for-historical-research_b200")
globals().update({name: getattr(_pkg, name) for name in _pkg.__all__})
__all__ = list(_pkg.__all__)
# `from xs_b200.diffusion import Diffusion` etc.: alias the already-loaded submodules (no second copy is imported)
for _name, _mod in list(_sys.modules.items()):
    if _name.startswith(_pkg.__name__ + "."):
        _sys.modules[__name__ + _name[len(_pkg.__name__):]] = _mod
