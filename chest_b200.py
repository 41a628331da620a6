"""Importable alias of the package directory ``channel-estimation_b200`` (hyphenated name):
``import chest_b200`` and ``from chest_b200.modulation import FBMC`` resolve to the very same
module objects as ``importlib.import_module("channel-estimation_b200")``."""
import importlib
import os
import sys

_REAL = "channel-estimation_b200"
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
_pkg = importlib.import_module(_REAL)
for _name, _mod in list(sys.modules.items()):
    if _name.startswith(_REAL + "."):
        sys.modules[__name__ + _name[len(_REAL):]] = _mod
sys.modules[__name__] = _pkg
