"""Importable alias of the package directory ``channel-estimation_b200`` (hyphenated name)."""
import importlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
_pkg = importlib.import_module("channel-estimation_b200")
sys.modules[__name__] = _pkg
