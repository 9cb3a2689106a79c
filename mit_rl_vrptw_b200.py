"""Importable alias of the package directory ``mit_rl-vrptw_b200`` (a hyphen is not a valid identifier)."""
import importlib
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
if _here not in sys.path:
    sys.path.insert(0, _here)
_pkg = importlib.import_module("mit_rl-vrptw_b200")
sys.modules[__name__] = _pkg
