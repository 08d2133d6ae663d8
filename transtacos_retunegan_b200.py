"""Importable alias of the package directory ``transtacos-retunegan_b200/`` (hyphens are not valid in
``import`` statements).  ``import transtacos_retunegan_b200 as sb`` returns the package itself."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("transtacos-retunegan_b200")
sys.modules[__name__] = _pkg
