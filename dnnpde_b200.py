"""Importable alias of the package directory `deep-neural-network-solutions-for-partial-differential-equations_b200`
(whose name contains hyphens): `import dnnpde_b200 as pde; pde.BlackScholesBarenblatt(...)`."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("deep-neural-network-solutions-for-partial-differential-equations_b200")
sys.modules[__name__] = _pkg
