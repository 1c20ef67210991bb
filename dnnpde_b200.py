"""Importable alias of the package directory `deep-neural-network-solutions-for-partial-differential-equations_b200`
(whose name contains hyphens): `import dnnpde_b200 as pde; pde.BlackScholesBarenblatt(...)`."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_REAL = "deep-neural-network-solutions-for-partial-differential-equations_b200"
_pkg = importlib.import_module(_REAL)
for _sub in ("DeepBSDE", "with_corr_high_dimension_pde", "hjb_implement", "nd_BSPDE_case", "bspde_1d_case",
             "numerics", "numerics.multidimensional_mc_pricer", "heston_dnnpde", "basket_pricer"):
    importlib.import_module(_REAL + "." + _sub)
# one module object per submodule, reachable under both names (no duplicate class objects)
for _name, _mod in list(sys.modules.items()):
    if _name.startswith(_REAL + "."):
        sys.modules[__name__ + _name[len(_REAL):]] = _mod
sys.modules[__name__] = _pkg
