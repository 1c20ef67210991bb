"""B200-native (sm_100a) FBSNN training step and basket Monte-Carlo pricer behind the reference's Python surface.

The directory name is not a Python identifier; import it with
    importlib.import_module("deep-neural-network-solutions-for-partial-differential-equations_b200")
or through the `dnnpde_b200` alias module at the repository root.
"""
from . import _lib, parallel, spec
from . import basket_pricer
from .comparators import BasicOptionPriceCalculator, BasketOptionPriceCalculator
from .drivers import PredictionGenerator, TrainingPhases
from .fbsnn import FBSNN
from .mc_pricer import (AnalyticalBlackScholes, BasketOption, BlackScholesModel, CorrelationMatrix,
                        MonteCarloPricer)
from .networks import Naisnet, Sine
from .problems import (BasketCallOption, BlackScholesBarenblatt, BSPDETestCase, CallOption1D, CallOptionND,
                       HamiltonJacobiBellman, HestonFBSNN, hjb_u_exact, u_exact)

build = _lib.build

__all__ = ["FBSNN", "Sine", "Naisnet", "BlackScholesBarenblatt", "BasketCallOption", "BSPDETestCase",
           "CallOption1D", "CallOptionND", "HamiltonJacobiBellman", "HestonFBSNN", "u_exact", "hjb_u_exact", "basket_pricer", "CorrelationMatrix",
           "BlackScholesModel", "BasketOption", "MonteCarloPricer", "AnalyticalBlackScholes", "build",
           "TrainingPhases", "PredictionGenerator", "BasketOptionPriceCalculator", "BasicOptionPriceCalculator",
           "parallel", "spec"]
