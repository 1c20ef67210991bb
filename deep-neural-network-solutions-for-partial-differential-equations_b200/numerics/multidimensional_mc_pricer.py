"""Entry point with the reference's module path numerics/multidimensional_mc_pricer.py."""
from ..mc_pricer import (AnalyticalBlackScholes, BasketOption, BlackScholesModel, CorrelationMatrix,
                         MonteCarloPricer)

__all__ = ["CorrelationMatrix", "BlackScholesModel", "BasketOption", "MonteCarloPricer", "AnalyticalBlackScholes"]
