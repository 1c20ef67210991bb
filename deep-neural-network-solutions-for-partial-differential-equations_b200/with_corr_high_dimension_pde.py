"""Names of the reference's with_corr_high_dimension_pde.py hot-path classes."""
from .drivers import PredictionGenerator, TrainingPhases
from .fbsnn import FBSNN
from .networks import Naisnet, Sine
from .problems import BasketCallOption as CallOption
from .problems import BSPDETestCase

__all__ = ["TrainingPhases", "PredictionGenerator", "Sine", "Naisnet", "FBSNN", "CallOption", "BSPDETestCase"]
