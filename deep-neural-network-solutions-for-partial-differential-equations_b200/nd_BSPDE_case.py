"""Names of the reference's nd_BSPDE_case.py hot-path classes."""
from .drivers import PredictionGenerator, TrainingPhases
from .fbsnn import FBSNN
from .networks import Naisnet, Sine
from .problems import CallOptionND as CallOption

__all__ = ["TrainingPhases", "PredictionGenerator", "Sine", "Naisnet", "FBSNN", "CallOption"]
