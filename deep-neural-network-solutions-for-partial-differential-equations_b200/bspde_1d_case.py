"""Names of the reference's 1d_BSPDE_case.py hot-path classes (a module name cannot start with a digit)."""
from .drivers import PredictionGenerator, TrainingPhases
from .fbsnn import FBSNN
from .networks import Naisnet, Sine
from .problems import CallOption1D as CallOption

__all__ = ["TrainingPhases", "PredictionGenerator", "Sine", "Naisnet", "FBSNN", "CallOption"]
