"""Names of heston_dnnpde.py (HestonFBSNN :519-699, TrainingPhases :955-975) on the fused sm_100a kernels."""
from .drivers import TrainingPhases
from .networks import Naisnet, Sine
from .problems import HestonFBSNN

__all__ = ["HestonFBSNN", "TrainingPhases", "Sine", "Naisnet"]
