"""Names of the reference's nd_BSPDE_case.py hot-path classes."""
from .fbsnn import FBSNN
from .networks import Naisnet, Sine
from .problems import CallOptionND as CallOption

__all__ = ["Sine", "Naisnet", "FBSNN", "CallOption"]
