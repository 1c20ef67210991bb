"""Names of the reference's DeepBSDE.py (Sine, FBSNN, BlackScholesBarenblatt, u_exact)."""
from .fbsnn import FBSNN
from .networks import Sine
from .problems import BlackScholesBarenblatt, u_exact

__all__ = ["Sine", "FBSNN", "BlackScholesBarenblatt", "u_exact"]
