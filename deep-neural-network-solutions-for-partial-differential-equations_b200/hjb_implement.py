"""Names of the reference's hjb_implement.py hot-path classes."""
from .drivers import PredictionGenerator, TrainingPhases
from .fbsnn import FBSNN
from .networks import Naisnet, Sine
from .problems import BasketCallOption as _BasketCallOption
from .problems import HamiltonJacobiBellman, hjb_u_exact


class CallOption(_BasketCallOption):
    """hjb_implement.py:543-586 -- same callables as the with_corr variant, Mm-style N-schedule (:403-406)."""
    _schedule_kind = "mm"

u_exact = hjb_u_exact   # the driver's Cole-Hopf comparator closure (hjb_implement.py:1088-1094), on the device

__all__ = ["TrainingPhases", "PredictionGenerator", "Sine", "Naisnet", "FBSNN", "CallOption", "HamiltonJacobiBellman",
           "hjb_u_exact", "u_exact"]
