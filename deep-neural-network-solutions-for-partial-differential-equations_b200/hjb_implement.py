"""Names of the reference's hjb_implement.py hot-path classes."""
from .drivers import PredictionGenerator, TrainingPhases
from .fbsnn import FBSNN
from .networks import Naisnet, Sine
from .problems import BasketCallOption as _BasketCallOption
from .problems import HamiltonJacobiBellman


class CallOption(_BasketCallOption):
    """hjb_implement.py:543-586 -- same callables as the with_corr variant, Mm-style N-schedule (:403-406)."""
    _schedule_kind = "mm"

__all__ = ["TrainingPhases", "PredictionGenerator", "Sine", "Naisnet", "FBSNN", "CallOption", "HamiltonJacobiBellman"]
