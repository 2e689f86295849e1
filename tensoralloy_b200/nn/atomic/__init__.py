from tensoralloy_b200.nn.atomic.atomic import AtomicNN
from tensoralloy_b200.nn.atomic.sf import SymmetryFunction

__all__ = ["AtomicNN", "SymmetryFunction"]
