from tensoralloy_b200.nn.atomic.atomic import AtomicNN
from tensoralloy_b200.nn.atomic.grap import GenericRadialAtomicPotential
from tensoralloy_b200.nn.atomic.sf import SymmetryFunction

__all__ = ["AtomicNN", "SymmetryFunction", "GenericRadialAtomicPotential"]
