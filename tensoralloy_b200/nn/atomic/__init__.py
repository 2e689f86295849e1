from tensoralloy_b200.nn.atomic.atomic import AtomicNN
from tensoralloy_b200.nn.atomic.finite_temperature import (BeNN,
                                                           TemperatureDependentAtomicNN)
from tensoralloy_b200.nn.atomic.grap import GenericRadialAtomicPotential
from tensoralloy_b200.nn.atomic.sf import SymmetryFunction

__all__ = ["AtomicNN", "SymmetryFunction", "GenericRadialAtomicPotential",
           "TemperatureDependentAtomicNN", "BeNN"]
