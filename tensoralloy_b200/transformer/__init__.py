from tensoralloy_b200.transformer.universal import UniversalTransformer
from tensoralloy_b200.transformer.vap import VirtualAtomMap

__all__ = ["UniversalTransformer", "VirtualAtomMap"]
