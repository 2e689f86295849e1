from tensoralloy_b200.transformer.universal import (BatchUniversalTransformer,
                                                    UniversalTransformer)
from tensoralloy_b200.transformer.vap import VirtualAtomMap

__all__ = ["UniversalTransformer", "BatchUniversalTransformer", "VirtualAtomMap"]
