from tensoralloy_b200.nn.eam.adp import AdpNN
from tensoralloy_b200.nn.eam.alloy import EamAlloyNN
from tensoralloy_b200.nn.eam.fs import EamFsNN

__all__ = ["EamAlloyNN", "EamFsNN", "AdpNN"]
