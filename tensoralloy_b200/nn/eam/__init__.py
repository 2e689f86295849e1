from tensoralloy_b200.nn.eam.alloy import EamAlloyNN

__all__ = ["EamAlloyNN"]
