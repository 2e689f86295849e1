from tensoralloy_b200.analysis.phonon import PhononCalculator

__all__ = ["PhononCalculator"]
