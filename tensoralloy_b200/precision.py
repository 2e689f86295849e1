"""
Global floating-point precision, mirror of the reference's
tensoralloy/precision.py:21-134: 'high' = float64 (eps 1e-14), 'medium' = float32
(eps 1e-8).  eps is the constant added under the square root of r_ij
(transformer/universal.py:470-473) and to the MSEs of the losses.
"""
import enum
from contextlib import contextmanager

import numpy as np


class Precision(enum.Enum):
    medium = 0
    high = 1


class DType:
    def __init__(self, np_dtype, eps, name, tab_precision):
        self.as_numpy_dtype = np_dtype
        self.eps = eps
        self.name = name
        self.tab_precision = tab_precision    # TAB_PRECISION_* of include/tab200.h

    def __repr__(self):
        return f"<DType {self.name}>"


float64 = DType(np.float64, 1e-14, 'float64', 0)
float32 = DType(np.float32, 1e-8, 'float32', 1)

_floating_point_precision = None


def set_float_precision(precision=Precision.medium):
    """precision.py:68-88 (the reference's default is 'medium')."""
    global _floating_point_precision
    if isinstance(precision, str):
        precision = Precision[precision]
    if not isinstance(precision, Precision):
        raise ValueError(f"Unknown precision: {precision}")
    _floating_point_precision = precision


def get_float_precision():
    global _floating_point_precision
    if _floating_point_precision is None:
        set_float_precision()
    return _floating_point_precision


def get_float_dtype():
    return float32 if get_float_precision() == Precision.medium else float64


@contextmanager
def precision_scope(precision):
    """precision.py:36-65."""
    if isinstance(precision, str):
        precision = Precision[precision]
    prev = get_float_precision()
    set_float_precision(precision)
    try:
        yield
    finally:
        set_float_precision(prev)
