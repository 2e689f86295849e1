"""
tensoralloy_b200 -- B200 (sm_100a) implementation of the TensorAlloy energy /
force / virial hot path behind the reference's operator surface.

Layout (only what the path needs):
  csrc/          CUDA kernels + the C ABI (include/tab200.h) -> libtab200.so
  _lib.py        ctypes binding (no fallback: raises if the library is missing)
  calculator.py  TensorAlloyCalculator          (reference: calculator.py)
  transformer/   UniversalTransformer, VirtualAtomMap (reference: transformer/)
  neighbor.py    find_neighbor_size_of_atoms    (reference: neighbor.py)
  nn/            EamAlloyNN / EamFsNN / AdpNN / AtomicNN  (reference: nn/)
  io/            frozen-.pb / setfl readers     (reference: io/lammps.py, nn/basic.py:1017)
"""
__version__ = '0.1.0'
