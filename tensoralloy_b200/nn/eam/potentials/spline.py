"""
Tabulated potentials: natural cubic splines over setfl / ADP tables, evaluated on
the GPU (TAB_FN_SPLINE, csrc/potentials.cuh).  This is the component the
reference references but does not ship (`tensoralloy.extension.interp.cubic.
CubicInterpolator`) together with the `spline@<file>`
potential names of train/training.py:260.

    EamAlloyNN(elements, custom_potentials="spline@/path/to/file.eam.alloy")
    AdpNN(elements, custom_potentials="spline@/path/to/file.adp")
"""
import numpy as np

from tensoralloy_b200 import _lib
from tensoralloy_b200.io.lammps import read_adp_setfl, read_eam_alloy_setfl
from tensoralloy_b200.utils import get_elements_from_kbody_term


class SplinePotential:
    name = 'spline'

    def __init__(self, filename, is_adp=None):
        if is_adp is None:
            is_adp = str(filename).endswith('.adp')
        self.filename = filename
        self.setfl = read_adp_setfl(filename) if is_adp else read_eam_alloy_setfl(filename)
        self._pool = []           # list of [n-1, 4] coefficient blocks
        self._offsets = {}        # (kind, key) -> interval offset
        self._n_intervals = 0
        self.params = {}

    def set_param(self, *args):   # tables have no scalar parameters
        pass

    def _entry(self, kind, key, spline):
        tag = (kind, key)
        if tag not in self._offsets:
            c = spline.coefficients()
            self._offsets[tag] = (self._n_intervals, len(c), spline.x[0],
                                  spline.x[1] - spline.x[0])
            self._pool.append(c)
            self._n_intervals += len(c)
        off, n, x0, dx = self._offsets[tag]
        fn = _lib.make_fn(_lib.FN_SPLINE, [x0, 1.0 / dx, n], aux=off)
        fn._owner = self          # lets EamNN rebase offsets when pools are merged
        return fn

    def pool(self):
        return np.concatenate(self._pool) if self._pool else np.zeros((0, 4))

    def _pair(self, table, kbody_term):
        a, b = get_elements_from_kbody_term(kbody_term)
        for key in (f'{a}{b}', f'{b}{a}'):
            if key in table:
                return key, table[key]
        raise KeyError(f"{self.filename}: no table for {kbody_term}")

    def rho(self, element_or_term):
        el = get_elements_from_kbody_term(element_or_term)[-1]
        return self._entry('rho', el, self.setfl.rho[el])

    def embed(self, element):
        return self._entry('embed', element, self.setfl.embed[element])

    def phi(self, kbody_term):
        key, sp = self._pair(self.setfl.phi, kbody_term)
        return self._entry('phi', key, sp)

    def dipole(self, kbody_term):
        key, sp = self._pair(self.setfl.dipole, kbody_term)
        return self._entry('dipole', key, sp)

    def quadrupole(self, kbody_term):
        key, sp = self._pair(self.setfl.quadrupole, kbody_term)
        return self._entry('quadrupole', key, sp)
