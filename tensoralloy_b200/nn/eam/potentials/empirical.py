"""
Further analytic potentials of the reference -> libtab200 function tables.
Mirrors (parameter names, defaults, section layout):
  AgSutton90  nn/eam/potentials/sutton90.py:18-121
  AgrawalBe   nn/eam/potentials/agrawal.py:35-170
  RWGrimes    nn/eam/potentials/grimmes.py:20-126
  MishinH     nn/eam/potentials/mishin.py:20-315 (embed / dipole / quadrupole; the
              reference's MishinH.rho/.phi read undefined variables -- SURVEY.md 0.1)
The arithmetic is evaluated on the GPU (csrc/potentials.cuh).
"""
import copy

from tensoralloy_b200 import _lib
from tensoralloy_b200.precision import get_float_dtype
from tensoralloy_b200.utils import get_elements_from_kbody_term


class _Potential:
    name = 'base'

    def __init__(self, params=None):
        self._params = copy.deepcopy(params) if params is not None else self.defaults

    @property
    def params(self):
        return self._params

    def set_param(self, section, key, value):
        self._params.setdefault(section, {})[key] = float(value)

    def _pair_section(self, kbody_term):
        a, b = get_elements_from_kbody_term(kbody_term)
        for key in (f'{a}{b}', f'{b}{a}'):
            if key in self._params:
                return self._params[key]
        raise KeyError(f"{self.name}: no parameters for {kbody_term}")


class AgSutton90(_Potential):
    name = 'sutton90'

    @property
    def defaults(self):
        return {'Ag': {'a': 2.928323832}, 'AgAg': {'b': 2.485883762}}

    def rho(self, element_or_term):
        el = get_elements_from_kbody_term(element_or_term)[-1]
        return _lib.make_fn(_lib.FN_SUTTON_RHO, [self._params[el]['a']])

    def phi(self, kbody_term):
        return _lib.make_fn(_lib.FN_SUTTON_PHI, [self._pair_section(kbody_term)['b']])

    def embed(self, element):
        return _lib.make_fn(_lib.FN_SQRT_EMBED, [1.0])


class AgrawalBe(_Potential):
    name = 'Be/1'

    @property
    def defaults(self):
        return {"Be": {"A": 1.597, "B": 9.49713, "D": 0.41246, "alpha": 0.36324,
                       "re": 2.29, "F0": -2.0393, "F1": 12.6178,
                       "beta": 0.18752, "gamma": -2.28827, "m": 10, "rc": 5.0}}

    def rho(self, element_or_term):
        el = get_elements_from_kbody_term(element_or_term)[-1]
        p = self._params[el]
        return _lib.make_fn(_lib.FN_AGRAWAL_RHO,
                            [p['A'], p['B'], p['re'], p['rc'], p['m']])

    def phi(self, kbody_term):
        el = get_elements_from_kbody_term(kbody_term)[0]
        p = self._params[el]
        return _lib.make_fn(_lib.FN_AGRAWAL_PHI,
                            [p['D'], p['alpha'], p['re'], p['rc'], p['m']])

    def embed(self, element):
        p = self._params[element]
        return _lib.make_fn(_lib.FN_AGRAWAL_EMBED,
                            [p['F0'], p['F1'], p['beta'], p['gamma']])


class RWGrimes(_Potential):
    name = 'grimes'

    @property
    def defaults(self):
        return {'PuPu': {'A': 18600.0, 'rho': 0.2637, 'C': 0.0, 'D': 0.70185,
                         'gamma': 1.98008, 'r0': 2.34591},
                'Pu': {'G': 2.168, 'n': 3980.058}}

    def rho(self, element_or_term):
        el = get_elements_from_kbody_term(element_or_term)[-1]
        return _lib.make_fn(_lib.FN_GRIMES_RHO, [self._params[el]['n']])

    def phi(self, kbody_term):
        p = self._pair_section(kbody_term)
        return _lib.make_fn(_lib.FN_GRIMES_PHI, [p['A'], p['rho'], p['C'], p['D'],
                                                 p['gamma'], p['r0']])

    def embed(self, element):
        return _lib.make_fn(_lib.FN_SQRT_EMBED, [self._params[element]['G']])


class MishinH(_Potential):
    name = 'mishinh'

    @property
    def defaults(self):
        params = {
            "Mo": {"s1": -2.00695289e-01, "s2": -3.12178751e-04, "s3": 7.86343222e-05,
                   "s4": 5.29721645e+00, "s5": 3.79481951e-02, "s6": 1.11800974e+02,
                   "s7": 4.05948858e+00},
            "Al": {"s1": -3.72848864e-01, "s2": 6.52035828e-03, "s3": 9.71742655e-05,
                   "s4": 7.64264116e+00, "s5": 6.88604789e-02, "s6": 1.55694016e+01,
                   "s7": 5.38646368e+00},
            "H": {"s1": 8.08612, "s2": 1.46294e-2, "s3": -6.86143e-3, "s4": 3.19616,
                  "s5": 1.17247e-1, "s6": 50, "s7": 15e5},
            "NiNi": {"d1": 4.4657e-3, "d2": -1.3702e0, "d3": -0.9611e-1,
                     "q1": 6.4502e0, "q2": 0.2608e-1, "q3": -6.0208e0,
                     "h": 3.323, "rc": 5.168},
            "FeFe": {"d1": 1.9135e-1, "d2": -1.0796e0, "d3": -0.8928e-1,
                     "q1": -5.8954e-2, "q2": -1.3872e0, "q3": 2.4790e0,
                     "h": 6.202, "rc": 5.055},
        }
        params['MoMo'] = dict(params['NiNi'])
        params['MoNi'] = dict(params['NiNi'])
        params['BeBe'] = dict(params['MoMo'])
        return params

    def rho(self, element_or_term):
        raise NotImplementedError(
            "MishinH.rho reads undefined variables in the reference (SURVEY.md 0.1)")

    phi = rho

    def embed(self, element):
        p = self._params[element]
        eps = get_float_dtype().eps
        return _lib.make_fn(_lib.FN_MISHIN_EMBED,
                            [p[f's{k}'] for k in range(1, 8)] + [eps])

    def dipole(self, kbody_term):
        p = self._pair_section(kbody_term)
        return _lib.make_fn(_lib.FN_MISHIN_POLAR,
                            [p['d1'], p['d2'], p['d3'], p['rc'], p['h']])

    def quadrupole(self, kbody_term):
        p = self._pair_section(kbody_term)
        return _lib.make_fn(_lib.FN_MISHIN_POLAR,
                            [p['q1'], p['q2'], p['q3'], p['rc'], p['h']])
