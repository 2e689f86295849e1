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

    # {section: [parameter, ...]} never trained by the reference (potentials.py:76-79);
    # none of the potentials of this module declares any
    fixed_parameters = property(lambda self: {})

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


class AlFeMsah11(_Potential):
    """nn/eam/potentials/msah11.py:28-424 -- Mendelev et al. (2011) Al-Fe Finnis-Sinclair
    potential: fixed constants, no trainable parameters.  rho is a function of the ordered
    pair (EamFsNN); the cross density and the cross pair function are shared by Al-Fe and
    Fe-Al.  The pair function's constants travel in the model's coefficient pool
    (TAB_FN_MSAH_PHI), the densities and embeddings in the table entries themselves."""
    name = 'msah11'

    _PHI = {
        'AlAl': dict(
            first=(1e-8, 1.60, [2433.5591473227, 0.1818, -22.713109144730, 0.5099,
                                -6.6883008584622, 0.2802, -2.8597223982536, 0.02817,
                                -1.4309258761180]),
            second=(1.6, 2.25, [6.0801330531321, -2.3092752322555, 0.042696494305190,
                                -0.07952189194038]),
            polys=[(2.25, 3.2, [17.222548257633, -13.838795389103, 26.724085544227,
                                -4.8730831082596, 0.26111775221382], [4, 5, 6, 7, 8]),
                   (2.25, 4.8, [-1.8864362756631, 2.4323070821980, -4.0022263154653,
                                1.3937173764119, -0.31993486318965], [4, 5, 6, 7, 8]),
                   (2.25, 6.5, [0.30601966016455, -0.63945082587403, 0.54057725028875,
                                -0.21210673993915, 0.03201431888287], [4, 5, 6, 7, 8])]),
        'FeFe': dict(
            first=(1e-8, 1.0, [9734.2365892908, 0.1818, -28.616724320005, 0.5099,
                               -8.4267310396064, 0.2802, -3.6030244464156, 0.02817,
                               -1.8028536321603]),
            second=(1.0, 2.05, [7.4122709384068, -0.64180690713367, -2.6043547961722,
                                0.62625393931230]),
            polys=[(2.05, hc, [a], [3]) for hc, a in zip(
                [2.2, 2.3, 2.4, 2.5, 2.6, 2.7, 2.8, 3.0, 3.3, 3.7, 4.2, 4.7, 5.3],
                [-27.444805994228, 15.738054058489, 2.2077118733936, -2.4989799053251,
                 4.2099676494795, -0.77361294129713, 0.80656414937789, -2.3194358924605,
                 2.6577406128280, -1.0260416933564, 0.35018615891957, -0.058531821042271,
                 -0.0030458824556234])]),
        'AlFe': dict(
            first=(1e-8, 1.2, [4867.1182946454, 0.1818, -25.834107666296, 0.5099,
                               -7.6073373918597, 0.2802, -3.2526756183596, 0.02817,
                               -1.6275487829767]),
            second=(1.2, 2.2, [6.6167846784367, -1.5208197629514, -0.73055022396300,
                               -0.03879272494264]),
            polys=[(2.2, 3.2, [-4.148701943924, 5.6697481153271, -1.7835153896441,
                               -3.3886912738827, 1.9720627768230], [4, 5, 6, 7, 8]),
                   (2.2, 6.2, [0.094200713038410, -0.16163849208165, 0.10154590006100,
                               -0.027624717063181, 0.0027505576632627], [4, 5, 6, 7, 8])]),
    }
    _RHO = {
        'AlAl': (4, [0.00019850823042883, 0.10046665347629, 1.0054338881951E-01,
                     0.099104582963213, 0.090086286376778, 0.0073022698419468,
                     0.014583614223199, -0.0010327381407070, 0.0073219994475288,
                     0.0095726042919017],
                 [2.5, 2.6, 2.7, 2.8, 3.0, 3.4, 4.2, 4.8, 5.6, 6.5]),
        'FeFe': (3, [11.686859407970, -0.014710740098830, 0.47193527075943],
                 [2.4, 3.2, 4.2]),
        'AlFe': (4, [0.010015421408039, 0.0098878643929526, 0.0098070326434207,
                     0.0084594444746494, 0.0038057610928282, -0.0014091094540309,
                     0.0074410802804324], [2.4, 2.5, 2.6, 2.8, 3.1, 5.0, 6.2]),
    }

    def __init__(self, params=None):
        super().__init__(params)
        self._blocks, self._offsets, self._n4 = [], {}, 0

    @property
    def defaults(self):
        return {"Al": {}, "Fe": {}}

    @staticmethod
    def _key(kbody_term):
        a, b = get_elements_from_kbody_term(kbody_term)
        if {a, b} - {'Al', 'Fe'}:
            raise KeyError(f"msah11: no parameters for {kbody_term}")
        return a + b if a == b else 'AlFe'

    def rho(self, kbody_term):
        els = get_elements_from_kbody_term(kbody_term)
        key = self._key(kbody_term) if len(els) == 2 else els[0] + els[0]
        order, factors, cutoffs = self._RHO[key]
        vals = [float(order), float(len(factors))]
        for f, rc in zip(factors, cutoffs):
            vals += [f, rc]
        return _lib.make_fn(_lib.FN_POWCUT_RHO, vals)

    def phi(self, kbody_term):
        import numpy as np
        key = self._key(kbody_term)
        if key not in self._offsets:
            d = self._PHI[key]
            flat = [float(len(d['polys']))]
            flat += [d['first'][0], d['first'][1]] + list(d['first'][2])
            flat += [d['second'][0], d['second'][1]] + list(d['second'][2])
            for lo, hi, coef, orders in d['polys']:
                flat += [lo, hi, float(len(coef))]
                for a, n in zip(coef, orders):
                    flat += [a, float(n)]
            flat = np.asarray(flat, dtype=np.float64)
            flat = np.concatenate([flat, np.zeros((-len(flat)) % 4)]).reshape(-1, 4)
            self._offsets[key] = self._n4
            self._blocks.append(flat)
            self._n4 += len(flat)
        entry = _lib.make_fn(_lib.FN_MSAH_PHI, [], aux=self._offsets[key])
        entry._owner = self
        return entry

    def embed(self, element):
        if element == 'Al':
            return _lib.make_fn(_lib.FN_MSAH_EMBED_AL,
                                [0.000093283590195398, 0.0023491751192724])
        if element == 'Fe':
            return _lib.make_fn(_lib.FN_MSAH_EMBED_FE,
                                [0.00067314115586063, 0.000000076514905604792])
        raise KeyError(f"msah11: no embedding function for {element}")

    def pool(self):
        import numpy as np
        return np.concatenate(self._blocks) if self._blocks else np.zeros((0, 4))
