"""
oracle/potentials.py -- TEST INFRASTRUCTURE (see oracle/__init__.py).

torch (CPU, float64 by default) restatement of the analytic EAM potential
functions of the reference, `tensoralloy/nn/eam/potentials/*`.  `torch.pow` with
a floating exponent stands in for `tf.pow` (`safe_pow` without the env switch,
extension/grad_ops.py:16-74).  Every function returns a tensor shaped like its
input; autograd provides the derivatives exactly as TF autograd does in the
reference.
"""
import json
from pathlib import Path

import torch

_DATA = Path(__file__).resolve().parent / 'data'


def _elements_of(term):
    """utils.py:210-234 get_elements_from_kbody_term."""
    out = []
    for ch in term:
        if ch.isupper():
            out.append(ch)
        else:
            out[-1] += ch
    return out


# --------------------------------------------------------------------------
# generic.py
# --------------------------------------------------------------------------

def density_exp(r, a, b, re):
    """generic.py:87-99: a * exp(-b (r/re - 1))."""
    return a * torch.exp(-b * (r / re - 1.0))


def zhou_exp(r, a, b, c, re, order=20):
    """generic.py:102-117: a exp(-b (r/re - 1)) / (1 + (r/re - c)^order)."""
    x = r / re
    upper = density_exp(r, a, b, re)
    lower = 1.0 + torch.pow(x - c, torch.tensor(float(order), dtype=r.dtype))
    return upper / lower


def morse(r, d, gamma, r0):
    """generic.py:15-30."""
    gd = gamma * (r - r0)
    return d * (torch.exp(-2.0 * gd) - 2.0 * torch.exp(-gd))


def buckingham(r, A, rho, C, order=6):
    """generic.py:33-49 (C / r^order via div_no_nan)."""
    rs = torch.pow(r, order)
    right = torch.where(rs == 0, torch.zeros_like(rs), C / rs)
    return A * torch.exp(-r / rho) - right


def mishin_cutoff(x):
    """generic.py:52-66: psi(x) = x^4 / (1 + x^4) for x < 0 else 0."""
    ix = torch.relu(-x)
    x4 = torch.pow(ix, 4.0)
    return x4 / (1.0 + x4)


def mishin_polar(x, p1, p2, p3, rc, h):
    """generic.py:69-84."""
    z = (x - rc) / h
    return (p1 * torch.exp(-p2 * x) + p3) * mishin_cutoff(z)


# --------------------------------------------------------------------------
# potential families
# --------------------------------------------------------------------------

class Potential:
    """Interface: rho(r, element_or_term), phi(r, term), embed(rho, element),
    optional dipole(r, term) / quadrupole(r, term)."""
    name = 'base'

    def __init__(self, params=None, dtype=torch.float64):
        self.dtype = dtype
        self.params = params if params is not None else self.defaults()

    def p(self, section, key):
        leaves = getattr(self, 'leaves', None)
        if leaves is not None:
            if (section, key) not in leaves:
                leaves[(section, key)] = torch.tensor(self.params[section][key],
                                                      dtype=self.dtype, requires_grad=True)
            return leaves[(section, key)]
        return torch.tensor(self.params[section][key], dtype=self.dtype)

    def track_parameters(self):
        """From now on every parameter is one shared leaf tensor (requires_grad): the
        reference's shared trainable variables, potentials.py:171-200.  Returns the dict
        (section, key) -> leaf, filled as the functions touch their parameters."""
        self.leaves = {}
        return self.leaves

    def defaults(self):
        raise NotImplementedError


class Zjw04(Potential):
    """zjw04.py:155-413."""
    name = 'zjw04'

    def defaults(self):
        return json.loads((_DATA / 'zjw04.json').read_text())['zjw04']

    def rho(self, r, element):
        # zjw04.py:245-277; for eam/alloy `element` is the NEIGHBOUR element
        # (alloy.py:162-176).
        element = _elements_of(element)[-1]
        P = lambda k: self.p(element, k)
        return zhou_exp(r, a=P('f_eq'), b=P('beta'), c=P('lamda'), re=P('r_eq'))

    def phi(self, r, term):
        # zjw04.py:187-243
        a, b = _elements_of(term)
        if a == b:
            P = lambda k: self.p(a, k)
            return (zhou_exp(r, P('A'), P('alpha'), P('kappa'), P('r_eq')) -
                    zhou_exp(r, P('B'), P('beta'), P('lamda'), P('r_eq')))
        phi_a = self.phi(r, a + a)
        rho_a = self.rho(r, a)
        phi_b = self.phi(r, b + b)
        rho_b = self.rho(r, b)
        return 0.5 * (rho_a / rho_b * phi_b + rho_b / rho_a * phi_a)

    def _branches(self, element):
        P = lambda k: self.p(element, k)
        rho_e = P('rho_e')
        rho_s = P('rho_s')
        rho_n = torch.tensor(0.85, dtype=self.dtype) * rho_e
        rho_0 = torch.tensor(1.15, dtype=self.dtype) * rho_e
        two = torch.tensor(2.0, dtype=self.dtype)
        three = torch.tensor(3.0, dtype=self.dtype)

        def e1(x):
            x1 = x / rho_n - 1.0
            return P('Fn0') + (P('Fn1') * x1 + P('Fn2') * torch.pow(x1, two)
                               + P('Fn3') * torch.pow(x1, three))

        def e2(x):
            x1 = x / rho_e - 1.0
            return P('F0') + (P('F1') * x1 + P('F2') * torch.pow(x1, two)
                              + P('F3') * torch.pow(x1, three))

        def e3(x, eps=0.0):
            xs = x / rho_s + eps
            return P('Fe') * (1.0 - P('eta') * torch.log(xs)) * torch.pow(
                xs, P('eta'))

        return rho_n, rho_0, e1, e2, e3

    def embed(self, rho, element):
        # zjw04.py:279-389: three branches selected by gather/scatter
        rho_n, rho_0, e1, e2, e3 = self._branches(element)
        out = torch.zeros_like(rho)
        m1 = rho < rho_n
        m2 = (rho >= rho_n) & (rho < rho_0)
        m3 = rho >= rho_0
        flat = rho.reshape(-1)
        pieces = torch.zeros_like(flat)
        for m, fn in ((m1, e1), (m2, e2), (m3, e3)):
            idx = torch.nonzero(m.reshape(-1)).reshape(-1)
            if idx.numel():
                pieces = pieces.index_put((idx,), fn(flat[idx]),
                                          accumulate=True)
        return (out.reshape(-1) + pieces).reshape(rho.shape)


class Zjw04xc(Zjw04):
    """zjw04.py:420-550: sigmoid-blended embedding; adds 'Be' = Mo."""
    name = 'zjw04xc'

    def defaults(self):
        params = dict(super().defaults())
        params['Be'] = dict(params['Mo'])
        return params

    def embed(self, rho, element):
        rho_n, rho_0, e1, e2, e3 = self._branches(element)
        y1 = e1(rho)
        y2 = e2(rho)
        y3 = e3(rho, eps=1e-8)
        c1 = torch.sigmoid(2.0 * (rho_n - rho))
        c3 = torch.sigmoid(2.0 * (rho - rho_0))
        c2 = 1.0 - (c1 + c3)
        return c1 * y1 + c2 * y2 + c3 * y3


class Zjw04uxc(Zjw04xc):
    """zjw04.py:553-567 (same arithmetic, different fixed-variable set)."""
    name = 'zjw04uxc'


class Zjw04xcp(Zjw04xc):
    """zjw04.py:570-696: A-B pair has its own zhou_exp parameter set."""
    name = 'zjw04xcp'

    def defaults(self):
        params = dict(super().defaults())
        over = json.loads((_DATA / 'zjw04.json').read_text())[
            'zjw04xcp_overrides']
        for k, v in over.items():
            params[k] = dict(v)
        return params

    def phi(self, r, term):
        a, b = _elements_of(term)
        sec = a if a == b else term
        if sec not in self.params:
            sec = b + a
        P = lambda k: self.p(sec, k)
        return (zhou_exp(r, P('A'), P('alpha'), P('kappa'), P('r_eq')) -
                zhou_exp(r, P('B'), P('beta'), P('lamda'), P('r_eq')))


class AgSutton90(Potential):
    """sutton90.py:18-121."""
    name = 'sutton90'

    def defaults(self):
        return {'Ag': {'a': 2.928323832}, 'AgAg': {'b': 2.485883762}}

    def rho(self, r, element):
        el = _elements_of(element)[-1]
        rinv = torch.where(r == 0, torch.zeros_like(r), 1.0 / r)   # div_no_nan
        return torch.pow(self.p(el, 'a') * rinv, 6.0)

    def phi(self, r, term):
        rinv = torch.where(r == 0, torch.zeros_like(r), 1.0 / r)
        return torch.pow(self.p(term, 'b') * rinv, 12.0)

    def embed(self, rho, element):
        return -torch.sqrt(rho)


def morse_prime(r, d, gamma, r0):
    """agrawal.py:20-32."""
    gd = gamma * (r - r0)
    return (torch.exp(-gd) - torch.exp(-2.0 * gd)) * (d * gamma * 2.0)


class AgrawalBe(Potential):
    """agrawal.py:35-170."""
    name = 'Be/1'

    def defaults(self):
        return {"Be": {"A": 1.597, "B": 9.49713, "D": 0.41246, "alpha": 0.36324,
                       "re": 2.29, "F0": -2.0393, "F1": 12.6178,
                       "beta": 0.18752, "gamma": -2.28827, "m": 10, "rc": 5.0}}

    def rho(self, r, element):
        el = _elements_of(element)[-1]
        P = lambda k: self.p(el, k)
        A, B, re, rc, m = P('A'), P('B'), P('re'), P('rc'), P('m')
        rho0 = A * torch.exp(-B * (r - re))
        rho1 = A * torch.exp(-B * (rc - re))
        drho = -A * B * torch.exp(-B * (rc - re))
        rho2 = rc / m * (1.0 - (r / rc) ** m) * drho
        return rho0 - rho1 + rho2

    def phi(self, r, term):
        el = _elements_of(term)[0]
        P = lambda k: self.p(el, k)
        D, alpha, re, rc, m = P('D'), P('alpha'), P('re'), P('rc'), P('m')
        phi0 = morse(r, D, alpha, re)
        phi1 = -morse(rc, D, alpha, re)
        z = torch.pow(r / rc, m)
        dphi = morse_prime(rc, D, alpha, re)
        return phi0 + phi1 + rc / m * ((1.0 - z) * dphi)

    def embed(self, rho, element):
        P = lambda k: self.p(element, k)
        x = torch.pow(rho, P('beta'))
        y = torch.pow(rho, P('gamma'))
        logrho = torch.log(torch.clamp(rho, min=1e-12))
        return P('F0') * (1.0 - P('beta') * logrho) * x + P('F1') * y


class RWGrimes(Potential):
    """grimmes.py:20-126."""
    name = 'grimes'

    def defaults(self):
        return {'PuPu': {'A': 18600.0, 'rho': 0.2637, 'C': 0.0, 'D': 0.70185,
                         'gamma': 1.98008, 'r0': 2.34591},
                'Pu': {'G': 2.168, 'n': 3980.058}}

    def phi(self, r, term):
        P = lambda k: self.p(term, k)
        return morse(r, P('D'), P('gamma'), P('r0')) + buckingham(
            r, P('A'), P('rho'), P('C'))

    def rho(self, r, element):
        el = _elements_of(element)[-1]
        rs = torch.pow(r, 8)
        left = torch.where(rs == 0, torch.zeros_like(rs), self.p(el, 'n') / rs)
        right = 0.5 + 0.5 * torch.erf(20.0 * (r - 1.0 - 0.5))
        return left * right

    def embed(self, rho, element):
        return -torch.sqrt(rho) * self.p(element, 'G')


class MishinH(Potential):
    """mishin.py:20-315 (embed, dipole, quadrupole; rho/phi are broken upstream)."""
    name = 'mishinh'

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

    def embed(self, rho, element):
        P = lambda k: self.p(element, k)
        eps = 1e-14 if self.dtype == torch.float64 else 1e-8
        rho2 = rho * rho
        rho3 = rho * rho2
        rho4 = rho2 * rho2
        rhos5 = torch.pow(rho + eps, P('s5'))
        a = 1.0 - P('s6') * rho2
        b = 1.0 + P('s7') * rho4
        omega = 1.0 - a / b
        return (P('s1') * rho + P('s2') * rho2 + P('s3') * rho3
                - P('s4') * rhos5) * omega

    def dipole(self, r, term):
        P = lambda k: self.p(term, k)
        return mishin_polar(r, P('d1'), P('d2'), P('d3'), P('rc'), P('h'))

    def quadrupole(self, r, term):
        P = lambda k: self.p(term, k)
        return mishin_polar(r, P('q1'), P('q2'), P('q3'), P('rc'), P('h'))


class SplineTable(Potential):
    """Natural cubic splines over setfl / ADP tables: io/lammps.py:60-221 `Spline`
    (+ the CubicInterpolator the reference imports from the missing
    extension/interp package).  Coefficients from scipy
    CubicSpline(bc_type='natural'); evaluation in torch so autograd differentiates
    the piecewise polynomial.  `tables`: dict kind -> key -> (x0, dx, y)."""
    name = 'spline'

    def __init__(self, tables, dtype=torch.float64):
        self.dtype = dtype
        self.params = {}
        self._c = {}
        from scipy.interpolate import CubicSpline
        import numpy as np
        for kind, group in tables.items():
            for key, (x0, dx, y) in group.items():
                x = x0 + dx * np.arange(len(y))
                cs = CubicSpline(x, np.asarray(y, dtype=np.float64), bc_type='natural')
                self._c[(kind, key)] = (x0, dx, torch.tensor(cs.c.copy(), dtype=dtype))

    def _eval(self, kind, key, x):
        if (kind, key) not in self._c:
            a, b = _elements_of(key) if len(_elements_of(key)) == 2 else (key, None)
            key = b + a
        x0, dx, c = self._c[(kind, key)]
        n = c.shape[1]
        k = torch.clamp(torch.floor((x.detach() - x0) / dx).long(), 0, n - 1)
        t = x - (x0 + k.to(x.dtype) * dx)
        return ((c[0, k] * t + c[1, k]) * t + c[2, k]) * t + c[3, k]

    def rho(self, r, element):
        return self._eval('rho', _elements_of(element)[-1], r)

    def phi(self, r, term):
        return self._eval('phi', term, r)

    def embed(self, rho, element):
        return self._eval('embed', element, rho)

    def dipole(self, r, term):
        return self._eval('dipole', term, r)

    def quadrupole(self, r, term):
        return self._eval('quadrupole', term, r)


class AlFeMsah11(Potential):
    """msah11.py:28-424 -- Mendelev et al. (2011) Al-Fe Finnis-Sinclair potential; no
    trainable parameters.  phi: screened-Coulomb head, exp(cubic) bridge, sums of
    a (r_k - r)^n H(r_k - r) H(r - r_lo) tails; rho: sum f_i max(r_i - r, 0)^order."""
    name = 'msah11'

    PHI = {
        'AlAl': dict(
            first=(1e-8, 1.60, [2433.5591473227, 0.1818, -22.713109144730, 0.5099,
                                -6.6883008584622, 0.2802, -2.8597223982536, 0.02817,
                                -1.4309258761180]),
            second=(1.6, 2.25, [6.0801330531321, -2.3092752322555, 0.042696494305190,
                                -0.07952189194038]),
            polys=[(2.25, 3.2, [(17.222548257633, 4), (-13.838795389103, 5),
                                (26.724085544227, 6), (-4.8730831082596, 7),
                                (0.26111775221382, 8)]),
                   (2.25, 4.8, [(-1.8864362756631, 4), (2.4323070821980, 5),
                                (-4.0022263154653, 6), (1.3937173764119, 7),
                                (-0.31993486318965, 8)]),
                   (2.25, 6.5, [(0.30601966016455, 4), (-0.63945082587403, 5),
                                (0.54057725028875, 6), (-0.21210673993915, 7),
                                (0.03201431888287, 8)])]),
        'FeFe': dict(
            first=(1e-8, 1.0, [9734.2365892908, 0.1818, -28.616724320005, 0.5099,
                               -8.4267310396064, 0.2802, -3.6030244464156, 0.02817,
                               -1.8028536321603]),
            second=(1.0, 2.05, [7.4122709384068, -0.64180690713367, -2.6043547961722,
                                0.62625393931230]),
            polys=[(2.05, hc, [(a, 3)]) for hc, a in zip(
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
            polys=[(2.2, 3.2, [(-4.148701943924, 4), (5.6697481153271, 5),
                               (-1.7835153896441, 6), (-3.3886912738827, 7),
                               (1.9720627768230, 8)]),
                   (2.2, 6.2, [(0.094200713038410, 4), (-0.16163849208165, 5),
                               (0.10154590006100, 6), (-0.027624717063181, 7),
                               (0.0027505576632627, 8)])]),
    }
    RHO = {
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

    def defaults(self):
        return {'Al': {}, 'Fe': {}}

    @staticmethod
    def _key(term):
        a, b = _elements_of(term)
        return a + b if a == b else 'AlFe'

    def phi(self, r, term):
        d = self.PHI[self._key(term)]
        zero = torch.zeros_like(r)
        lo, hi, c = d['first']
        m = (r >= lo) & (r < hi)
        x = torch.where(m, r, torch.ones_like(r))
        y = (c[0] / x) * sum(c[1 + 2 * i] * torch.exp(c[2 + 2 * i] * x) for i in range(4))
        out = torch.where(m, y, zero)
        lo, hi, c = d['second']
        m = (r >= lo) & (r < hi)
        y = torch.exp(c[0] + c[1] * r + c[2] * r ** 2 + c[3] * r ** 3)
        out = out + torch.where(m, y, zero)
        for lo, hi, terms in d['polys']:
            m = (r >= lo) & (r < hi)
            x = torch.where(m, hi - r, zero)
            out = out + sum(a * x ** n for a, n in terms)
        return out

    def rho(self, r, term):
        els = _elements_of(term)
        key = self._key(term) if len(els) == 2 else els[0] + els[0]
        order, factors, cutoffs = self.RHO[key]
        return sum(f * torch.clamp(rc - r, min=0.0) ** order
                   for f, rc in zip(factors, cutoffs))

    def embed(self, rho, element):
        if element == 'Al':
            m = rho >= 1e-12
            x = torch.where(m, rho, torch.ones_like(rho))
            y = -torch.sqrt(x) + 0.000093283590195398 * x ** 2 - \
                0.0023491751192724 * x * torch.log(x)
            return torch.where(m, y, torch.zeros_like(rho))
        return -torch.sqrt(rho) - 0.00067314115586063 * rho ** 2 + \
            0.000000076514905604792 * rho ** 4


REGISTRY = {
    'zjw04': Zjw04, 'zjw04xc': Zjw04xc, 'zjw04uxc': Zjw04uxc,
    'zjw04xcp': Zjw04xcp, 'sutton90': AgSutton90, 'Be/1': AgrawalBe,
    'grimes': RWGrimes, 'mishinh': MishinH, 'msah11': AlFeMsah11,
}


def get_potential(name, dtype=torch.float64, params=None):
    return REGISTRY[name](params=params, dtype=dtype)
