"""
Zhou-Johnson-Wadley (2004) EAM parameter sets -> libtab200 function tables.

Mirror of the reference classes `Zjw04`, `Zjw04xc`, `Zjw04uxc`, `Zjw04xcp`
(nn/eam/potentials/zjw04.py:155-696): same parameter names and defaults (the
published table, stored in data/zjw04.json), same section naming
(`Shared/<element>/<param>` in exported graphs).  The functions themselves are
evaluated on the GPU (csrc/potentials.cuh: zhou_exp, zhou_embed).
"""
import copy
import json
from pathlib import Path

from tensoralloy_b200 import _lib
from tensoralloy_b200.utils import get_elements_from_kbody_term

_DATA = Path(__file__).resolve().parents[3] / 'data' / 'zjw04.json'

RHO_KEYS = ('f_eq', 'beta', 'lamda', 'r_eq')
PHI_KEYS = ('A', 'alpha', 'kappa', 'B', 'beta', 'lamda', 'r_eq')
EMBED_KEYS = ('Fn0', 'Fn1', 'Fn2', 'Fn3', 'F0', 'F1', 'F2', 'F3', 'eta', 'Fe',
              'rho_e', 'rho_s')


def _load():
    return json.loads(_DATA.read_text())


class Zjw04:
    """zjw04.py:155-413."""
    name = 'zjw04'
    embed_kind = _lib.FN_ZHOU_EMBED

    def __init__(self, params=None):
        self._params = copy.deepcopy(params) if params is not None \
            else self.defaults
        self._name = 'Zjw04'

    @property
    def defaults(self):
        return copy.deepcopy(_load()['zjw04'])

    @property
    def params(self):
        return self._params

    @property
    def fixed_parameters(self):
        """{section: [parameter, ...]} the reference never trains (zjw04.py:174-177: the
        embedding parameters and r_eq of every element)."""
        keys = ['F0', 'F1', 'F2', 'F3', 'Fn0', 'Fn1', 'Fn2', 'Fn3', 'Fe', 'eta', 'rho_e',
                'rho_s', 'r_eq']
        return {el: list(keys) for el in _load()['zjw04']}

    def set_param(self, section, key, value):
        """Override one variable (a loaded .pb constant or a trained value)."""
        self._params.setdefault(section, {})[key] = float(value)

    # -- table builders ---------------------------------------------------
    def rho(self, element_or_term):
        """Density contributed by a NEIGHBOUR of `element` (alloy.py:162-176)."""
        el = get_elements_from_kbody_term(element_or_term)[-1]
        p = self._params[el]
        return _lib.make_fn(_lib.FN_ZHOU_RHO, [p[k] for k in RHO_KEYS])

    def phi(self, kbody_term):
        a, b = get_elements_from_kbody_term(kbody_term)
        if a == b:
            p = self._params[a]
            return _lib.make_fn(_lib.FN_ZHOU_PHI, [p[k] for k in PHI_KEYS])
        pa, pb = self._params[a], self._params[b]
        vals = ([pa[k] for k in PHI_KEYS] + [pa[k] for k in RHO_KEYS] +
                [pb[k] for k in PHI_KEYS] + [pb[k] for k in RHO_KEYS])
        return _lib.make_fn(_lib.FN_ZHOU_PHI_MIX, vals)

    def embed(self, element):
        p = self._params[element]
        return _lib.make_fn(self.embed_kind, [p[k] for k in EMBED_KEYS])


class Zjw04xc(Zjw04):
    """zjw04.py:420-550 (sigmoid-blended embedding, adds Be = Mo)."""
    name = 'zjw04xc'
    embed_kind = _lib.FN_ZHOU_EMBED_XC

    def __init__(self, params=None):
        super().__init__(params)
        self._name = 'Zjw04xc'

    @property
    def defaults(self):
        p = copy.deepcopy(_load()['zjw04'])
        p['Be'] = copy.deepcopy(p['Mo'])
        return p

    @property
    def fixed_parameters(self):
        """zjw04.py:427-429 (and :585-587 for xcp): r_eq of the single elements."""
        return {sec: ['r_eq'] for sec in self.defaults
                if len(get_elements_from_kbody_term(sec)) == 1}


class Zjw04uxc(Zjw04xc):
    """zjw04.py:553-567."""
    name = 'zjw04uxc'

    def __init__(self, params=None):
        super().__init__(params)
        self._name = 'Zjw04uxc'

    fixed_parameters = property(lambda self: {})        # zjw04.py:566


class Zjw04xcp(Zjw04xc):
    """zjw04.py:570-696: the A-B pair has its own zhou_exp parameters."""
    name = 'zjw04xcp'

    def __init__(self, params=None):
        super().__init__(params)
        self._name = 'Zjw04xcp'

    @property
    def defaults(self):
        p = super().defaults
        for k, v in _load()['zjw04xcp_overrides'].items():
            p[k] = dict(v)
        return p

    def phi(self, kbody_term):
        a, b = get_elements_from_kbody_term(kbody_term)
        sec = a if a == b else kbody_term
        if sec not in self._params:
            sec = f'{b}{a}'
        p = self._params[sec]
        return _lib.make_fn(_lib.FN_ZHOU_PHI, [p[k] for k in PHI_KEYS])
