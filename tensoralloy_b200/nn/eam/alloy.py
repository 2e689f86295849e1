"""
EamAlloyNN -- mirror of the reference's tensoralloy/nn/eam/alloy.py:27-196:
rho is a function of the NEIGHBOUR element (alloy.py:162-176).
"""
from tensoralloy_b200 import _lib
from tensoralloy_b200.nn.eam.eam import EamNN
from tensoralloy_b200.utils import Defaults


class EamAlloyNN(EamNN):
    tag = "alloy"
    kind = _lib.EAM_ALLOY

    def _get_hidden_sizes(self, hidden_sizes):
        """alloy.py:41-91 (only consulted by 'nn' functions)."""
        self._setup_kbody_terms()
        results = {}
        for element in self._elements:
            results[element] = {'rho': Defaults.hidden_sizes,
                                'embed': Defaults.hidden_sizes}
        for term in self._unique_kbody_terms:
            results[term] = {'phi': Defaults.hidden_sizes}
        if isinstance(hidden_sizes, dict):
            for section, val in hidden_sizes.items():
                if section in results:
                    results[section].update(val)
        else:
            import numpy as np
            value = np.atleast_1d(hidden_sizes).tolist()
            for section in results:
                for key in results[section]:
                    results[section][key] = value
        return results

    def _setup_potentials(self, custom_potentials=None):
        """alloy.py:93-126."""
        if isinstance(custom_potentials, str):
            potentials = {el: {"rho": custom_potentials, "embed": custom_potentials}
                          for el in self._elements}
            potentials.update({t: {"phi": custom_potentials}
                               for t in self._unique_kbody_terms})
            return potentials
        potentials = {el: {"rho": "nn", "embed": "nn"} for el in self._elements}
        potentials.update({t: {"phi": "nn"} for t in self._unique_kbody_terms})
        custom_potentials = custom_potentials or {}
        for element in self._elements:
            for key in ('rho', 'embed'):
                if key in custom_potentials.get(element, {}):
                    value = custom_potentials[element][key]
                    assert self._check_fn_avail(value)
                    potentials[element][key] = value
        for term in self._unique_kbody_terms:
            if 'phi' in custom_potentials.get(term, {}):
                value = custom_potentials[term]['phi']
                assert self._check_fn_avail(value)
                potentials[term]['phi'] = value
        return potentials

    def _rho_entry(self, centre, other):
        return self._fn_of(other, 'rho').rho(other)
