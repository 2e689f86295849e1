"""
EamFsNN -- mirror of the reference's tensoralloy/nn/eam/fs.py:27-203
(Finnis-Sinclair): rho is a function of the ORDERED pair (centre, neighbour)
(fs.py:180-203); phi is keyed by the sorted pair.
"""
import numpy as np

from tensoralloy_b200 import _lib
from tensoralloy_b200.nn.eam.eam import EamNN
from tensoralloy_b200.utils import Defaults, get_kbody_terms


class EamFsNN(EamNN):
    tag = "fs"
    kind = _lib.EAM_FS

    def _get_hidden_sizes(self, hidden_sizes):
        """fs.py:41-101."""
        self._setup_kbody_terms()
        self._all_kbody_terms = get_kbody_terms(self._elements, angular=False)[0]
        results = {}
        for element in self._elements:
            results[element] = {'embed': Defaults.hidden_sizes}
        for term in self._all_kbody_terms:
            results[term] = {'phi': Defaults.hidden_sizes, 'rho': Defaults.hidden_sizes}
        if isinstance(hidden_sizes, dict):
            for section, val in hidden_sizes.items():
                if section in results:
                    results[section].update(val)
        else:
            value = np.atleast_1d(hidden_sizes).tolist()
            for section in results:
                for key in results[section]:
                    results[section][key] = value
        for term in self._all_kbody_terms:
            if term not in self._unique_kbody_terms:
                del results[term]['phi']
        return results

    def _setup_potentials(self, custom_potentials=None):
        """fs.py:103-144."""
        if isinstance(custom_potentials, str):
            potentials = {el: {"embed": custom_potentials} for el in self._elements}
            potentials.update({t: {"phi": custom_potentials, "rho": custom_potentials}
                               for t in self._all_kbody_terms})
            return potentials
        potentials = {el: {"embed": "nn"} for el in self._elements}
        potentials.update({t: {"phi": "nn", "rho": "nn"}
                           for t in self._all_kbody_terms})
        custom_potentials = custom_potentials or {}
        for element in self._elements:
            if 'embed' in custom_potentials.get(element, {}):
                value = custom_potentials[element]['embed']
                assert self._check_fn_avail(value)
                potentials[element]['embed'] = value
        for term in self._all_kbody_terms:
            if 'rho' in custom_potentials.get(term, {}):
                value = custom_potentials[term]['rho']
                assert self._check_fn_avail(value)
                potentials[term]['rho'] = value
            if term not in self._unique_kbody_terms:
                del potentials[term]['phi']
        for term in self._unique_kbody_terms:
            if 'phi' in custom_potentials.get(term, {}):
                value = custom_potentials[term]['phi']
                assert self._check_fn_avail(value)
                potentials[term]['phi'] = value
        return potentials

    def _rho_entry(self, centre, other):
        term = f'{centre}{other}'
        return self._fn_of(term, 'rho').rho(term)
