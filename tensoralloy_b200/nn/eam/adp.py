"""
AdpNN -- mirror of the reference's tensoralloy/nn/eam/adp.py:30-586: EAM/alloy
plus dipole and quadrupole terms.  The dipole / quadrupole energies are squared
PER K-BODY TERM (per neighbour species), exactly as the reference does
(adp.py:370-389, 455-494).

`_may_insert_spline_fn`, called by the reference at adp.py:154,176, is defined
nowhere upstream; the intent -- an optional tabulated spline in
place of the analytic function -- is served by the 'spline@...' potentials of
this build, not reproduced as a crash.
"""
import numpy as np

from tensoralloy_b200 import _lib
from tensoralloy_b200.nn.eam.alloy import EamAlloyNN
from tensoralloy_b200.utils import Defaults


class AdpNN(EamAlloyNN):
    scope = "ADP"
    tag = "adp"
    kind = _lib.EAM_ADP

    def _get_hidden_sizes(self, hidden_sizes):
        """adp.py:52-106."""
        self._setup_kbody_terms()
        results = {}
        for element in self._elements:
            results[element] = {'rho': Defaults.hidden_sizes,
                                'embed': Defaults.hidden_sizes}
        for term in self._unique_kbody_terms:
            results[term] = {'phi': Defaults.hidden_sizes,
                             'dipole': Defaults.hidden_sizes,
                             'quadrupole': Defaults.hidden_sizes}
        if isinstance(hidden_sizes, dict):
            for section, val in hidden_sizes.items():
                if section in results:
                    results[section].update(val)
        else:
            value = np.atleast_1d(hidden_sizes).tolist()
            for section in results:
                for key in results[section]:
                    results[section][key] = value
        return results

    def _setup_potentials(self, custom_potentials=None):
        """adp.py:108-150."""
        keys = ('phi', 'dipole', 'quadrupole')
        if isinstance(custom_potentials, str):
            potentials = {el: {"rho": custom_potentials, "embed": custom_potentials}
                          for el in self._elements}
            potentials.update({t: {k: custom_potentials for k in keys}
                               for t in self._unique_kbody_terms})
            return potentials
        potentials = {el: {"rho": "nn", "embed": "nn"} for el in self._elements}
        potentials.update({t: {k: "nn" for k in keys}
                           for t in self._unique_kbody_terms})
        custom_potentials = custom_potentials or {}
        for element in self._elements:
            for key in ('rho', 'embed'):
                if key in custom_potentials.get(element, {}):
                    value = custom_potentials[element][key]
                    assert self._check_fn_avail(value)
                    potentials[element][key] = value
        for term in self._unique_kbody_terms:
            for key in keys:
                if key in custom_potentials.get(term, {}):
                    value = custom_potentials[term][key]
                    assert self._check_fn_avail(value)
                    potentials[term][key] = value
        return potentials
