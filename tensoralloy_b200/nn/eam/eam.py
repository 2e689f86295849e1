"""
EamNN: base of the EAM-family models.  Mirror of the reference's
tensoralloy/nn/eam/eam.py:80-570 (constructor signature :88-94, `as_dict`
:137-147, potential bookkeeping :149-260); `_get_model_outputs` (:495-570) is
executed by csrc/eam.cu instead of a TF graph.
"""
from typing import List

import numpy as np

from tensoralloy_b200 import _lib
from tensoralloy_b200.nn.basic import BasicNN
from tensoralloy_b200.nn.eam.potentials import available_potentials
from tensoralloy_b200.precision import get_float_dtype
from tensoralloy_b200.utils import (Defaults, get_elements_from_kbody_term,
                                    get_kbody_terms)


class EamNN(BasicNN):
    scope = "EAM"
    tag = None
    kind = _lib.EAM_ALLOY

    def __init__(self, elements: List[str], custom_potentials=None,
                 hidden_sizes=None, fixed_functions=None,
                 minimize_properties=('energy', 'forces'),
                 export_properties=('energy', 'forces', 'stress')):
        self._fixed_functions = fixed_functions or []
        self._unique_kbody_terms = None
        self._kbody_terms = None
        super().__init__(elements=elements, hidden_sizes=hidden_sizes,
                         minimize_properties=minimize_properties,
                         export_properties=export_properties)
        self._potentials = self._setup_potentials(custom_potentials)
        self._empirical_functions = {
            key: cls() for key, cls in available_potentials.items()}
        # 'nn' functions (eam.py:174-190): one MLP per (section, function)
        from tensoralloy_b200.nn.eam.potentials.nn import NNFunctions
        self._nn = NNFunctions(self.scope, self._activation)
        for section, fns in self._potentials.items():
            for fn, name in fns.items():
                if name == 'nn':
                    self._nn.declare(fn, section, self._hidden_sizes[section][fn])
        self._nn.initialize(np.random.default_rng(Defaults.seed))
        self._empirical_functions['nn'] = self._nn
        self._model = None
        self._out = None
        assert self._kbody_terms and self._unique_kbody_terms

    unique_kbody_terms = property(lambda self: self._unique_kbody_terms)
    potentials = property(lambda self: self._potentials)

    def as_dict(self):
        return {"class": self.__class__.__name__,
                "elements": self._elements,
                "custom_potentials": self._potentials,
                "hidden_sizes": self._hidden_sizes,
                "fixed_functions": self._fixed_functions,
                "minimize_properties": self._minimize_properties,
                "export_properties": self._export_properties}

    @staticmethod
    def _check_fn_avail(name: str):
        if name.startswith("spline@"):
            return True
        return name.lower() == "nn" or name in available_potentials or \
            name.lower() in available_potentials

    def _setup_kbody_terms(self):
        kbody_terms = get_kbody_terms(self._elements, angular=False)[1]
        unique = []
        for element in self._elements:
            for term in kbody_terms[element]:
                a, b = get_elements_from_kbody_term(term)
                if a == b:
                    unique.append(term)
                else:
                    ab = "".join(sorted([a, b]))
                    if ab not in unique:
                        unique.append(ab)
        self._unique_kbody_terms = unique
        self._kbody_terms = kbody_terms

    # -- variables ---------------------------------------------------------
    def initialize_variables(self, seed=Defaults.seed):
        """he_normal kernels / zero biases of the 'nn' functions (init_ops.py:81-123)."""
        self._nn.initialize(np.random.default_rng(seed))
        self._model = None

    @property
    def variables(self):
        """The trainable 'nn' variables by reference name."""
        return self._nn.variables()

    def get_variable(self, name):
        """Value of `EAM/Shared/<section>/<param>` (potentials.py:171-200) or of an
        'nn' kernel / bias."""
        if self._nn.owns(name):
            return self._nn.variables()[name]
        scope, shared, section, key = name.split('/')
        for fn in self._empirical_functions.values():
            if section in fn.params and key in fn.params[section]:
                return fn.params[section][key]
        raise KeyError(name)

    def set_variable(self, name, value, potential=None):
        """Set one shared variable (e.g. a Const node of a frozen .pb).  Shared
        variables are keyed by section/param only (potentials.py:171-200), so
        the value is applied to every registered potential unless one is named."""
        if self._nn.owns(name):
            self._nn.set_variable(name, value)
            self._model = None
            return
        parts = name.split('/')
        section, key = parts[-2], parts[-1]
        for pname, fn in self._empirical_functions.items():
            if potential is None or pname == potential:
                if section in fn.params or potential is not None:
                    fn.set_param(section, key, value)
        self._model = None

    # -- device model --------------------------------------------------------
    def _rho_entry(self, centre, other):
        raise NotImplementedError

    def _fn_of(self, section, key):
        name = self._potentials[section][key]
        if name == 'nn':
            return self._nn
        if name.startswith("spline@"):
            if name not in self._empirical_functions:
                from tensoralloy_b200.nn.eam.potentials.spline import SplinePotential
                self._empirical_functions[name] = SplinePotential(name[len("spline@"):])
            return self._empirical_functions[name]
        return self._empirical_functions[name]

    def _device_model(self):
        if self._model is not None:
            return self._model
        els = self._elements
        rho = [self._rho_entry(a, b) for a in els for b in els]
        phi = []
        for a in els:
            for b in els:
                key = "".join(sorted([a, b])) if a != b else f"{a}{a}"
                phi.append(self._fn_of(key, 'phi').phi(key))
        embed = [self._fn_of(a, 'embed').embed(a) for a in els]
        dipole = quadrupole = None
        if self.kind == _lib.EAM_ADP:
            dipole, quadrupole = [], []
            for a in els:
                for b in els:
                    key = "".join(sorted([a, b])) if a != b else f"{a}{a}"
                    dipole.append(self._fn_of(key, 'dipole').dipole(key))
                    quadrupole.append(self._fn_of(key, 'quadrupole').quadrupole(key))
        # tabulated and 'nn' functions: one coefficient pool per model (units of 4
        # doubles); rebase the offsets of every provider after the first
        splines = [f for k, f in self._empirical_functions.items()
                   if k.startswith("spline@") or k in ('nn', 'msah11')]
        base = 0
        pools = []
        for sp in splines:
            pool = sp.pool()
            if not len(pool):
                continue
            if base:
                for fns in (rho, phi, embed, dipole or [], quadrupole or []):
                    for fn in fns:
                        if getattr(fn, '_owner', None) is sp and not getattr(fn, '_rebased', False):
                            fn.aux += base
                            fn._rebased = True
            pools.append(pool)
            base += len(pool)
        self._model = _lib.EamModel(self.kind, len(els), rho, phi, embed, dipole,
                                    quadrupole)
        if pools and base:
            import numpy as _np
            self._model.set_splines(_np.concatenate(pools))
        return self._model

    def _evaluate(self, features, want_forces, want_virial, want_atomic):
        return self._evaluate_single(features, want_forces, want_virial, want_atomic)

    # -- LAMMPS tables ---------------------------------------------------------
    def export_to_setfl(self, setfl: str, nr: int, dr: float, nrho: int, drho: float,
                        r0=1.0, rt=None, rho0=0.0, rhot=None, checkpoint=None,
                        lattice_constants=None, lattice_types=None,
                        use_ema_variables=True):
        """Export this model to a LAMMPS setfl table (eam/alloy, eam/fs or adp by model
        kind) -- the reference's `export_to_setfl` (alloy.py:198-381, fs.py, adp.py:588-794).
        The functions are tabulated by the GPU evaluators that also serve the energy /
        force kernels (`tab_eam_tabulate`) on r = k dr and rho = k drho, k = 0 .. n-1.  The
        plotting arguments (r0, rt, rho0, rhot) and the checkpoint arguments of the reference
        are accepted and ignored (no figures are drawn; the variables are the model's)."""
        from tensoralloy_b200.io import lammps as io
        from tensoralloy_b200.precision import precision_scope
        lattice_constants = lattice_constants or {}
        lattice_types = lattice_types or {}
        els = self._elements
        n_el = len(els)
        with precision_scope('high'):
            model = self._device_model()
            r = np.arange(nr) * float(dr)
            rho_grid = np.arange(nrho) * float(drho)
            tab = lambda which, idx, x: model.tabulate(which, idx, x)[0]
            embed = {el: tab('embed', a, rho_grid) for a, el in enumerate(els)}
            pair_keys = [("".join(sorted([els[i], els[j]])), i, j)
                         for i in range(n_el) for j in range(i, n_el)]
            phi = {key: tab('phi', i * n_el + j, r) for key, i, j in pair_keys}
        masses = self._atomic_masses()
        lat = [float(lattice_constants.get(el, 0.0)) for el in els]
        typ = [lattice_types.get(el, 'fcc') for el in els]
        rcut = float(nr) * float(dr)
        comments = (f"Contributor: tensoralloy_b200 ({self.__class__.__name__})",
                    "LAMMPS setfl format",
                    f"Conversion by tensoralloy_b200.nn.eam.{self.__class__.__name__}")
        if self.kind == _lib.EAM_FS:
            with precision_scope('high'):
                rho = {f"{a}{b}": tab('rho', i * n_el + j, r)
                       for i, a in enumerate(els) for j, b in enumerate(els)}
            io.write_fs_setfl(setfl, els, nrho, drho, nr, dr, rcut, embed, rho, phi, masses,
                              lat, typ, comments)
            return
        with precision_scope('high'):
            # eam/alloy: the density contributed BY an atom of `el` (alloy.py:162-176)
            rho = {el: tab('rho', j, r) for j, el in enumerate(els)}
            dip = quad = {}
            if self.kind == _lib.EAM_ADP:
                dip = {key: tab('dipole', i * n_el + j, r) for key, i, j in pair_keys}
                quad = {key: tab('quadrupole', i * n_el + j, r) for key, i, j in pair_keys}
        # the file lists pairs as (i, j <= i) keyed el_i el_j in the reader's sequence
        order = [f"{els[i]}{els[j]}" for i in range(n_el) for j in range(i, n_el)]
        rekey = lambda d: {f"{els[i]}{els[j]}": d[key] for key, i, j in pair_keys} if d else {}
        sp = lambda n, dx, y: io._spline(n, dx, y)
        table = io.SetFL(
            elements=list(els), rho={el: sp(nr, dr, rho[el]) for el in els},
            phi={k: sp(nr, dr, v) for k, v in rekey(phi).items()},
            embed={el: sp(nrho, drho, embed[el]) for el in els},
            dipole={k: sp(nr, dr, v) for k, v in rekey(dip).items()},
            quadrupole={k: sp(nr, dr, v) for k, v in rekey(quad).items()},
            nr=nr, dr=dr, nrho=nrho, drho=drho, rcut=rcut, atomic_masses=masses,
            lattice_constants=lat, lattice_types=typ)
        assert list(table.phi) == order
        io.write_setfl(setfl, table, comments, is_adp=self.kind == _lib.EAM_ADP)

    def _atomic_masses(self):
        from tensoralloy_b200.analysis.phonon import _MASSES
        return [float(_MASSES.get(el, 0.0)) for el in self._elements]

    def _elastic(self, features):
        """Elastic tensor [6,6] in GPa (nn/constraint/elastic.py:24-91) from the closed-form
        second derivatives (`tab_eam_elastic`); models without that kernel (ADP, 'nn'
        functions) use the central difference of the virial of BasicNN._elastic."""
        from tensoralloy_b200.atoms import GPa
        model = self._device_model()
        try:
            C = model.elastic(features.nbr).cpu().numpy()
        except _lib.TabError as exc:
            if 'status -4' not in str(exc):
                raise
            return super()._elastic(features)
        return C / features.volume / GPa

    def _hessian(self, features):
        """[Nvap, 3, Nvap, 3] Hessian in GSL order incl. the virtual atom, the
        layout of the reference's `Output/Hessian` op (basic.py:410-421)."""
        model = self._device_model()
        try:
            H = model.hessian(features.nbr).cpu().numpy()
        except _lib.TabError as exc:
            # ADP and 'nn' (MLP) functions have no closed-form second-derivative kernel: the
            # analytic forces are differentiated instead (BasicNN._hessian_from_forces)
            if 'status -4' not in str(exc):
                raise
            H = self._hessian_from_forces(features)
        return self._embed_hessian(H, features)
