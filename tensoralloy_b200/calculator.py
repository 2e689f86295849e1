"""
TensorAlloyCalculator -- the ASE-compatible calculator of the reference
(tensoralloy/calculator.py:31-383) on top of libtab200.

Same constructor, properties, getters and result conventions:
  results['energy']       scalar eV
  results['forces']       [Nvap-1, 3] in GSL order  -> get_forces() maps to [N,3]
  results['stress']       [6] Voigt xx yy zz yz xz xy, eV/A^3
  results['energy/atom']  [Nvap-1]
  results['hessian']      [Nvap, 3, Nvap, 3]
The `sess.run(ops, feed_dict)` boundary (calculator.py:368-369) becomes: build
the GPU neighbour lists, run the fused kernels, copy the results back.

ASE itself is not installed in this image; when it is, this class subclasses
`ase.calculators.calculator.Calculator`, otherwise a minimal stand-in with the
same `get_property` / caching semantics.
"""
import json
from os.path import dirname
from typing import List

import numpy as np

from tensoralloy_b200.atoms import GPa
from tensoralloy_b200.nn.basic import BasicNN, exportable_properties
from tensoralloy_b200.precision import get_float_precision, precision_scope
from tensoralloy_b200.utils import ModeKeys

try:   # pragma: no cover - ASE is absent in the build image
    from ase.calculators.calculator import Calculator as _AseCalculator
    from ase.calculators.calculator import all_changes
    _HAVE_ASE = True
except Exception:   # noqa
    _HAVE_ASE = False
    all_changes = ['positions', 'numbers', 'cell', 'pbc',
                   'initial_charges', 'initial_magmoms']

    class _AseCalculator:
        """The subset of ase.calculators.calculator.Calculator the reference
        relies on: result caching keyed on the atoms' state + get_property."""
        implemented_properties: List[str] = []

        def __init__(self, restart=None, ignore_bad_restart_file=False,
                     label=None, atoms=None, **kwargs):
            self.atoms = None
            self.results = {}
            if atoms is not None:
                atoms.calc = self
                self.atoms = atoms.copy()

        def check_state(self, atoms):
            if self.atoms is None:
                return list(all_changes)
            changes = []
            if len(atoms) != len(self.atoms) or \
                    not np.array_equal(atoms.positions, self.atoms.positions):
                changes.append('positions')
            if atoms.get_chemical_symbols() != self.atoms.get_chemical_symbols():
                changes.append('numbers')
            if not np.array_equal(atoms.cell, self.atoms.cell):
                changes.append('cell')
            if not np.array_equal(atoms.pbc, self.atoms.pbc):
                changes.append('pbc')
            return changes

        def calculate(self, atoms=None, properties=('energy',),
                      system_changes=all_changes):
            if atoms is None:
                return
            mine = self.atoms
            if mine is not None and mine is not atoms and len(mine) == len(atoms) and \
                    np.array_equal(mine.numbers, atoms.numbers):
                # same atoms, new geometry: refresh the snapshot in place (a fresh copy of a
                # million-atom structure costs more in page faults than the evaluation)
                np.copyto(mine.positions, atoms.positions)
                mine.cell = np.array(atoms.cell, dtype=np.float64)
                mine.pbc = np.array(atoms.pbc, dtype=bool)
                mine.info = dict(atoms.info)
            else:
                self.atoms = atoms.copy()

        def get_property(self, name, atoms=None, allow_calculation=True):
            if name not in self.implemented_properties:
                raise NotImplementedError(f'{name} property not implemented')
            if atoms is None:
                atoms = self.atoms
                system_changes = []
            else:
                system_changes = self.check_state(atoms)
                if system_changes:
                    self.results = {}
            if name not in self.results:
                if not allow_calculation:
                    return None
                self.calculate(atoms, [name], system_changes)
            result = self.results[name]
            if isinstance(result, np.ndarray):
                result = result.copy()
            return result

        def get_potential_energy(self, atoms=None, force_consistent=False):
            return self.get_property('energy', atoms)


class TensorAlloyCalculator(_AseCalculator):
    """ASE-Calculator for TensorAlloy models (calculator.py:31)."""

    implemented_properties = list(exportable_properties) + ['energy/atom']
    default_parameters = {}
    nolabel = True

    def __init__(self, graph_model_path, atoms=None, serial_mode=False):
        """
        graph_model_path : str or BasicNN
            A `.npz` in the LAMMPS-native layout (`export_to_lammps_native`,
            atomic.py:304-480), a frozen `.pb` exported by the reference (`BasicNN.export`,
            basic.py:1017-1153) -- its transformer JSON and parameter constants
            are read, the TF graph itself is NOT executed -- or a model object
            with an attached transformer.
        serial_mode : bool
            Accepted for compatibility (calculator.py:70-75); the GPU path has
            no host thread pool to restrict.
        """
        super().__init__(restart=None, ignore_bad_restart_file=False, label=None,
                         atoms=atoms)
        self._serial_mode = serial_mode
        self._mode = ModeKeys.PREDICT
        self._model_timestamp = None
        if isinstance(graph_model_path, BasicNN):
            nn = graph_model_path
            if nn.transformer is None:
                raise ValueError("A descriptor transformer must be attached.")
            self._graph_model_path = None
            self._model_dir = None
            self._nn = nn
            self._fp_precision = get_float_precision().name
            self._api_version = "1.1"
            self._predict_properties = list(nn.predict_properties)
        elif str(graph_model_path).endswith('.npz'):
            # the LAMMPS-native layout (atomic.py:304-480), io/native.py
            from tensoralloy_b200.io.native import read_lammps_native
            nn, precision = read_lammps_native(graph_model_path)
            self._graph_model_path = graph_model_path
            self._model_dir = dirname(graph_model_path)
            self._nn = nn
            self._fp_precision = precision
            self._api_version = "1.1"
            self._predict_properties = list(nn.predict_properties)
        else:
            from tensoralloy_b200.io.graph_model import load_graph_model
            loaded = load_graph_model(graph_model_path)
            self._graph_model_path = graph_model_path
            self._model_dir = dirname(graph_model_path)
            self._nn = loaded.nn
            self._fp_precision = loaded.precision
            self._api_version = loaded.api_version
            self._model_timestamp = loaded.timestamp
            self._predict_properties = list(loaded.predict_properties)
        self._transformer = self._nn.transformer
        self._is_finite_temperature = self._nn.is_finite_temperature
        self._variational_energy = self._nn.variational_energy
        if 'energy/atom' not in self._predict_properties:
            self._predict_properties.append('energy/atom')
        self.implemented_properties = self._predict_properties
        self._ncalls = 0
        self._prerequisite_properties = []

    # -- reference properties ---------------------------------------------
    elements = property(lambda self: self._transformer.elements)
    transformer = property(lambda self: self._transformer)
    predict_properties = property(lambda self: self._predict_properties)
    api_version = property(lambda self: self._api_version)
    variational_energy = property(lambda self: self._variational_energy)
    ncalls = property(lambda self: self._ncalls)
    nn = property(lambda self: self._nn)

    def get_model_timestamp(self):
        return self._model_timestamp

    def get_magnetic_moment(self, atoms=None):
        return None

    def get_magnetic_moments(self, atoms=None):
        return None

    def get_electron_entropy(self, atoms=None):
        return self.get_property('eentropy', atoms=atoms)

    def get_free_energy(self, atoms=None):
        return self.get_property('free_energy', atoms=atoms)

    def get_atomic(self, atoms=None, prop="energy"):
        """calculator.py:217-226."""
        atoms = atoms or self.atoms
        values = self.get_property(f'{prop}/atom', atoms=atoms)
        values = np.insert(values, 0, 0, 0)
        clf = self.transformer.get_vap_transformer(atoms)
        return clf.map_array(values.reshape((-1, 1)), reverse=True).flatten()

    def get_hessian(self, atoms=None):
        """[3N, 3N] Hessian (calculator.py:228-241)."""
        atoms = atoms or self.atoms
        hessian = self.get_property('hessian', atoms)
        clf = self.transformer.get_vap_transformer(atoms)
        return clf.reverse_map_hessian(hessian)

    def get_forces(self, atoms=None):
        """calculator.py:243-249."""
        atoms = atoms or self.atoms
        forces = self.get_property('forces', atoms)
        clf = self.transformer.get_vap_transformer(atoms)
        if clf.is_identity:
            return forces
        forces = np.insert(forces, 0, 0, 0)
        return clf.map_forces(forces, reverse=True)

    def get_stress(self, atoms=None, voigt=True):
        """calculator.py:251-277."""
        atoms = atoms or self.atoms
        stress = self.get_property('stress', atoms)
        if not voigt:
            xx, yy, zz, yz, xz, xy = stress
            stress = np.array([[xx, xy, xz], [xy, yy, yz], [xz, yz, zz]])
        return stress

    def get_total_pressure(self, atoms=None):
        """calculator.py:279-295 (GPa)."""
        stress = self.get_stress(atoms)
        return np.mean(stress[:3]) * (-1.0) / GPa

    def get_elastic_constant_tensor(self, atoms=None):
        """calculator.py:297-322."""
        atoms = atoms or self.atoms
        assert atoms.pbc.all()
        elastic = self.get_property('elastic', atoms, allow_calculation=True)
        for i in range(6):
            for j in range(i + 1, 6):
                elastic[j, i] = elastic[i, j]
        return elastic

    def set_prerequisite_properties(self, properties: List[str]):
        for prop in properties:
            if prop in self.implemented_properties:
                self._prerequisite_properties.append(prop)

    def reset_call_counter(self):
        self._ncalls = 0

    def calculate_batch(self, images, properties=('energy', 'forces', 'stress')):
        """Batched inference (BASELINE config 2: "batched E/F/stress inference"): the
        structures of `images` share ONE neighbour handle and every kernel runs once over
        the whole batch (`tab_nbr_build_batch`).  The reference has no calculator-level
        batch call (its batches exist only inside the tf.data training pipeline,
        universal.py:921-1388); this is the inference-side equivalent.  Returns one dict
        per structure: 'energy', 'forces' [N,3] in the structure's own (ASE) atom order,
        'stress' [6] Voigt eV/A^3, 'energy/atom' [N] (+ 'eentropy', 'free_energy' for
        finite-temperature models, whose forces derive from the free energy)."""
        with precision_scope(self._fp_precision):
            properties = set(properties)
            for prop in properties:
                if prop not in self.implemented_properties:
                    raise KeyError(prop)
            want_stress = bool({'stress', 'total_pressure'} & properties)
            want_forces = 'forces' in properties or want_stress
            batch = self._transformer.get_batch_features(images)
            raw = self._nn.evaluate_batch(batch, want_forces, want_stress, True)
            dtype = np.float64 if self._fp_precision == 'high' else np.float32
            energy = raw['energy'].astype(dtype)
            eatom = raw['energy/atom'].astype(dtype)
            forces = raw['forces'].astype(dtype) if want_forces else None
            if want_stress:
                with np.errstate(divide='ignore', invalid='ignore'):   # clusters: V = 0
                    stress = raw['virial'] / batch.volumes[:, None, None]
                    pressure = (np.trace(stress, axis1=1, axis2=2) /
                                (-3.0 * GPa)).astype(dtype)
                virial = raw['virial'].astype(dtype)
                voigt = stress[:, [0, 1, 2, 1, 0, 0], [0, 1, 2, 2, 2, 1]].astype(dtype)
            off = batch.offsets
            out = []
            for s in range(batch.n_struct):
                lo, hi = off[s], off[s + 1]
                res = {'energy': energy[s], 'energy/atom': eatom[lo:hi]}
                for key in ('eentropy', 'free_energy'):      # finite-temperature models
                    if key in raw:
                        res[key] = dtype(raw[key][s])
                if want_forces:
                    res['forces'] = forces[lo:hi]
                if want_stress:
                    res['virial'] = virial[s]
                    res['stress'] = voigt[s]
                    res['total_pressure'] = pressure[s]
                out.append(res)
            self._ncalls += 1
        return out

    # -- the hot call ---------------------------------------------------------
    def calculate(self, atoms=None, properties=('energy', 'forces'),
                  system_changes=all_changes, debug_mode=False, extra_ops=None):
        """calculator.py:335-370."""
        _AseCalculator.calculate(self, atoms, properties, system_changes)
        atoms = atoms if atoms is not None else self.atoms
        with precision_scope(self._fp_precision):
            properties = set(properties).union(self._prerequisite_properties)
            for prop in properties:
                if prop not in self.implemented_properties:
                    raise KeyError(prop)
            features = self._transformer.get_constant_features(atoms)
            want_stress = bool({'stress', 'total_pressure', 'elastic', 'total_stress'}
                               & properties)
            want_forces = 'forces' in properties or want_stress
            want_atomic = bool({'energy/atom', 'atomic'} & properties) or \
                self._is_finite_temperature
            raw = self._nn._evaluate(features, want_forces, want_stress, want_atomic)
            if 'hessian' in properties:
                raw['hessian'] = self._nn._hessian(features)
            if 'elastic' in properties:
                raw['elastic'] = self._nn._elastic(features)
            wanted = list(properties)
            if want_stress:
                wanted.append('stress')
            pred = self._nn._finalize(raw, features, wanted)
            if 'elastic' in raw:
                pred['elastic'] = raw['elastic']
            # property names of legacy exports (Metadata/ops of api 1.0 files)
            if 'atomic' in properties:
                pred['atomic'] = pred['energy/atom']
            if 'free_energy' in properties and 'free_energy' not in pred:
                pred['free_energy'] = pred['energy']
            if 'total_stress' in properties and 'virial' in pred:
                pred['total_stress'] = pred['virial'] / features.volume
            self.results = pred
            self._ncalls += 1
