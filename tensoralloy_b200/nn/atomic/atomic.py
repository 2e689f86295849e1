"""
AtomicNN -- mirror of the reference's tensoralloy/nn/atomic/atomic.py:60-302:
per-element MLP (1x1 convolutions = per-atom dense layers,
nn/convolutional.py:154-300) over atomic descriptors, optional min-max input
normalisation (:157-195) and an output bias used as atomic static energy
(:245-259).  Variable names follow the reference so that exported graphs load:
    Atomic/<El>/Conv1d{k}/kernel [1,in,out], .../bias [out],
    Atomic/<El>/Output/kernel [1,in,1], .../bias [1], Atomic/<El>/xlo|xhi [1,1,D].
"""
from typing import List

import numpy as np

from tensoralloy_b200 import _lib
from tensoralloy_b200.nn.atomic.sf import SymmetryFunction
from tensoralloy_b200.nn.basic import BasicNN
from tensoralloy_b200.precision import get_float_dtype
from tensoralloy_b200.utils import Defaults


class AtomicNN(BasicNN):
    scope = "Atomic"

    def __init__(self, elements: List[str], descriptor=None, hidden_sizes=None,
                 activation=None, kernel_initializer='he_normal', minmax_scale=True,
                 use_resnet_dt=False, atomic_static_energy=None,
                 use_atomic_static_energy=True, fixed_atomic_static_energy=False,
                 minimize_properties=('energy', 'forces'),
                 export_properties=('energy', 'forces')):
        super().__init__(elements=elements, hidden_sizes=hidden_sizes,
                         activation=activation,
                         minimize_properties=minimize_properties,
                         export_properties=export_properties)
        self._kernel_initializer = kernel_initializer
        self._minmax_scale = minmax_scale
        self._use_resnet_dt = use_resnet_dt
        self._atomic_static_energy = atomic_static_energy or {}
        self._use_atomic_static_energy = use_atomic_static_energy
        self._fixed_atomic_static_energy = fixed_atomic_static_energy
        if isinstance(descriptor, dict):
            d = dict(descriptor)
            cls = d.pop('@class', d.pop('class', 'SymmetryFunction'))
            d.pop('@module', None)
            if cls == 'GenericRadialAtomicPotential':
                from tensoralloy_b200.nn.atomic.grap import GenericRadialAtomicPotential
                descriptor = GenericRadialAtomicPotential(**d)
            else:
                descriptor = SymmetryFunction(**d)
        if descriptor is None:
            descriptor = SymmetryFunction(self._elements)
        self._descriptor = descriptor
        self._variables = {}
        self._model = None
        self._out = None

    descriptor = property(lambda self: self._descriptor)
    activation = property(lambda self: self._activation)
    use_resnet_dt = property(lambda self: self._use_resnet_dt)
    use_atomic_static_energy = property(lambda self: self._use_atomic_static_energy)

    def export_to_lammps_native(self, model_path, checkpoint=None,
                                use_ema_variables=True, dtype=np.float64):
        """The `.npz` model file of LAMMPS `pair_style tensoralloy/native`
        (atomic.py:304-480), written from the variables this object holds
        (`checkpoint` / `use_ema_variables` are accepted for signature parity: there is
        no TF checkpoint to restore).  io/native.py holds the layout."""
        from tensoralloy_b200.io.native import write_lammps_native
        write_lammps_native(self, model_path, dtype=dtype)

    def as_dict(self):
        return {"class": self.__class__.__name__, "elements": self._elements,
                "hidden_sizes": self._hidden_sizes, "activation": self._activation,
                'kernel_initializer': self._kernel_initializer,
                'minmax_scale': self._minmax_scale,
                'use_resnet_dt': self._use_resnet_dt,
                'use_atomic_static_energy': self._use_atomic_static_energy,
                'fixed_atomic_static_energy': self._fixed_atomic_static_energy,
                'atomic_static_energy': self._atomic_static_energy,
                "minimize_properties": self._minimize_properties,
                "export_properties": self._export_properties,
                "descriptor": self._descriptor.as_dict()}

    # -- variables -----------------------------------------------------------
    def _dim(self):
        return self._descriptor.dimension(self._transformer.angular)

    def initialize_variables(self, seed=Defaults.seed):
        """he_normal kernels / zero biases (init_ops.py:81-123), output bias =
        atomic static energy (atomic.py:245-259), xlo = 1000, xhi = 0
        (atomic.py:181-182)."""
        rng = np.random.default_rng(seed)
        dim = self._dim()
        if getattr(self._descriptor, 'algorithm', None) == 'nn':
            from tensoralloy_b200.nn.atomic.grap_nn import initialize_filter_variables
            initialize_filter_variables(self, np.random.default_rng(seed + 1))
        for el in self._elements:
            sizes = [dim] + list(self._hidden_sizes[el])
            for k in range(len(sizes) - 1):
                fan_in = sizes[k]
                # truncated normal, stddev sqrt(2/fan_in) (keras he_normal)
                w = rng.normal(size=(sizes[k], sizes[k + 1]))
                w = np.clip(w, -2.0, 2.0) * np.sqrt(2.0 / fan_in) / 0.87962566103423978
                self.set_variable(f"{self.scope}/{el}/Conv1d{k + 1}/kernel", w[None])
                self.set_variable(f"{self.scope}/{el}/Conv1d{k + 1}/bias",
                                  np.zeros(sizes[k + 1]))
            w = np.clip(rng.normal(size=(sizes[-1], 1)), -2.0, 2.0) * \
                np.sqrt(2.0 / sizes[-1]) / 0.87962566103423978
            self.set_variable(f"{self.scope}/{el}/Output/kernel", w[None])
            if self._use_atomic_static_energy:
                self.set_variable(f"{self.scope}/{el}/Output/bias",
                                  np.array([self._atomic_static_energy.get(el, 0.0)]))
            if self._minmax_scale:
                self.set_variable(f"{self.scope}/{el}/xlo", np.full((1, 1, dim), 1000.0))
                self.set_variable(f"{self.scope}/{el}/xhi", np.zeros((1, 1, dim)))

    def set_variable(self, name, value):
        self._variables[name] = np.asarray(value, dtype=np.float64)
        self._model = None
        self._filter_eval = None

    def get_variable(self, name):
        return self._variables[name]

    variables = property(lambda self: self._variables)

    def mlp_params(self, element):
        """The element's layers as plain arrays (also what the oracle consumes)."""
        W, b = [], []
        k = 1
        while f"{self.scope}/{element}/Conv1d{k}/kernel" in self._variables:
            w = self._variables[f"{self.scope}/{element}/Conv1d{k}/kernel"]
            W.append(w.reshape(w.shape[-2], w.shape[-1]))
            b.append(self._variables[f"{self.scope}/{element}/Conv1d{k}/bias"].reshape(-1))
            k += 1
        w = self._variables[f"{self.scope}/{element}/Output/kernel"]
        W.append(w.reshape(w.shape[-2], w.shape[-1]))
        ob = self._variables.get(f"{self.scope}/{element}/Output/bias")
        use_ob = self._use_atomic_static_energy and ob is not None
        b.append(ob.reshape(-1) if use_ob else None)
        xlo = self._variables.get(f"{self.scope}/{element}/xlo")
        xhi = self._variables.get(f"{self.scope}/{element}/xhi")
        mm = self._minmax_scale and xlo is not None and xhi is not None
        return dict(weights=W, biases=b, activation=self._activation,
                    use_resnet_dt=self._use_resnet_dt, output_bias=use_ob,
                    out_bias=(ob.reshape(-1) if use_ob else None),
                    xlo=xlo.reshape(-1) if mm else None,
                    xhi=xhi.reshape(-1) if mm else None)

    # -- device model ----------------------------------------------------------
    def _device_model(self):
        if self._model is not None:
            return self._model
        if not self._variables:
            self.initialize_variables()
        clf = self._transformer
        sf = self._descriptor
        if getattr(sf, 'uses_torch_path', lambda: False)():
            raise ValueError("GRAP with the `nn` filter network or moments 4 / 5 is not "
                             "served by the descriptor kernels: use "
                             "nn.atomic.grap_nn.GrapFilterTrainer")
        self._model = _lib.AtomicModel(
            len(self._elements), clf.rcut, clf.acut,
            sf.radial_sets(), sf.angular_sets() if clf.angular else None,
            sf.cutoff_function, [self.mlp_params(el) for el in self._elements],
            radial_kind=sf.radial_kind(), moments=sf.moments(),
            grap_flags=getattr(sf, 'grap_flags', lambda: 0)())
        return self._model

    def required_cutoff(self):
        clf = self._transformer
        return max(clf.rcut, clf.acut or 0.0) if clf.angular else clf.rcut

    def get_descriptors(self, features):
        """Raw G2/G4 descriptors [N, D] in caller atom order (sf.py:184-215)."""
        g = self._device_model().descriptors(features.nbr,
                                             get_float_dtype().tab_precision)
        return g.cpu().numpy()

    def _evaluate(self, features, want_forces, want_virial, want_atomic):
        if getattr(self._descriptor, 'uses_torch_path', lambda: False)():
            return self._evaluate_filter_network(features, want_forces or want_virial)
        return self._evaluate_single(features, want_forces, want_virial, want_atomic)

    def evaluate_batch(self, batch, want_forces=True, want_virial=True, want_atomic=True):
        if getattr(self._descriptor, 'uses_torch_path', lambda: False)():
            raise NotImplementedError(
                "batched inference of a GRAP model on the torch path: queue the structures in "
                "nn.atomic.grap_nn.GrapFilterTrainer and call .evaluate()")
        return super().evaluate_batch(batch, want_forces, want_virial, want_atomic)

    def _evaluate_filter_network(self, features, want_forces):
        """GRAP `nn` algorithm (nn/atomic/grap_nn.py): torch on the device over the
        library's pair vectors and pair-force op."""
        from tensoralloy_b200.nn.atomic.grap_nn import FilterEvaluator
        if getattr(self, '_filter_eval', None) is None:      # reset by set_variable
            self._filter_eval = FilterEvaluator(self)
        e_atom, f, w = self._filter_eval(features.nbr, features.types, want_forces=want_forces)
        raw = {'energy': float(e_atom.sum().item()),
               'energy/atom': e_atom.double().cpu().numpy()}
        if want_forces:
            raw['forces'] = f.double().cpu().numpy()
            raw['virial'] = w[0].double().cpu().numpy()
        return raw
