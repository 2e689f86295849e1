"""
TemperatureDependentAtomicNN / BeNN -- mirror of the reference's finite-temperature
AtomicNN (tensoralloy/nn/atomic/finite_temperature.py:29-304,
nn/atomic/special/beryllium.py:17-77, options nn/atomic/dataclasses.py:27-33).

Per atom of element `el` with descriptors x [D] and the electron temperature T
(eV, `atoms.info['etemperature']`, transformer/universal.py:880):
    H  = net_H(x)                 hidden = options.layers[:-1], out = layers[-1] (linear,
                                  with bias), activation = options.activation
    Ht = [H, T]                   finite_temperature.py:92-118
    S  = net_S(Ht) (* T if algo == 'Sommerfeld')      :120-166   (BeNN: analytic free-
                                  electron S(T) * softplus(net_S(Ht)), beryllium.py:23-77)
    U  = net_U(Ht)                output bias = atomic static energy   :168-209
    F  = U - T S                  :296-301
'energy' = sum U, 'eentropy' = sum S, 'free_energy' = sum F; forces and stress derive
from the FREE energy (variational_energy, basic.py:190-202).
Variables keep the reference's names:  TD/<El>/{H,S,U}/Conv1d{k}/{kernel,bias},
TD/<El>/{H,S,U}/Output/{kernel,bias}, TD/<El>/{xlo,xhi}.

Device path: descriptors G and the force / virial assembly F = J(R)^T dF/dG are the
libtab200 kernels (tab_atomic_descriptors, tab_atomic_forces: k_sf_forward,
k_sf_backward, k_sf_collect); the three heads, the entropy model and dF/dG between them are
ONE kernel (tab_td_eval, csrc/td_heads.cu: tiles of 8 atoms, activations in shared memory).
The torch formulation of the heads (`_net`, `_entropy`) remains for the training step
(`TemperatureDependentTrainer`: autograd over the parameters).
There is no CPU path: `_lib` raises without libtab200.so and the tensors live on cuda.
"""
from typing import List

import numpy as np

from tensoralloy_b200.atoms_utils import get_electron_temperature
from tensoralloy_b200.nn.atomic.atomic import AtomicNN
from tensoralloy_b200.precision import get_float_dtype


class FiniteTemperatureOptions:
    """nn/atomic/dataclasses.py:27-33."""

    def __init__(self, activation="softplus", layers=(128, 128), algo="default"):
        self.activation = activation
        self.layers = list(layers)
        self.algo = algo


class TemperatureDependentAtomicNN(AtomicNN):
    scope = "TD"
    special = None

    def __init__(self, elements: List[str], descriptor=None, hidden_sizes=None,
                 activation=None, kernel_initializer='he_normal', minmax_scale=True,
                 use_resnet_dt=False, atomic_static_energy=None,
                 use_atomic_static_energy=True, fixed_atomic_static_energy=False,
                 minimize_properties=('energy', 'forces'),
                 export_properties=('energy', 'forces'),
                 finite_temperature=None):
        super().__init__(elements=elements, descriptor=descriptor,
                         hidden_sizes=hidden_sizes, activation=activation,
                         kernel_initializer=kernel_initializer, minmax_scale=minmax_scale,
                         use_resnet_dt=use_resnet_dt,
                         atomic_static_energy=atomic_static_energy,
                         use_atomic_static_energy=use_atomic_static_energy,
                         fixed_atomic_static_energy=fixed_atomic_static_energy,
                         minimize_properties=minimize_properties,
                         export_properties=export_properties)
        if finite_temperature is None:
            finite_temperature = FiniteTemperatureOptions()
        elif isinstance(finite_temperature, dict):
            finite_temperature = FiniteTemperatureOptions(**finite_temperature)
        self._finite_temperature = finite_temperature
        self._torch_params = None

    finite_temperature_options = property(lambda self: self._finite_temperature)

    @property
    def is_finite_temperature(self) -> bool:
        return True

    def as_dict(self):
        d = super().as_dict()
        d['finite_temperature'] = dict(self._finite_temperature.__dict__)
        return d

    # -- variables -----------------------------------------------------------
    def _head_sizes(self, el):
        ft = self._finite_temperature
        dim = self._dim()
        hid = list(self._hidden_sizes[el])
        return {'H': ([dim] + list(ft.layers[:-1]), ft.layers[-1]),
                'S': ([ft.layers[-1] + 1] + hid, 1),
                'U': ([ft.layers[-1] + 1] + hid, 1)}

    def _s_output_bias(self):
        return True          # finite_temperature.py:152 (BeNN: False)

    def initialize_variables(self, seed=611):
        rng = np.random.default_rng(seed)

        def he(n_in, n_out):
            w = np.clip(rng.normal(size=(n_in, n_out)), -2.0, 2.0)
            return w * np.sqrt(2.0 / n_in) / 0.87962566103423978

        dim = self._dim()
        for el in self._elements:
            for head, (sizes, n_out) in self._head_sizes(el).items():
                base = f"{self.scope}/{el}/{head}"
                for k in range(len(sizes) - 1):
                    self.set_variable(f"{base}/Conv1d{k + 1}/kernel",
                                      he(sizes[k], sizes[k + 1])[None])
                    self.set_variable(f"{base}/Conv1d{k + 1}/bias", np.zeros(sizes[k + 1]))
                self.set_variable(f"{base}/Output/kernel", he(sizes[-1], n_out)[None])
                if head == 'U':
                    mean = self._atomic_static_energy.get(el, 0.0) \
                        if self._use_atomic_static_energy else 0.0
                    self.set_variable(f"{base}/Output/bias", np.full(n_out, mean))
                elif head == 'H' or self._s_output_bias():
                    self.set_variable(f"{base}/Output/bias", np.zeros(n_out))
            if self._minmax_scale:
                self.set_variable(f"{self.scope}/{el}/xlo", np.full((1, 1, dim), 1000.0))
                self.set_variable(f"{self.scope}/{el}/xhi", np.zeros((1, 1, dim)))

    def set_variable(self, name, value):
        super().set_variable(name, value)
        self._torch_params = None
        self._td_heads = None

    def head_params(self, el, head):
        """One head's layers as plain arrays (what the oracle consumes)."""
        base = f"{self.scope}/{el}/{head}"
        W, b = [], []
        k = 1
        while f"{base}/Conv1d{k}/kernel" in self._variables:
            w = self._variables[f"{base}/Conv1d{k}/kernel"]
            W.append(w.reshape(w.shape[-2], w.shape[-1]))
            b.append(self._variables[f"{base}/Conv1d{k}/bias"].reshape(-1))
            k += 1
        w = self._variables[f"{base}/Output/kernel"]
        W.append(w.reshape(w.shape[-2], w.shape[-1]))
        ob = self._variables.get(f"{base}/Output/bias")
        b.append(None)
        act = self._finite_temperature.activation if head == 'H' else self._activation
        return dict(weights=W, biases=b, out_bias=None if ob is None else ob.reshape(-1),
                    activation=act, use_resnet_dt=self._use_resnet_dt)

    def td_params(self, el):
        p = {h: self.head_params(el, h) for h in ('H', 'S', 'U')}
        p['algo'] = self._finite_temperature.algo
        p['special'] = self.special
        return p

    def mlp_params(self, element):
        """The geometry-side device model (descriptors + force assembly) carries no
        network; a one-layer placeholder keeps tab_atomic_create's contract."""
        return dict(weights=[np.zeros((self._dim(), 1))], biases=[None],
                    activation='softplus', use_resnet_dt=False, output_bias=False,
                    out_bias=None, xlo=None, xhi=None)

    def minmax(self, element):
        """(xlo, xhi) of atomic.py:157-195, or None."""
        xlo = self._variables.get(f"{self.scope}/{element}/xlo")
        xhi = self._variables.get(f"{self.scope}/{element}/xhi")
        if self._minmax_scale and xlo is not None and xhi is not None:
            return xlo.reshape(-1), xhi.reshape(-1)
        return None

    # -- evaluation --------------------------------------------------------------
    def _torch_heads(self, tdtype):
        import torch
        if self._torch_params is not None and self._torch_params[0] == tdtype:
            return self._torch_params[1]
        t = lambda a: None if a is None else torch.tensor(np.asarray(a), dtype=tdtype,
                                                          device='cuda')
        out = {}
        for el in self._elements:
            heads = {}
            for h in ('H', 'S', 'U'):
                p = self.head_params(el, h)
                heads[h] = dict(W=[t(w) for w in p['weights']],
                                b=[t(v) for v in p['biases']], ob=t(p['out_bias']),
                                act=p['activation'], resnet=p['use_resnet_dt'])
            mm = self.minmax(el)
            heads['xlo'], heads['xhi'] = (t(mm[0]), t(mm[1])) if mm else (None, None)
            out[el] = heads
        self._torch_params = (tdtype, out)
        return out

    @staticmethod
    def _net(q, x):
        from tensoralloy_b200.nn.atomic.training import _activation
        fn = _activation(q['act'])
        h = x
        nh = len(q['W']) - 1
        for k in range(nh):
            y = fn(h @ q['W'][k] + q['b'][k])
            if k and q['resnet'] and q['W'][k].shape[1] == q['W'][k - 1].shape[1]:
                h = y + h
            else:
                h = y
        out = h @ q['W'][nh]
        if q['ob'] is not None:
            out = out + q['ob']
        return out

    def _entropy(self, heads, Ht, T):
        """finite_temperature.py:120-166."""
        S = self._net(heads['S'], Ht)[:, 0]
        if self._finite_temperature.algo == "Sommerfeld":
            S = S * T
        return S

    def _device_heads(self):
        """The three networks of every element as one `tab_td` handle (csrc/td_heads.cu)."""
        from tensoralloy_b200 import _lib
        if getattr(self, '_td_heads', None) is not None:
            return self._td_heads
        heads = []
        for el in self._elements:
            nets = {}
            for h in ('H', 'S', 'U'):
                p = self.head_params(el, h)
                biases = list(p['biases'])
                biases[-1] = p['out_bias']
                nets[h] = dict(weights=p['weights'], biases=biases, activation=p['activation'],
                               use_resnet_dt=p['use_resnet_dt'],
                               output_bias=p['out_bias'] is not None)
            mm = self.minmax(el)
            if mm:
                nets['H']['xlo'], nets['H']['xhi'] = mm
            heads.append(nets)
        self._td_heads = _lib.TdHeads(self._dim(), heads, self._finite_temperature.algo,
                                      self.special)
        return self._td_heads

    def _run(self, nbr, types, t_atom, sid, nb, want_grad):
        """Shared core of the single-structure and the batched evaluation.  t_atom: per-atom
        electron temperature (cuda); sid: structure of each atom or None.  Returns per-atom
        U, S, F (float64 numpy), per-structure sums, and forces / virials [nb, 9].
        Descriptors (`tab_atomic_descriptors`) -> heads, entropy model and dF/dG
        (`tab_td_eval`) -> forces and virial of the free energy (`tab_atomic_forces`)."""
        import torch
        dt = get_float_dtype()
        model = self._device_model()
        n = int(types.shape[0])
        G = model.descriptors(nbr, dt.tab_precision)
        U, S, F, dfdg = self._device_heads().eval(types.to(torch.int32).contiguous(), G,
                                                  t_atom.to(torch.float64).contiguous(),
                                                  dt.tab_precision)
        forces = virial = None
        if want_grad:
            forces = torch.empty((n, 3), dtype=torch.float64, device='cuda')
            virial = torch.empty(nb * 9, dtype=torch.float64, device='cuda')
            model.forces_from_dedg(nbr, dfdg, forces, virial, dt.tab_precision)
        per_atom = torch.stack([U, S, F])
        if self._device_heads().status():
            raise RuntimeError("tab_td_eval: a weight transfer timed out on the device")
        if sid is None:
            sums = per_atom.sum(dim=1, keepdim=True)
        else:
            sums = torch.zeros((3, nb), dtype=torch.float64, device='cuda').index_add(
                1, sid, per_atom)
        return (per_atom.cpu().numpy(), sums.cpu().numpy(),
                None if forces is None else forces.cpu().numpy(),
                None if virial is None else virial.cpu().numpy().reshape(nb, 3, 3))

    def _evaluate(self, features, want_forces, want_virial, want_atomic):
        import torch
        n = features.n_atoms
        etemp = float(get_electron_temperature(features.atoms))
        types = torch.as_tensor(np.asarray(features.types), device='cuda').long()
        t_atom = torch.full((n,), etemp, dtype=torch.float64, device='cuda')
        per_atom, sums, forces, virial = self._run(features.nbr, types, t_atom, None, 1,
                                                   want_forces or want_virial)
        raw = {'energy': sums[0, 0], 'eentropy': sums[1, 0], 'free_energy': sums[2, 0]}
        if want_forces:
            raw['forces'] = forces
        if want_virial:
            raw['virial'] = virial[0].copy()
        if want_atomic:
            raw.update({'energy/atom': per_atom[0], 'eentropy/atom': per_atom[1],
                        'free_energy/atom': per_atom[2]})
        return raw

    def evaluate_batch(self, batch, want_forces=True, want_virial=True, want_atomic=True):
        """Batched evaluation (one neighbour handle, `tab_nbr_build_batch`): every
        structure carries its own electron temperature (`atoms.info['etemperature']`)."""
        import torch
        nb = batch.n_struct
        lens = np.diff(batch.offsets)
        sid = torch.as_tensor(np.repeat(np.arange(nb), lens), device='cuda').long()
        temps = np.array([float(get_electron_temperature(a)) for a in batch.images])
        t_atom = torch.as_tensor(np.repeat(temps, lens), device='cuda')
        types = torch.as_tensor(np.asarray(batch.types), device='cuda').long()
        per_atom, sums, forces, virial = self._run(batch.nbr, types, t_atom, sid, nb,
                                                   want_forces or want_virial)
        raw = {'energy': sums[0], 'eentropy': sums[1], 'free_energy': sums[2]}
        if want_forces:
            raw['forces'] = forces
        if want_virial:
            raw['virial'] = virial
        if want_atomic:
            raw.update({'energy/atom': per_atom[0], 'eentropy/atom': per_atom[1],
                        'free_energy/atom': per_atom[2]})
        return raw

    def _finalize(self, raw, features, properties):
        pred = super()._finalize(raw, features, properties)
        dtype = get_float_dtype().as_numpy_dtype
        for key in ('eentropy', 'free_energy'):
            pred[key] = dtype(raw[key])
        return pred


class BeNN(TemperatureDependentAtomicNN):
    """special/beryllium.py:17-77: the entropy head multiplies a fitted free-electron
    S(T) by softplus(net_S(Ht)); the net has no output bias."""
    special = 'Be'

    def _s_output_bias(self):
        return False

    def _entropy(self, heads, Ht, T):
        import torch
        t2 = T * T
        ft = torch.relu(1.0 - 1.45 * T) ** 2
        base = -0.5718444 * t2 * ft + 0.83744317 * T + (-0.2110962) * (1.0 - ft)
        dev = torch.nn.functional.softplus(self._net(heads['S'], Ht)[:, 0])
        return base * dev
