"""
'nn' potential functions: rho(r), phi(r), F(rho), u(r), w(r) parametrised by a small
MLP of the scalar argument -- the reference's `EamNN._get_nn_fn`
(nn/eam/eam.py:174-190 -> `convolution1x1`, nn/convolutional.py:257-290): hidden
layers with bias and activation, a linear output unit WITHOUT bias, no ResNet
link.  Variable names follow the reference's scopes
    <EAM|ADP>/<Rho|Phi|Embed|Dipole|Quadrupole>/<key>/Conv{rank}d{k}/{kernel,bias}
    <...>/<key>/Output/kernel
(rank 2 for the pair functions, 1 for the embedding, eam.py:300-449).  The weights
are evaluated on the GPU as TAB_FN_MLP entries (csrc/potentials.cuh).
"""
import numpy as np

from tensoralloy_b200 import _lib

SCOPES = {'rho': ('Rho', 2), 'phi': ('Phi', 2), 'embed': ('Embed', 1),
          'dipole': ('Dipole', 2), 'quadrupole': ('Quadrupole', 2)}


class NNFunctions:
    name = 'nn'

    def __init__(self, scope, activation):
        self.scope = scope
        self.activation = activation
        self.params = {}
        self._sizes = {}          # (fn, key) -> hidden sizes
        self._weights = {}        # (fn, key) -> [W0 [1,h0], b0, W1 [h0,h1], b1, ..., Wout [h,1]]
        self._blocks = []
        self._offsets = {}
        self._n4 = 0

    def set_param(self, *args):
        pass

    # -- variables -------------------------------------------------------------
    def declare(self, fn, key, hidden_sizes):
        self._sizes[(fn, key)] = [int(h) for h in hidden_sizes]

    def prefix(self, fn, key):
        return f"{self.scope}/{SCOPES[fn][0]}/{key}"

    def initialize(self, rng):
        """he_normal kernels, zero biases (init_ops.py:81-123)."""
        for (fn, key), hidden in self._sizes.items():
            sizes = [1] + hidden
            arrays = []
            for k in range(len(hidden)):
                w = np.clip(rng.normal(size=(sizes[k], sizes[k + 1])), -2.0, 2.0) * \
                    np.sqrt(2.0 / sizes[k]) / 0.87962566103423978
                arrays += [w, np.zeros(sizes[k + 1])]
            arrays.append(np.clip(rng.normal(size=(sizes[-1], 1)), -2.0, 2.0) *
                          np.sqrt(2.0 / sizes[-1]) / 0.87962566103423978)
            self._weights[(fn, key)] = arrays
        self._invalidate()

    def variables(self):
        out = {}
        for (fn, key), arrays in self._weights.items():
            rank = SCOPES[fn][1]
            pre = self.prefix(fn, key)
            nh = (len(arrays) - 1) // 2
            for k in range(nh):
                out[f"{pre}/Conv{rank}d{k + 1}/kernel"] = arrays[2 * k]
                out[f"{pre}/Conv{rank}d{k + 1}/bias"] = arrays[2 * k + 1]
            out[f"{pre}/Output/kernel"] = arrays[-1]
        return out

    def set_variable(self, name, value):
        """name = <scope>/<Section>/<key>/Conv?d{k}/{kernel|bias} or .../Output/kernel."""
        parts = name.split('/')
        section, key, layer, what = parts[-4], parts[-3], parts[-2], parts[-1]
        fn = {v[0]: k for k, v in SCOPES.items()}[section]
        arrays = self._weights[(fn, key)]
        value = np.asarray(value, dtype=np.float64)
        if layer == 'Output':
            arrays[-1] = value.reshape(-1, 1)
        else:
            k = int(layer.split('d')[-1]) - 1
            if what == 'kernel':
                arrays[2 * k] = value.reshape(value.shape[-2], value.shape[-1])
            else:
                arrays[2 * k + 1] = value.reshape(-1)
        self._invalidate()

    def owns(self, name):
        parts = name.split('/')
        return len(parts) >= 5 and (parts[-2].startswith('Conv') or parts[-2] == 'Output')

    def weights(self, fn, key):
        return self._weights[(fn, key)]

    # -- device entries ----------------------------------------------------------
    def _invalidate(self):
        self._blocks, self._offsets, self._n4 = [], {}, 0

    def _entry(self, fn, key):
        tag = (fn, key)
        if tag not in self._offsets:
            arrays = self._weights[tag]
            flat = np.concatenate([np.asarray(a, dtype=np.float64).reshape(-1) for a in arrays])
            pad = (-len(flat)) % 4
            flat = np.concatenate([flat, np.zeros(pad)]).reshape(-1, 4)
            self._offsets[tag] = self._n4
            self._blocks.append(flat)
            self._n4 += len(flat)
        hidden = self._sizes[tag]
        entry = _lib.make_fn(_lib.FN_MLP,
                             [len(hidden), _lib.ACTIVATIONS[self.activation.lower()]] + hidden,
                             aux=self._offsets[tag])
        entry._owner = self
        return entry

    def pool(self):
        return np.concatenate(self._blocks) if self._blocks else np.zeros((0, 4))

    def rho(self, key):
        return self._entry('rho', key)

    def phi(self, key):
        return self._entry('phi', key)

    def embed(self, key):
        return self._entry('embed', key)

    def dipole(self, key):
        return self._entry('dipole', key)

    def quadrupole(self, key):
        return self._entry('quadrupole', key)
