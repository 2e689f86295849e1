"""
GenericRadialAtomicPotential with the trainable `nn` algorithm: the radial functions
H_k(r), k = 1..num_filters, are ONE small filter network r -> R^K shared by every
(centre, neighbour) element pair (reference: NNAlgorithm nn/atomic/grap.py:211-269,
`apply_model` :619-646 -> convolution1x1(variable_scope="Filters", output_bias=False),
convolutional.py:154-300; new mode only, grap.py:296-299).

Because the descriptor itself carries trainable variables, the loss gradients need
d G / d(filter weights) and, for the force / stress terms, its second-order companion.  The
path follows the EAM / ADP trainer (nn/eam/training.py): everything that touches the
neighbour lists stays in libtab200 --

    (i, j, D_p) = tab_pairs_export         one batch handle for all structures of the rank
    F, W        = PairForce(dE/dD_p)       tab_pair_forces (linear in dE/dD)
    backward    = tab_pair_jvp             its transpose

-- and the filter network, the moment sums P = sum_j H_k M_d, the T_dm contraction
(grap.py:470-535, 647-676) and the per-element atomic networks are evaluated by torch on the
same device so that autograd (create_graph) provides dE/dD_p and every parameter gradient,
which the reference obtains from TF second-order autograd (nn/opt.py:132-157).

`h_abck_modifier` 1 / 2 (filter input r / r_cov, exp(-r / r_cov) with the covalent radius of
the centre element; grap.py:621-632) use the Cordero-2008 table stated in `atoms.py` (the
source of ASE's `covalent_radii`; ASE itself is not installed here).
"""
import math
import os

import numpy as np
import torch

from tensoralloy_b200 import _lib
from tensoralloy_b200.nn import losses
from tensoralloy_b200.nn.atomic.training import VOIGT, AtomicNNTrainer, _activation



def filter_scope(nn):
    """`{scope}/Filters` -- the reference reads `Atomic/Filters/Conv3d{k}/kernel:0` when it
    exports the model (atomic.py:421-438)."""
    return f"{nn.scope}/Filters"

# unique Cartesian index tuples and multiplicities of T_dm (grap.py:470-512)
_AB = ((0, 0), (0, 1), (0, 2), (1, 1), (1, 2), (2, 2))
_AB_MULT = (1.0, 2.0, 2.0, 1.0, 2.0, 1.0)
_ABC = ((0, 0, 0), (0, 0, 1), (0, 0, 2), (0, 1, 1), (0, 1, 2), (0, 2, 2),
        (1, 1, 1), (1, 1, 2), (1, 2, 2), (2, 2, 2))
_ABC_MULT = (1.0, 3.0, 3.0, 3.0, 6.0, 3.0, 1.0, 3.0, 3.0, 1.0)


class NNAlgorithm:
    """Host description of the filter network (grap.py:211-269)."""
    name = "nn"

    def __init__(self, parameters=None):
        p = dict(parameters or {})
        self.use_resnet_dt = bool(p.get("use_resnet_dt", True))
        self.hidden_sizes = [int(x) for x in p.get("hidden_sizes", [32, 32, 32])]
        self.activation = p.get("activation", "softplus")
        self.num_filters = int(p.get("num_filters", 16))
        self.ckpt = p.get("ckpt", None)
        self.trainable = bool(p.get("trainable", True))
        self.h_abck_modifier = int(p.get("h_abck_modifier", 0))
        if isinstance(self.ckpt, str) and os.path.exists(self.ckpt):
            # grap.py:248-262: an npz checkpoint (a `use_fnn` model file) overrides the
            # architecture keys.  A path that no longer exists (a model file read back on
            # another machine: `as_dict` keeps the key) leaves the stored keys in force.
            npz = np.load(self.ckpt)
            actfn = {0: "relu", 1: "softplus", 2: "tanh", 3: "squareplus"}
            self.activation = actfn.get(int(npz["fnn::actfn"]))
            self.hidden_sizes = [int(x) for x in npz["fnn::layer_sizes"].tolist()]
            self.num_filters = self.hidden_sizes.pop(-1)
            self.use_resnet_dt = int(npz["fnn::use_resnet_dt"]) == 1
            if "fnn::trainable" in npz.files:
                self.trainable = int(npz["fnn::trainable"]) == 1
            if "fnn::h_abck_modifier" in npz.files:
                self.h_abck_modifier = int(npz["fnn::h_abck_modifier"])
        elif self.ckpt is not None and not isinstance(self.ckpt, str):
            raise ValueError("GRAP/nn: `ckpt` must be the path of an npz file or None")
        if self.h_abck_modifier not in (0, 1, 2):
            raise ValueError(f"GRAP/nn: unknown h_abck_modifier {self.h_abck_modifier} "
                             "(0: r, 1: r / r_cov, 2: exp(-r / r_cov))")
        if self.num_filters < 1 or not self.hidden_sizes:
            raise ValueError("GRAP/nn: num_filters >= 1 and at least one hidden layer")

    def __len__(self):
        return self.num_filters

    def as_dict(self):
        return {"use_resnet_dt": self.use_resnet_dt, "hidden_sizes": self.hidden_sizes,
                "activation": self.activation, "num_filters": self.num_filters,
                "trainable": self.trainable, "ckpt": self.ckpt,
                "h_abck_modifier": self.h_abck_modifier}


def initialize_filter_variables(nn, rng):
    """he_normal kernels / zero biases of `Filters/Conv3d{k}` and the bias-free
    `Filters/Output` layer (rank-5 input -> Conv3d names, convolutional.py:219,266-288)."""
    algo = nn.descriptor.algorithm_object
    sizes = [1] + list(algo.hidden_sizes)
    if isinstance(algo.ckpt, str):
        if not os.path.exists(algo.ckpt):
            raise FileNotFoundError(f"GRAP/nn: filter checkpoint {algo.ckpt} not found")
        # convolutional.py:219-254: constant initialisers from `fnn::weights_0_{j}` /
        # `fnn::biases_0_{j}` (NNAlgorithm took the architecture from the same file)
        npz = np.load(algo.ckpt)
        if not int(npz["use_fnn"]) if "use_fnn" in npz.files else True:
            raise ValueError(f"GRAP/nn: {algo.ckpt} holds no filter network (use_fnn = 0)")
        full = sizes + [algo.num_filters]
        for k in range(len(full) - 1):
            name = f"Conv3d{k + 1}" if k < len(full) - 2 else "Output"
            w = np.asarray(npz[f"fnn::weights_0_{k}"], dtype=np.float64).reshape(
                full[k], full[k + 1])
            nn.set_variable(f"{filter_scope(nn)}/{name}/kernel", w[None, None, None])
            if k < len(full) - 2:
                nn.set_variable(f"{filter_scope(nn)}/{name}/bias",
                                np.asarray(npz[f"fnn::biases_0_{k}"],
                                           dtype=np.float64).reshape(-1))
        return
    for k in range(len(sizes) - 1):
        w = np.clip(rng.normal(size=(sizes[k], sizes[k + 1])), -2.0, 2.0) * \
            np.sqrt(2.0 / sizes[k]) / 0.87962566103423978
        nn.set_variable(f"{filter_scope(nn)}/Conv3d{k + 1}/kernel", w[None, None, None])
        nn.set_variable(f"{filter_scope(nn)}/Conv3d{k + 1}/bias", np.zeros(sizes[k + 1]))
    w = np.clip(rng.normal(size=(sizes[-1], algo.num_filters)), -2.0, 2.0) * \
        np.sqrt(2.0 / sizes[-1]) / 0.87962566103423978
    nn.set_variable(f"{filter_scope(nn)}/Output/kernel", w[None, None, None])


def filter_params(nn):
    """The filter network as plain arrays (also what the oracle consumes)."""
    algo = nn.descriptor.algorithm_object
    W, b = [], []
    k = 1
    while f"{filter_scope(nn)}/Conv3d{k}/kernel" in nn.variables:
        w = nn.get_variable(f"{filter_scope(nn)}/Conv3d{k}/kernel")
        W.append(w.reshape(w.shape[-2], w.shape[-1]))
        b.append(nn.get_variable(f"{filter_scope(nn)}/Conv3d{k}/bias").reshape(-1))
        k += 1
    if not W:
        raise ValueError("GRAP/nn: the `Atomic/Filters/*` variables are not initialised")
    w = nn.get_variable(f"{filter_scope(nn)}/Output/kernel")
    W.append(w.reshape(w.shape[-2], w.shape[-1]))
    b.append(None)
    return dict(weights=W, biases=b, activation=algo.activation,
                use_resnet_dt=algo.use_resnet_dt)


def _cutoff(name, r, rc):
    """nn/cutoff.py:20-85."""
    z = torch.clamp(r / rc, max=1.0)
    if name == 'cosine':
        return 0.5 * (torch.cos(z * math.pi) + 1.0)
    return 1.0 + 5.0 * z ** 6 - 6.0 * z ** 5            # polynomial, gamma = 5


def filter_network(r, W, b, act, resnet):
    """convolution1x1 on the scalar input r: [P] -> [P, K] (no output bias)."""
    h = r[:, None]
    nh = len(W) - 1
    for k in range(nh):
        y = act(h @ W[k] + b[k])
        h = y + h if (k and resnet and W[k].shape[1] == W[k - 1].shape[1]) else y
    return h @ W[nh]


def closed_form_radial(algorithm, grid, rc):
    """f_tau(r) of the closed-form algorithms (grap.py:121-209) as one torch function
    r [P] -> [P, K]; `grid` = descriptor.radial_sets()."""
    def fn(r):
        cols = []
        for prm in grid:
            if algorithm == 'sf':
                cols.append(torch.exp(-prm[0] * (r - prm[1]) ** 2 / rc ** 2))
            elif algorithm == 'morse':
                d, g, r0 = prm
                cols.append(d * (torch.exp(-2.0 * g * (r - r0)) - 2.0 * torch.exp(-g * (r - r0))))
            elif algorithm == 'density':
                a, beta, re = prm
                cols.append(a * torch.exp(-beta * (r / re - 1.0)))
            elif algorithm == 'pexp':
                rl, pl = prm
                cols.append(torch.exp(-(r / rl) ** pl))
            else:
                raise ValueError(algorithm)
        return torch.stack(cols, dim=1)
    return fn


def centre_covalent_radii(elements, centre_types, dtype, device):
    """r_cov of the CENTRE atom of every pair (grap.py:623-628: the reference evaluates the
    filter input per centre element)."""
    from tensoralloy_b200.atoms import atomic_numbers, covalent_radii
    table = torch.tensor([covalent_radii[atomic_numbers[e]] for e in elements], dtype=dtype,
                         device=device)
    return table[centre_types]


def filter_descriptors(D, key, n_rows, radial, cutoff, rc, max_moment, symmetric, eps,
                       modifier=0, rcov=None):
    """New-mode GRAP descriptors from the directed pair vectors (grap.py:596-680).
    D [P, 3]; key [P] = centre * n_el + term (row of the moment sums); n_rows = n * n_el;
    radial: r [P] -> H [P, K] without the cutoff (the filter network or the closed forms).
    Returns [n_rows, K, max_moment + 1]; the caller reshapes to [n, n_el * K * (M + 1)].

    Moments <= 3 use the unique Cartesian index tuples with their multiplicities
    (`_get_moment_coeff_tensor` / `_get_multiplicity_tensor`, grap.py:470-535); a descriptor
    with max_moment 4 or 5 uses the full 3^m products with unit weights for EVERY moment and
    has no traceless form (`get_moment_tensor` / `get_T_dm`, grap.py:537-594, 655-660)."""
    r = torch.sqrt(torch.sum(D * D, dim=1) + eps)                 # universal.py:470-473
    # input of the radial functions (`h_abck_modifier`, grap.py:621-632): r, r / r_cov or
    # exp(-r / r_cov) with the covalent radius of the centre atom (`rcov` [P]); the cutoff and
    # the moments always see r itself
    h_in = r if modifier == 0 else (r / rcov if modifier == 1 else torch.exp(-r / rcov))
    H = radial(h_in) * _cutoff(cutoff, r, rc)[:, None]             # [P, K]
    K = H.shape[1]
    z = lambda *shape: torch.zeros(*shape, dtype=D.dtype, device=D.device)
    P0 = z(n_rows, K).index_add(0, key, H)
    cols = [torch.sign(P0) * torch.sqrt(P0 * P0 + 1e-16)]         # grap.py:667-676
    if max_moment == 0:
        return torch.stack(cols, dim=2)
    u = D / r[:, None]
    if max_moment > 3:
        m = torch.ones(D.shape[0], 1, dtype=D.dtype, device=D.device)
        for _ in range(max_moment):
            m = (m[:, :, None] * u[:, None, :]).reshape(D.shape[0], -1)       # [P, 3^k]
            P = z(n_rows, K, m.shape[1]).index_add(0, key, H[:, :, None] * m[:, None, :])
            cols.append(torch.sum(P * P, dim=2))
        return torch.stack(cols, dim=2)
    P1 = z(n_rows, K, 3).index_add(0, key, H[:, :, None] * u[:, None, :])
    S1 = torch.sum(P1 * P1, dim=2)
    cols.append(S1)
    if max_moment >= 2:
        m2 = torch.stack([u[:, a] * u[:, c] for a, c in _AB], dim=1)          # [P, 6]
        P2 = z(n_rows, K, 6).index_add(0, key, H[:, :, None] * m2[:, None, :])
        mult = torch.tensor(_AB_MULT, dtype=D.dtype, device=D.device)
        S2 = torch.sum(P2 * P2 * mult, dim=2)
        if symmetric:
            S2 = S2 - P0 * P0 / 3.0                                # grap.py:485-486
        cols.append(S2)
    if max_moment >= 3:
        m3 = torch.stack([u[:, a] * u[:, c] * u[:, d] for a, c, d in _ABC], dim=1)
        P3 = z(n_rows, K, 10).index_add(0, key, H[:, :, None] * m3[:, None, :])
        mult = torch.tensor(_ABC_MULT, dtype=D.dtype, device=D.device)
        S3 = torch.sum(P3 * P3 * mult, dim=2)
        if symmetric:
            S3 = S3 - 0.6 * S1                                     # grap.py:492-494
        cols.append(S3)
    return torch.stack(cols, dim=2)


def _modifier_of(nn):
    desc = nn.descriptor
    return int(desc.algorithm_object.h_abck_modifier) if desc.algorithm == 'nn' else 0


def _radial_of(nn, filters):
    """The radial part of the model's descriptor as a torch function of r."""
    desc = nn.descriptor
    if desc.algorithm == 'nn':
        return lambda r: filter_network(r, filters['W'], filters['b'], filters['act'],
                                        filters['resnet'])
    return closed_form_radial(desc.algorithm, desc.radial_sets(), nn.transformer.rcut)


class FilterEvaluator:
    """E, per-atom E, forces and virial of ONE structure (or one batch handle) for a model
    with the `nn` algorithm: the inference side of `AtomicNN._evaluate`
    (`TensorAlloyCalculator.calculate`).  Same ops as the trainer, parameters as constants:
    pair vectors from the library, filters + moment sums + atomic networks in torch on the
    device, forces / virial through `tab_pair_forces`."""

    def __init__(self, nn, device='cuda', pair_force=None):
        from tensoralloy_b200.precision import get_float_dtype
        self.nn, self.device = nn, device
        self.dt = get_float_dtype()
        self.tdtype = torch.float64 if self.dt.name == 'float64' else torch.float32
        t = lambda a: torch.tensor(np.asarray(a), dtype=self.tdtype, device=device)
        self.filters = None
        if nn.descriptor.algorithm == 'nn':
            fp = filter_params(nn)
            self.filters = dict(W=[t(w) for w in fp['weights']],
                                b=[None if v is None else t(v) for v in fp['biases']],
                                act=_activation(fp['activation']),
                                resnet=fp['use_resnet_dt'])
        self.layers = {}
        for el in nn.elements:
            p = nn.mlp_params(el)
            self.layers[el] = dict(
                W=[t(w) for w in p['weights']],
                b=[None if v is None else t(v) for v in p['biases']],
                act=_activation(p['activation']), resnet=p['use_resnet_dt'],
                xlo=None if p['xlo'] is None else t(p['xlo']),
                xhi=None if p['xhi'] is None else t(p['xhi']))
        if pair_force is None:
            from tensoralloy_b200.nn.eam.training import PairForce
            pair_force = PairForce.apply
        self._pair_force = pair_force

    _mlp = AtomicNNTrainer._mlp

    def __call__(self, nbr, types, pairs=None, want_forces=True):
        """types: element index per atom (caller order).  Returns (e_atom [N], F [N,3] or
        None, W [B,3,3] or None) as tensors of the working precision."""
        nn, desc = self.nn, self.nn.descriptor
        i, j, D = pairs if pairs is not None else nbr.pairs()
        i, j = i.long(), j.long()
        types = torch.as_tensor(np.asarray(types), device=self.device).long()
        n, nel = types.shape[0], len(nn.elements)
        ti, tj = types[i], types[j]
        term = torch.where(ti == tj, torch.zeros_like(ti), tj - (tj > ti).long() + 1)
        D = D.to(self.tdtype).detach().requires_grad_(want_forces)
        mod = _modifier_of(nn)
        G = filter_descriptors(D, i * nel + term, n * nel, _radial_of(nn, self.filters),
                               desc.cutoff_function, nn.transformer.rcut,
                               desc.max_moment, desc.is_T_symmetric, self.dt.eps, mod,
                               centre_covalent_radii(nn.elements, ti, self.tdtype, self.device)
                               if mod else None)
        G = G.reshape(n, -1)
        e_atom = torch.zeros(n, dtype=self.tdtype, device=self.device)
        for a, el in enumerate(nn.elements):
            sel = torch.nonzero(types == a).reshape(-1)
            if sel.numel():
                e_atom = e_atom.index_add(0, sel, self._mlp(el, G[sel]))
        if not want_forces:
            return e_atom.detach(), None, None
        g = torch.autograd.grad(e_atom.sum(), D)[0]
        forces, W = self._pair_force(g, nbr)
        return e_atom.detach(), forces, W


class GrapFilterTrainer(AtomicNNTrainer):
    """Training step (and, through `evaluate`, E / F / stress) of an AtomicNN over a GRAP
    descriptor with the `nn` algorithm.  Leaves: the per-element atomic networks (as in
    AtomicNNTrainer) plus the filter network (`filters['W']`, `filters['b']`)."""

    def __init__(self, nn, device='cuda', loss_weights=None, per_atom_energy=True,
                 pair_force=None):
        desc = nn.descriptor
        if not getattr(desc, 'uses_torch_path', lambda: False)():
            raise ValueError("GrapFilterTrainer serves GenericRadialAtomicPotential in new "
                             "mode with algorithm='nn' or moments 4 / 5; other descriptors "
                             "train through AtomicNNTrainer (descriptor kernels)")
        self._init_common(nn, device, loss_weights, per_atom_energy)
        self.filters = None
        if desc.algorithm == 'nn':
            fp = filter_params(nn)
            trainable = desc.algorithm_object.trainable
            mk = self._leaf if trainable else \
                (lambda a: torch.tensor(np.asarray(a), dtype=self.tdtype, device=device))
            self.filters = dict(W=[mk(w) for w in fp['weights']],
                                b=[None if v is None else mk(v) for v in fp['biases']],
                                act=_activation(fp['activation']),
                                resnet=fp['use_resnet_dt'])
        if pair_force is None:
            from tensoralloy_b200.nn.eam.training import PairForce
            pair_force = PairForce.apply
        self._pair_force = pair_force

    # -- data ------------------------------------------------------------------
    def _ensure_batch(self):
        if self._batch is not None:
            return self._batch
        clf = self.nn.transformer
        S = self.structures
        bf = clf.get_batch_features([s['atoms'] for s in S], nbr=_lib.NeighborList())
        i, j, D = bf.nbr.pairs()
        self._batch = self._make_batch(bf.nbr, bf.types, i.long(), j.long(), D)
        return self._batch

    def _make_batch(self, nbr, types, i, j, D):
        """types [N] (element index per atom, structures back to back), (i, j, D) the
        directed pairs in export order."""
        S = self.structures
        dev = self.device
        types = torch.as_tensor(np.asarray(types), device=dev).long()
        n_atoms = torch.tensor([s['n'] for s in S], device=dev)
        nel = len(self.elements)
        ti, tj = types[i], types[j]
        # index of the k-body term inside the centre's list [cc, c-x1, ...] (utils.py:262-273)
        term = torch.where(ti == tj, torch.zeros_like(ti), tj - (tj > ti).long() + 1)
        return dict(
            nbr=nbr, key=i * nel + term, D=D.to(self.tdtype), types=types, centre=ti,
            sel=[torch.nonzero(types == a).reshape(-1) for a in range(nel)],
            sid=torch.repeat_interleave(torch.arange(len(S), device=dev), n_atoms),
            n_atoms=n_atoms,
            volume=torch.tensor([s['volume'] for s in S], dtype=self.tdtype, device=dev),
            energy=torch.stack([s['energy'] for s in S]),
            forces=torch.cat([s['forces'] for s in S]),
            stress=torch.stack([s['stress'] for s in S]))

    # -- model -------------------------------------------------------------------
    def descriptors(self, D):
        B = self._batch
        desc = self.nn.descriptor
        n, nel = B['types'].shape[0], len(self.elements)
        mod = _modifier_of(self.nn)
        G = filter_descriptors(D, B['key'], n * nel, _radial_of(self.nn, self.filters),
                               desc.cutoff_function, self.nn.transformer.rcut,
                               desc.max_moment, desc.is_T_symmetric, self.dt.eps, mod,
                               centre_covalent_radii(self.elements, B['centre'], self.tdtype,
                                                     self.device) if mod else None)
        return G.reshape(n, -1)                   # [n, term * K * (M + 1)]

    def energies(self, D):
        """Per-structure energies as a function of the pair vectors."""
        B = self._batch
        G = self.descriptors(D)
        e_atom = torch.zeros(G.shape[0], dtype=self.tdtype, device=self.device)
        for a, el in enumerate(self.elements):
            sel = B['sel'][a]
            if sel.numel():
                e_atom = e_atom.index_add(0, sel, self._mlp(el, G[sel]))
        E = torch.zeros(len(self.structures), dtype=self.tdtype,
                        device=self.device).index_add(0, B['sid'], e_atom)
        return E, e_atom

    def total_loss(self, want_forces=True, want_stress=True):
        B = self._ensure_batch()
        D = B['D'].detach().requires_grad_(True)
        E, _ = self.energies(D)
        w = self.loss_weights
        loss = losses.energy_loss(B['energy'], E, B['n_atoms'], self.per_atom_energy,
                                  w['energy'])
        parts = {'energy': loss.detach()}
        if want_forces or want_stress:
            g = torch.autograd.grad(E.sum(), D, create_graph=True)[0]
            F, W = self._pair_force(g, B['nbr'])
            if want_forces:
                lf = losses.forces_loss(B['forces'], F, w['forces'])
                loss = loss + lf
                parts['forces'] = lf.detach()
            if want_stress:
                st = W / B['volume'][:, None, None]
                voigt = torch.stack([st[:, a, b] for a, b in VOIGT], dim=1)
                ls = losses.stress_loss(B['stress'], voigt, w['stress'])
                loss = loss + ls
                parts['stress'] = ls.detach()
        return loss, parts

    def evaluate(self):
        """Energies [B], forces [N, 3] and Voigt stresses [B, 6] of the queued structures
        with the current parameters (no labels needed beyond placeholders)."""
        B = self._ensure_batch()
        D = B['D'].detach().requires_grad_(True)
        E, _ = self.energies(D)
        g = torch.autograd.grad(E.sum(), D)[0]
        F, W = self._pair_force(g, B['nbr'])
        st = W / B['volume'][:, None, None]
        voigt = torch.stack([st[:, a, b] for a, b in VOIGT], dim=1)
        return E.detach(), F.detach(), voigt.detach()

    def sync_to_model(self):
        """Write the trained values back into the model's variables."""
        nn = self.nn
        for el in self.elements:
            L = self.layers[el]
            nh = len(L['W']) - 1
            for k, w in enumerate(L['W']):
                name = f"{nn.scope}/{el}/" + (f"Conv1d{k + 1}" if k < nh else "Output")
                nn.set_variable(f"{name}/kernel", w.detach().cpu().numpy()[None])
                if L['b'][k] is not None:
                    nn.set_variable(f"{name}/bias", L['b'][k].detach().cpu().numpy())
        if self.filters is None:
            return
        nh = len(self.filters['W']) - 1
        for k, w in enumerate(self.filters['W']):
            name = f"{filter_scope(nn)}/" + (f"Conv3d{k + 1}" if k < nh else "Output")
            nn.set_variable(f"{name}/kernel", w.detach().cpu().numpy()[None, None, None])
            if self.filters['b'][k] is not None:
                nn.set_variable(f"{name}/bias", self.filters['b'][k].detach().cpu().numpy())
