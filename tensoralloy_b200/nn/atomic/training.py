"""
Training step of AtomicNN (energy + force + stress loss with parameter
gradients) -- the hot part of the reference's `BasicNN.model_fn` in TRAIN mode
(nn/basic.py:920-1015: build -> get_total_loss :446-631 -> get_train_op
nn/opt.py:89-166), structure-parallel over the GPUs with one flat gradient
all-reduce (the reference's MirroredStrategy with MEAN variable aggregation,
train/distribute_utils.py:84-159, potentials.py:41, atomic.py:155).

"PyTorch custom ops where tensors cross into training": the geometry work stays
in libtab200 --
    G        = descriptors (tab_atomic_descriptors; constant w.r.t. parameters)
    F, W     = SfForce(dE/dG)     (tab_atomic_forces: linear in dE/dG)
    backward = tab_atomic_jvp     (the transpose of that linear map)
-- while the tiny per-element MLPs run in torch so that autograd provides
dE/dG (create_graph) and the parameter gradients of the force / stress terms,
which the reference obtains from TF second-order autograd.
"""
import numpy as np
import torch

from tensoralloy_b200 import _lib
from tensoralloy_b200.nn import losses
from tensoralloy_b200.precision import get_float_dtype

VOIGT = ((0, 0), (1, 1), (2, 2), (1, 2), (0, 2), (0, 1))


class SfForce(torch.autograd.Function):
    """forces [N,3] and virials [B,3,3] of a BATCH of structures (one neighbour handle,
    `tab_nbr_build_batch`) from c = dE/dG [N, dim]: one kernel pass over all structures."""

    @staticmethod
    def forward(ctx, dedg, model, nbr, precision):
        n = dedg.shape[0]
        nb = max(nbr.n_struct, 1)
        c = dedg.detach().to(torch.float64).contiguous()
        forces = torch.empty((n, 3), dtype=torch.float64, device=c.device)
        virial = torch.empty(nb * 9, dtype=torch.float64, device=c.device)
        model.forces_from_dedg(nbr, c, forces, virial, precision)
        ctx.model, ctx.nbr, ctx.precision = model, nbr, precision
        ctx.shape, ctx.dtype = dedg.shape, dedg.dtype
        return forces.to(dedg.dtype), virial.reshape(nb, 3, 3).to(dedg.dtype)

    @staticmethod
    def backward(ctx, g_forces, g_virial):
        u = g_forces.detach().to(torch.float64).contiguous()
        A = g_virial.detach().to(torch.float64).contiguous().reshape(-1)
        out = torch.empty(ctx.shape, dtype=torch.float64, device=u.device)
        ctx.model.jvp(ctx.nbr, u, A, out, ctx.precision)
        return out.to(ctx.dtype), None, None, None


def allreduce_mean_(params, dist, world):
    """Average the gradients of `params` over the ranks with ONE flat-buffer
    all-reduce (reference: MirroredStrategy + VariableAggregation.MEAN,
    train/distribute_utils.py:84-159, nn/eam/potentials/potentials.py:41)."""
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1)
                      for p in params])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat /= world
    off = 0
    for p in params:
        n = p.numel()
        if p.grad is None:
            p.grad = flat[off:off + n].reshape(p.shape).clone()
        else:       # in place: a captured step keeps writing into the same tensors
            p.grad.copy_(flat[off:off + n].reshape(p.shape))
        off += n
    return flat


def _activation(name):
    name = name.lower()
    F = torch.nn.functional
    table = {'softplus': F.softplus, 'tanh': torch.tanh, 'relu': torch.relu,
             'leaky_relu': lambda x: F.leaky_relu(x, 0.2), 'sigmoid': torch.sigmoid,
             'softsign': F.softsign, 'elu': F.elu,
             'squareplus': lambda x: 0.5 * (x + torch.sqrt(x * x + 4.0))}
    return table[name]


class AtomicNNTrainer:
    """Holds the structures of this rank (lists built once, descriptors cached:
    the geometry does not change during training) and the torch parameters."""

    def __init__(self, nn, device='cuda', loss_weights=None, per_atom_energy=True):
        self.model = nn._device_model()          # geometry side (weights unused)
        self._init_common(nn, device, loss_weights, per_atom_energy)

    def _init_common(self, nn, device, loss_weights, per_atom_energy):
        self.nn = nn
        self.device = device
        self.dt = get_float_dtype()
        self.tdtype = torch.float64 if self.dt.name == 'float64' else torch.float32
        self.elements = nn.elements
        self.loss_weights = dict(energy=1.0, forces=1.0, stress=1.0)
        self.loss_weights.update(loss_weights or {})
        self.per_atom_energy = per_atom_energy
        self.params = []        # flat list of leaf tensors
        self.layers = {}        # el -> dict(W=[...], b=[...], out_b, xlo, xhi)
        for el in self.elements:
            p = nn.mlp_params(el)
            W = [self._leaf(w) for w in p['weights']]
            b = [self._leaf(v) if v is not None else None for v in p['biases']]
            self.layers[el] = dict(
                W=W, b=b, act=_activation(p['activation']), resnet=p['use_resnet_dt'],
                xlo=None if p['xlo'] is None else torch.tensor(p['xlo'], dtype=self.tdtype,
                                                               device=device),
                xhi=None if p['xhi'] is None else torch.tensor(p['xhi'], dtype=self.tdtype,
                                                               device=device))
        self.structures = []
        self._batch = None

    def _leaf(self, arr):
        t = torch.tensor(np.asarray(arr), dtype=self.tdtype, device=self.device,
                         requires_grad=True)
        self.params.append(t)
        return t

    # -- data ------------------------------------------------------------------
    def add_structure(self, atoms, energy, forces, stress):
        """Queue one labelled structure of this rank's sub-batch."""
        t = lambda a: torch.tensor(np.asarray(a), dtype=self.tdtype, device=self.device)
        self.structures.append(dict(atoms=atoms, n=len(atoms),
                                    volume=float(atoms.get_volume()), energy=t(energy),
                                    forces=t(forces), stress=t(stress)))
        self._batch = None
        self._graph = None      # a captured step belongs to the old batch

    def _ensure_batch(self):
        """Lists + descriptors of ALL queued structures in one device handle (kept: the
        geometry does not change during training)."""
        if self._batch is not None:
            return self._batch
        clf = self.nn.transformer
        bf = clf.get_batch_features([s['atoms'] for s in self.structures],
                                    rc=self.nn.required_cutoff(), nbr=_lib.NeighborList())
        G = self.model.descriptors(bf.nbr, self.dt.tab_precision).to(self.tdtype)
        dev = self.device
        n_atoms = torch.tensor([s['n'] for s in self.structures], device=dev)
        self._batch = dict(
            nbr=bf.nbr, G=G, types=torch.as_tensor(bf.types, device=dev).long(),
            sel=[torch.nonzero(torch.as_tensor(bf.types, device=dev) == a).reshape(-1)
                 for a in range(len(self.elements))],
            sid=torch.repeat_interleave(torch.arange(len(self.structures), device=dev),
                                        n_atoms),
            n_atoms=n_atoms,
            volume=torch.tensor([s['volume'] for s in self.structures], dtype=self.tdtype,
                                device=dev),
            energy=torch.stack([s['energy'] for s in self.structures]),
            forces=torch.cat([s['forces'] for s in self.structures]),
            stress=torch.stack([s['stress'] for s in self.structures]))
        return self._batch

    # -- model -------------------------------------------------------------------
    def _mlp(self, el, x):
        L = self.layers[el]
        if L['xlo'] is not None:
            den = L['xhi'] - L['xlo']
            x = torch.where(den == 0, torch.zeros_like(x), (L['xhi'] - x) / den)
        h = x
        nh = len(L['W']) - 1
        for k in range(nh):
            y = L['act'](h @ L['W'][k] + L['b'][k])
            if k and L['resnet'] and L['W'][k].shape[1] == L['W'][k - 1].shape[1]:
                h = y + h
            else:
                h = y
        out = h @ L['W'][nh]
        if L['b'][nh] is not None:
            out = out + L['b'][nh]
        return out[:, 0]

    def total_loss(self, want_forces=True, want_stress=True):
        """Loss of this rank's structures (reference semantics: each replica
        evaluates the loss of its own sub-batch)."""
        B = self._ensure_batch()
        nb = len(self.structures)
        G_all = B['G'].detach().requires_grad_(True)
        e_atom = torch.zeros(G_all.shape[0], dtype=self.tdtype, device=self.device)
        for a, el in enumerate(self.elements):
            sel = B['sel'][a]
            if sel.numel():
                e_atom = e_atom.index_add(0, sel, self._mlp(el, G_all[sel]))
        E = torch.zeros(nb, dtype=self.tdtype, device=self.device).index_add(
            0, B['sid'], e_atom)
        w = self.loss_weights
        loss = losses.energy_loss(B['energy'], E, B['n_atoms'], self.per_atom_energy,
                                  w['energy'])
        parts = {'energy': loss.detach()}
        if want_forces or want_stress:
            dedg = torch.autograd.grad(E.sum(), G_all, create_graph=True)[0]
            F, W = SfForce.apply(dedg, self.model, B['nbr'], self.dt.tab_precision)
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

    # -- one optimisation step ---------------------------------------------------
    def gradients(self, want_forces=True, want_stress=True):
        for p in self.params:
            p.grad = None
        loss, parts = self.total_loss(want_forces, want_stress)
        loss.backward()
        return loss.detach(), parts

    def allreduce_gradients(self, dist, world):
        allreduce_mean_(self.params, dist, world)

    def enable_graph(self, warmup=3):
        """Capture loss + backward of this rank's (fixed) batch in ONE CUDA graph: the
        training step is launch-bound (hundreds of small torch kernels around the force
        operator and its JVP), the geometry and all shapes are static.  `train_step`
        then replays the graph; gradients land in the same tensors every step.  Returns
        True when the capture worked (otherwise the eager path stays in use)."""
        self._graph = None
        try:
            self._ensure_batch()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(warmup):
                    for p in self.params:
                        p.grad = None
                    loss, _ = self.total_loss()
                    loss.backward()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            for p in self.params:
                p.grad = None
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                loss, parts = self.total_loss()
                loss.backward()
            # the replay writes into the gradient tensors allocated during the capture: keep
            # them, so that `train_step` can re-attach them if a caller detached them
            # (`optimizer.zero_grad()` sets p.grad = None by default)
            self._graph = (g, loss, parts, [p.grad for p in self.params])
        except Exception as exc:      # capture is an optimisation, never a requirement
            self._graph = None
            self.graph_error = f"{type(exc).__name__}: {exc}"
            for p in self.params:
                p.grad = None
        return self._graph is not None

    def train_step(self, optimizer, dist=None, world=1):
        graph = getattr(self, '_graph', None)
        if graph is not None:
            for p, g in zip(self.params, graph[3]):
                if p.grad is not g:          # detached or replaced since the capture
                    p.grad = g
            graph[0].replay()
            loss, parts = graph[1].detach(), graph[2]
        else:
            loss, parts = self.gradients()
        if dist is not None and world > 1:
            self.allreduce_gradients(dist, world)
        optimizer.step()
        return loss, parts

    def sync_to_model(self):
        """Write the trained parameters back into the AtomicNN variables."""
        for el in self.elements:
            L = self.layers[el]
            nh = len(L['W']) - 1
            for k in range(nh):
                self.nn.set_variable(f"{self.nn.scope}/{el}/Conv1d{k + 1}/kernel",
                                     L['W'][k].detach().cpu().numpy()[None])
                self.nn.set_variable(f"{self.nn.scope}/{el}/Conv1d{k + 1}/bias",
                                     L['b'][k].detach().cpu().numpy())
            self.nn.set_variable(f"{self.nn.scope}/{el}/Output/kernel",
                                 L['W'][nh].detach().cpu().numpy()[None])
            if L['b'][nh] is not None:
                self.nn.set_variable(f"{self.nn.scope}/{el}/Output/bias",
                                     L['b'][nh].detach().cpu().numpy())


class TemperatureDependentTrainer(AtomicNNTrainer):
    """Training step of `TemperatureDependentAtomicNN` / `BeNN`
    (nn/atomic/finite_temperature.py:211-304,338-366): the energy losses cover the
    properties listed in `minimize_properties` among 'energy' (U), 'free_energy' (F = U - T S)
    and 'eentropy' (S); forces and stress derive from the FREE energy (basic.py:190-202).
    Geometry side as in `AtomicNNTrainer`: descriptors, force operator and its JVP are the
    libtab200 kernels on one batch handle; the H / S / U heads run in torch."""

    def __init__(self, nn, device='cuda', loss_weights=None, per_atom_energy=True):
        self.nn = nn
        self.device = device
        self.dt = get_float_dtype()
        self.tdtype = torch.float64 if self.dt.name == 'float64' else torch.float32
        self.model = nn._device_model()
        self.elements = nn.elements
        self.loss_weights = dict(energy=1.0, free_energy=1.0, eentropy=1.0, forces=1.0,
                                 stress=1.0)
        self.loss_weights.update(loss_weights or {})
        self.per_atom_energy = per_atom_energy
        self.params = []
        self.named = {}             # reference variable name -> leaf
        self.heads = {}
        for el in self.elements:
            hd = {}
            for head in ('H', 'S', 'U'):
                p = nn.head_params(el, head)
                base = f"{nn.scope}/{el}/{head}"
                W, b = [], []
                for k, w in enumerate(p['weights'][:-1]):
                    W.append(self._named_leaf(f"{base}/Conv1d{k + 1}/kernel", w))
                    b.append(self._named_leaf(f"{base}/Conv1d{k + 1}/bias", p['biases'][k]))
                W.append(self._named_leaf(f"{base}/Output/kernel", p['weights'][-1]))
                ob = None if p['out_bias'] is None else \
                    self._named_leaf(f"{base}/Output/bias", p['out_bias'])
                hd[head] = dict(W=W, b=b, ob=ob, act=p['activation'],
                                resnet=p['use_resnet_dt'])
            mm = nn.minmax(el)
            t = lambda a: torch.tensor(np.asarray(a), dtype=self.tdtype, device=device)
            hd['xlo'], hd['xhi'] = (t(mm[0]), t(mm[1])) if mm else (None, None)
            self.heads[el] = hd
        self.structures = []
        self._batch = None

    def _named_leaf(self, name, arr):
        t = self._leaf(arr)
        self.named[name] = t
        return t

    def add_structure(self, atoms, energy, forces, stress, free_energy=None, eentropy=None):
        super().add_structure(atoms, energy, forces, stress)
        t = lambda a: torch.tensor(float(a), dtype=self.tdtype, device=self.device)
        s = self.structures[-1]
        s['free_energy'] = t(energy if free_energy is None else free_energy)
        s['eentropy'] = t(0.0 if eentropy is None else eentropy)
        s['etemperature'] = float(atoms.info.get('etemperature', 0.0))

    def total_loss(self, want_forces=True, want_stress=True):
        B = self._ensure_batch()
        S_ = self.structures
        nb = len(S_)
        if 't_atom' not in B:
            temps = torch.tensor([s['etemperature'] for s in S_], dtype=self.tdtype,
                                 device=self.device)
            B['t_atom'] = temps[B['sid']]
            B['free_energy'] = torch.stack([s['free_energy'] for s in S_])
            B['eentropy'] = torch.stack([s['eentropy'] for s in S_])
        G_all = B['G'].detach().requires_grad_(True)
        n = G_all.shape[0]
        z = lambda: torch.zeros(n, dtype=self.tdtype, device=self.device)
        U, S = z(), z()
        net = self.nn._net
        for a, el in enumerate(self.elements):
            sel = B['sel'][a]
            if not sel.numel():
                continue
            hd = self.heads[el]
            x = G_all[sel]
            if hd['xlo'] is not None:
                den = hd['xhi'] - hd['xlo']
                x = torch.where(den == 0, torch.zeros_like(x), (hd['xhi'] - x) / den)
            T = B['t_atom'][sel]
            Ht = torch.cat([net(hd['H'], x), T[:, None]], dim=1)
            U = U.index_add(0, sel, net(hd['U'], Ht)[:, 0])
            S = S.index_add(0, sel, self.nn._entropy(hd, Ht, T))
        F = U - B['t_atom'] * S
        seg = lambda v: torch.zeros(nb, dtype=self.tdtype, device=self.device).index_add(
            0, B['sid'], v)
        totals = {'energy': seg(U), 'eentropy': seg(S), 'free_energy': seg(F)}
        w = self.loss_weights
        loss = torch.zeros((), dtype=self.tdtype, device=self.device)
        parts = {}
        for prop in ('energy', 'free_energy', 'eentropy'):
            if prop in self.nn.minimize_properties:
                lp = losses.energy_loss(B[prop], totals[prop], B['n_atoms'],
                                        self.per_atom_energy, w[prop])
                loss = loss + lp
                parts[prop] = lp.detach()
        if want_forces or want_stress:
            dedg = torch.autograd.grad(totals['free_energy'].sum(), G_all,
                                       create_graph=True)[0]
            Fo, W = SfForce.apply(dedg, self.model, B['nbr'], self.dt.tab_precision)
            if want_forces:
                lf = losses.forces_loss(B['forces'], Fo, w['forces'])
                loss = loss + lf
                parts['forces'] = lf.detach()
            if want_stress:
                st = W / B['volume'][:, None, None]
                voigt = torch.stack([st[:, a, b] for a, b in VOIGT], dim=1)
                ls = losses.stress_loss(B['stress'], voigt, w['stress'])
                loss = loss + ls
                parts['stress'] = ls.detach()
        return loss, parts

    def sync_to_model(self):
        for name, t in self.named.items():
            v = t.detach().cpu().numpy()
            self.nn.set_variable(name, v[None] if name.endswith('/kernel') else v)
