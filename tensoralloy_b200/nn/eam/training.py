"""
Training step of the EAM / ADP family (energy + force + stress loss with parameter
gradients): the hot part of the reference's `BasicNN.model_fn` in TRAIN mode for
`EamAlloyNN` / `EamFsNN` / `AdpNN` (nn/basic.py:920-1015 -> get_total_loss :446-631 ->
get_train_op nn/opt.py:89-166; model graph nn/eam/eam.py:495-570, alloy.py:128-196,
fs.py:146-203, adp.py:315-586), structure-parallel over the GPUs with one flat gradient
all-reduce (MirroredStrategy + MEAN aggregation, potentials.py:41).

"PyTorch custom ops where tensors cross into training": everything that touches the
neighbour lists stays in libtab200 --
    (i, j, D_p)  = tab_pairs_export     one batch handle for all structures of the rank
    F, W         = PairForce(dE/dD_p)   tab_pair_forces (linear in dE/dD)
    backward     = tab_pair_jvp         its transpose
-- while the scalar functions of the pair table (rho(r), phi(r), F(rho), u(r), w(r): the
trainable 'nn' MLPs of eam.py:174-190 and the zjw04 forms with their shared empirical
variables, potentials.py:171-200) are evaluated by torch so that autograd provides dE/dD
(create_graph) and the parameter gradients of the force / stress terms, which the
reference obtains from TF second-order autograd.
"""
import numpy as np
import torch

from tensoralloy_b200 import _lib
from tensoralloy_b200.nn import losses
from tensoralloy_b200.nn.atomic.training import _activation, allreduce_mean_
from tensoralloy_b200.precision import get_float_dtype
from tensoralloy_b200.utils import get_elements_from_kbody_term

VOIGT = ((0, 0), (1, 1), (2, 2), (1, 2), (0, 2), (0, 1))


class PairForce(torch.autograd.Function):
    """forces [N,3] and virials [B,3,3] from g = dE/dD [nij,3] (export order)."""

    @staticmethod
    def forward(ctx, g, nbr):
        nb = max(nbr.n_struct, 1)
        c = g.detach().to(torch.float64).contiguous()
        forces = torch.empty((nbr.n, 3), dtype=torch.float64, device=c.device)
        virial = torch.empty(nb * 9, dtype=torch.float64, device=c.device)
        nbr.pair_forces(c, forces, virial)
        ctx.nbr, ctx.shape, ctx.dtype = nbr, g.shape, g.dtype
        return forces.to(g.dtype), virial.reshape(nb, 3, 3).to(g.dtype)

    @staticmethod
    def backward(ctx, g_forces, g_virial):
        u = g_forces.detach().to(torch.float64).contiguous()
        A = g_virial.detach().to(torch.float64).contiguous().reshape(-1)
        out = torch.empty(ctx.shape, dtype=torch.float64, device=u.device)
        ctx.nbr.pair_jvp(u, A, out)
        return out.to(ctx.dtype), None


# ---------------------------------------------------------------------------------------
# torch forms of the trainable functions
# ---------------------------------------------------------------------------------------
def _zhou_exp(r, a, b, c, re):
    """generic.py:102-117: a exp(-b (r/re - 1)) / (1 + (r/re - c)^20)."""
    x = r / re
    return a * torch.exp(-b * (x - 1.0)) / (1.0 + (x - c) ** 20)


class _Functions:
    """Callable table of one model: `fn(kind, key)(x)`; owns the leaf tensors."""

    ZJW = ('zjw04', 'zjw04xc', 'zjw04uxc', 'zjw04xcp')

    def __init__(self, nn, tdtype, device, freeze_reference_fixed=True):
        """freeze_reference_fixed: keep the parameters a potential class declares as
        never trained (`fixed_parameters`, potentials.py:76-79,119-127 -- e.g. the
        embedding parameters and r_eq of zjw04) out of the optimiser, as the reference
        does.  False makes every shared variable of a non-fixed function a trainable
        leaf (gradient checks)."""
        self.freeze = freeze_reference_fixed
        self.nn = nn
        self.tdtype = tdtype
        self.device = device
        self.named = {}          # reference variable name -> leaf tensor
        self._nn_layers = {}     # (fn, key) -> list of leaves [W0, b0, ..., Wout]
        self._shared = {}        # (section, key) -> leaf (empirical shared variables)
        self.act = _activation(nn._activation)

    def params(self):
        return [t for t in self.named.values() if t.requires_grad]

    def _fixed(self, fn, section):
        return f"{section}.{fn}" in self.nn._fixed_functions

    # -- 'nn' ---------------------------------------------------------------------------
    def _mlp(self, fn, section):
        tag = (fn, section)
        if tag not in self._nn_layers:
            provider = self.nn._nn
            arrays = provider.weights(fn, section)
            names = list(k for k in provider.variables()
                         if k.startswith(provider.prefix(fn, section) + '/'))
            rank = 1 if fn == 'embed' else 2
            pre = provider.prefix(fn, section)
            order = []
            nh = (len(arrays) - 1) // 2
            for k in range(nh):
                order += [f"{pre}/Conv{rank}d{k + 1}/kernel", f"{pre}/Conv{rank}d{k + 1}/bias"]
            order.append(f"{pre}/Output/kernel")
            assert set(order) == set(names)
            leaves = []
            for name, arr in zip(order, arrays):
                t = torch.tensor(np.asarray(arr), dtype=self.tdtype, device=self.device,
                                 requires_grad=not self._fixed(fn, section))
                self.named[name] = t
                leaves.append(t)
            self._nn_layers[tag] = leaves
        L = self._nn_layers[tag]

        def call(x):
            h = x[:, None]
            for k in range((len(L) - 1) // 2):
                h = self.act(h @ L[2 * k] + L[2 * k + 1])
            return (h @ L[-1])[:, 0]          # linear output unit, no bias (eam.py:174-190)
        return call

    # -- zjw04 --------------------------------------------------------------------------
    def _p(self, potential, section, key, fixed):
        tag = (section, key)
        pot = self.nn._empirical_functions[potential]
        if self.freeze and key in pot.fixed_parameters.get(section, ()):
            fixed = True
        if tag not in self._shared:
            value = pot.params[section][key]
            t = torch.tensor(float(value), dtype=self.tdtype, device=self.device,
                             requires_grad=not fixed)
            self._shared[tag] = t
            self.named[f"{self.nn.scope}/Shared/{section}/{key}"] = t
        elif not fixed and not self._shared[tag].requires_grad:
            self._shared[tag].requires_grad_(True)      # shared with a trainable function
        return self._shared[tag]

    # (the leaves are created when the function table is built, not at the first call:
    #  `EamTrainer.params` must list every trainable variable before the first step)
    def _zjw_rho(self, pot, el, fixed):
        fe, be, la, re = (self._p(pot, el, k, fixed) for k in ('f_eq', 'beta', 'lamda', 'r_eq'))
        return lambda r: _zhou_exp(r, fe, be, la, re)

    def _zjw_phi(self, pot, term, fixed):
        a, b = get_elements_from_kbody_term(term)

        def same(el):
            A, al, ka, B, be, la, re = (self._p(pot, el, k, fixed) for k in
                                        ('A', 'alpha', 'kappa', 'B', 'beta', 'lamda', 'r_eq'))
            return lambda r: _zhou_exp(r, A, al, ka, re) - _zhou_exp(r, B, be, la, re)
        if a == b:
            return same(a)
        pa, pb, ra, rb = same(a), same(b), self._zjw_rho(pot, a, fixed), \
            self._zjw_rho(pot, b, fixed)
        # zjw04.py:229-243
        return lambda r: 0.5 * (ra(r) / rb(r) * pb(r) + rb(r) / ra(r) * pa(r))

    def _zjw_phi_xcp(self, pot, term, fixed):
        # zjw04.py:570-696: the A-B pair has its own parameter section
        a, b = get_elements_from_kbody_term(term)
        params = self.nn._empirical_functions[pot].params
        sec = a if a == b else (term if term in params else f"{b}{a}")
        A, al, ka, B, be, la, re = (self._p(pot, sec, k, fixed) for k in
                                    ('A', 'alpha', 'kappa', 'B', 'beta', 'lamda', 'r_eq'))
        return lambda r: _zhou_exp(r, A, al, ka, re) - _zhou_exp(r, B, be, la, re)

    def _zjw_embed(self, pot, el, fixed):
        keys = ('Fn0', 'Fn1', 'Fn2', 'Fn3', 'F0', 'F1', 'F2', 'F3', 'eta', 'Fe', 'rho_e',
                'rho_s')
        Fn0, Fn1, Fn2, Fn3, F0, F1, F2, F3, eta, Fe, rho_e, rho_s = (
            self._p(pot, el, k, fixed) for k in keys)

        def blended(rho):
            # zjw04.py:440-550: the three branches blended by sigmoids
            rho_n, rho_0 = 0.85 * rho_e, 1.15 * rho_e
            x1 = rho / rho_n - 1.0
            e1 = Fn0 + Fn1 * x1 + Fn2 * x1 ** 2 + Fn3 * x1 ** 3
            x2 = rho / rho_e - 1.0
            e2 = F0 + F1 * x2 + F2 * x2 ** 2 + F3 * x2 ** 3
            x3 = rho / rho_s + 1e-8
            e3 = Fe * (1.0 - eta * torch.log(x3)) * x3 ** eta
            c1 = torch.sigmoid(2.0 * (rho_n - rho))
            c3 = torch.sigmoid(2.0 * (rho - rho_0))
            return c1 * e1 + (1.0 - (c1 + c3)) * e2 + c3 * e3
        if pot != 'zjw04':
            return blended

        def call(rho):
            # zjw04.py:279-389 (three branches selected by rho)
            rho_n, rho_0 = 0.85 * rho_e, 1.15 * rho_e
            x1 = rho / rho_n - 1.0
            e1 = Fn0 + Fn1 * x1 + Fn2 * x1 ** 2 + Fn3 * x1 ** 3
            x2 = rho / rho_e - 1.0
            e2 = F0 + F1 * x2 + F2 * x2 ** 2 + F3 * x2 ** 3
            safe = torch.where(rho >= rho_0, rho, torch.ones_like(rho) * rho_0.detach())
            x3 = safe / rho_s
            e3 = Fe * (1.0 - eta * torch.log(x3)) * x3 ** eta
            return torch.where(rho < rho_n, e1, torch.where(rho < rho_0, e2, e3))
        return call

    # -- further empirical forms (same formulas as csrc/potentials.cuh) -------------------
    def _pair_section(self, pot, term):
        a, b = get_elements_from_kbody_term(term)
        params = self.nn._empirical_functions[pot].params
        return f"{a}{b}" if f"{a}{b}" in params else f"{b}{a}"

    def _sutton(self, fn, section, fixed):
        # sutton90.py:43-97
        if fn == 'rho':
            el = get_elements_from_kbody_term(section)[-1]
            a = self._p('sutton90', el, 'a', fixed)
            return lambda r: (a / r) ** 6
        if fn == 'phi':
            b = self._p('sutton90', self._pair_section('sutton90', section), 'b', fixed)
            return lambda r: (b / r) ** 12
        return lambda rho: -torch.sqrt(rho)

    def _agrawal(self, fn, section, fixed):
        # agrawal.py:57-152
        el = get_elements_from_kbody_term(section)[-1] if fn == 'rho' else \
            get_elements_from_kbody_term(section)[0]
        P = lambda k: self._p('Be/1', el, k, fixed)
        if fn == 'rho':
            A, B, re, rc, m = P('A'), P('B'), P('re'), P('rc'), P('m')

            def rho(r):
                tail = A * torch.exp(-B * (rc - re))
                return A * torch.exp(-B * (r - re)) - tail + \
                    rc / m * (1.0 - (r / rc) ** m) * (-B * tail)
            return rho
        if fn == 'phi':
            D, al, re, rc, m = P('D'), P('alpha'), P('re'), P('rc'), P('m')
            morse = lambda x: D * (torch.exp(-2.0 * al * (x - re)) -
                                   2.0 * torch.exp(-al * (x - re)))
            dmorse = lambda x: (torch.exp(-al * (x - re)) -
                                torch.exp(-2.0 * al * (x - re))) * (D * al * 2.0)
            return lambda r: morse(r) - morse(rc) + rc / m * ((1.0 - (r / rc) ** m) *
                                                              dmorse(rc))
        F0, F1, be, ga = P('F0'), P('F1'), P('beta'), P('gamma')
        return lambda rho: F0 * (1.0 - be * torch.log(torch.clamp(rho, min=1e-12))) * \
            rho ** be + F1 * rho ** ga

    def _grimes(self, fn, section, fixed):
        # grimmes.py:41-101
        if fn == 'phi':
            sec = self._pair_section('grimes', section)
            A, rh, C, D, ga, r0 = (self._p('grimes', sec, k, fixed)
                                   for k in ('A', 'rho', 'C', 'D', 'gamma', 'r0'))
            return lambda r: D * (torch.exp(-2.0 * ga * (r - r0)) -
                                  2.0 * torch.exp(-ga * (r - r0))) + \
                A * torch.exp(-r / rh) - C / r ** 6
        el = get_elements_from_kbody_term(section)[-1]
        if fn == 'rho':
            n = self._p('grimes', el, 'n', fixed)
            return lambda r: n / r ** 8 * (0.5 + 0.5 * torch.erf(20.0 * (r - 1.5)))
        G = self._p('grimes', el, 'G', fixed)
        return lambda rho: -G * torch.sqrt(rho)

    def _mishin(self, fn, section, fixed):
        # mishin.py:20-315 (embed, dipole, quadrupole; its rho / phi read undefined
        # variables upstream)
        if fn == 'embed':
            s1, s2, s3, s4, s5, s6, s7 = (self._p('mishinh', section, f's{k}', fixed)
                                          for k in range(1, 8))
            eps = 1e-14 if self.tdtype == torch.float64 else 1e-8

            def embed(rho):
                rho2 = rho * rho
                omega = 1.0 - (1.0 - s6 * rho2) / (1.0 + s7 * rho2 * rho2)
                return (s1 * rho + s2 * rho2 + s3 * rho * rho2 -
                        s4 * torch.pow(rho + eps, s5)) * omega
            return embed
        sec = self._pair_section('mishinh', section)
        pre = 'd' if fn == 'dipole' else 'q'
        p1, p2, p3, rc, h = (self._p('mishinh', sec, k, fixed)
                             for k in (pre + '1', pre + '2', pre + '3', 'rc', 'h'))

        def polar(r):
            # generic.py:52-84: (p1 exp(-p2 r) + p3) psi((r - rc) / h)
            x4 = torch.relu(-(r - rc) / h) ** 4
            return (p1 * torch.exp(-p2 * r) + p3) * (x4 / (1.0 + x4))
        return polar

    def _msah11(self, fn, section):
        # msah11.py:28-424: constants only -- nothing to train, but the functions may sit
        # next to trainable ones in a model
        pot = self.nn._empirical_functions['msah11']
        if fn == 'embed':
            if section == 'Al':
                def embed_al(rho):
                    m = rho >= 1e-12
                    x = torch.where(m, rho, torch.ones_like(rho))
                    y = -torch.sqrt(x) + 0.000093283590195398 * x ** 2 - \
                        0.0023491751192724 * x * torch.log(x)
                    return torch.where(m, y, torch.zeros_like(rho))
                return embed_al
            if section != 'Fe':
                raise KeyError(f"msah11: no embedding function for {section}")
            return lambda rho: -torch.sqrt(rho) - 0.00067314115586063 * rho ** 2 + \
                0.000000076514905604792 * rho ** 4
        if fn == 'rho':
            els = get_elements_from_kbody_term(section)
            key = pot._key(section) if len(els) == 2 else els[0] + els[0]
            order, factors, cutoffs = pot._RHO[key]
            return lambda r: sum(f * torch.clamp(rc - r, min=0.0) ** order
                                 for f, rc in zip(factors, cutoffs))
        d = pot._PHI[pot._key(section)]

        def phi(r):
            zero = torch.zeros_like(r)
            lo, hi, c = d['first']
            m = (r >= lo) & (r < hi)
            x = torch.where(m, r, torch.ones_like(r))
            y = (c[0] / x) * sum(c[1 + 2 * i] * torch.exp(c[2 + 2 * i] * x) for i in range(4))
            out = torch.where(m, y, zero)
            lo, hi, c = d['second']
            m = (r >= lo) & (r < hi)
            out = out + torch.where(m, torch.exp(c[0] + c[1] * r + c[2] * r ** 2 +
                                                 c[3] * r ** 3), zero)
            for lo, hi, coef, orders in d['polys']:
                m = (r >= lo) & (r < hi)
                x = torch.where(m, hi - r, zero)
                out = out + sum(a * x ** n for a, n in zip(coef, orders))
            return out
        return phi

    # -- dispatch -----------------------------------------------------------------------
    def get(self, fn, section):
        name = self.nn.potentials[section][fn]
        fixed = self._fixed(fn, section)
        if name == 'nn':
            return self._mlp(fn, section)
        if name == 'sutton90' and fn in ('rho', 'phi', 'embed'):
            return self._sutton(fn, section, fixed)
        if name == 'Be/1' and fn in ('rho', 'phi', 'embed'):
            return self._agrawal(fn, section, fixed)
        if name == 'grimes' and fn in ('rho', 'phi', 'embed'):
            return self._grimes(fn, section, fixed)
        if name == 'mishinh' and fn in ('embed', 'dipole', 'quadrupole'):
            return self._mishin(fn, section, fixed)
        if name == 'msah11' and fn in ('rho', 'phi', 'embed'):
            return self._msah11(fn, section)
        if name in self.ZJW:
            if fn == 'rho':
                return self._zjw_rho(name, get_elements_from_kbody_term(section)[-1], fixed)
            if fn == 'phi':
                if name == 'zjw04xcp':
                    return self._zjw_phi_xcp(name, section, fixed)
                return self._zjw_phi(name, section, fixed)
            if fn == 'embed':
                return self._zjw_embed(name, section, fixed)
        raise NotImplementedError(
            f"training of '{name}' {fn} functions is not implemented "
            "(trainable forms: 'nn', 'zjw04' / xc / uxc / xcp, 'sutton90', 'Be/1', "
            "'grimes', 'mishinh' embed / dipole / quadrupole; 'msah11' as constants)")


class EamTrainer:
    """Holds the structures of this rank in ONE batch handle (lists and pair vectors built
    once: the geometry does not change during training) and the torch leaves."""

    def __init__(self, nn, device='cuda', loss_weights=None, per_atom_energy=True,
                 freeze_reference_fixed=True):
        self.nn = nn
        self.device = device
        self.dt = get_float_dtype()
        self.tdtype = torch.float64 if self.dt.name == 'float64' else torch.float32
        self.elements = nn.elements
        self.kind = {_lib.EAM_ALLOY: 'alloy', _lib.EAM_FS: 'fs', _lib.EAM_ADP: 'adp'}[nn.kind]
        self.loss_weights = dict(energy=1.0, forces=1.0, stress=1.0)
        self.loss_weights.update(loss_weights or {})
        self.per_atom_energy = per_atom_energy
        self.fns = _Functions(nn, self.tdtype, device, freeze_reference_fixed)
        els = self.elements
        self._rho, self._phi, self._embed, self._dip, self._quad = {}, {}, {}, {}, {}
        for a, ea in enumerate(els):
            self._embed[a] = self.fns.get('embed', ea)
            for b, eb in enumerate(els):
                key = "".join(sorted([ea, eb])) if ea != eb else f"{ea}{ea}"
                # alloy: density contributed by the NEIGHBOUR element (alloy.py:162-176);
                # FS: function of the ordered pair centre-neighbour (fs.py:180-203)
                self._rho[(a, b)] = self.fns.get('rho', f"{ea}{eb}" if self.kind == 'fs' else eb)
                self._phi[(a, b)] = self.fns.get('phi', key)
                if self.kind == 'adp':
                    self._dip[(a, b)] = self.fns.get('dipole', key)
                    self._quad[(a, b)] = self.fns.get('quadrupole', key)
        self.params = self.fns.params()
        self.structures = []
        self._batch = None

    # -- data ------------------------------------------------------------------
    def add_structure(self, atoms, energy, forces, stress):
        t = lambda a: torch.tensor(np.asarray(a), dtype=self.tdtype, device=self.device)
        self.structures.append(dict(atoms=atoms, n=len(atoms),
                                    volume=float(atoms.get_volume()), energy=t(energy),
                                    forces=t(forces), stress=t(stress)))
        self._batch = None
        self._graph = None      # a captured step belongs to the old batch

    def _ensure_batch(self):
        if self._batch is not None:
            return self._batch
        clf = self.nn.transformer
        S = self.structures
        bf = clf.get_batch_features([s['atoms'] for s in S], nbr=_lib.NeighborList())
        i, j, D = bf.nbr.pairs()
        dev = self.device
        types = torch.as_tensor(bf.types, device=dev).long()
        i, j = i.long(), j.long()
        n_atoms = torch.tensor([s['n'] for s in S], device=dev)
        ti, tj = types[i], types[j]
        nel = len(self.elements)
        groups = {}
        for a in range(nel):
            for b in range(nel):
                sel = torch.nonzero((ti == a) & (tj == b)).reshape(-1)
                if sel.numel():
                    groups[(a, b)] = sel
        self._batch = dict(
            nbr=bf.nbr, i=i, D=D.to(self.tdtype), types=types, groups=groups,
            by_type={a: torch.nonzero(types == a).reshape(-1) for a in range(nel)},
            sid=torch.repeat_interleave(torch.arange(len(S), device=dev), n_atoms),
            n_atoms=n_atoms,
            volume=torch.tensor([s['volume'] for s in S], dtype=self.tdtype, device=dev),
            energy=torch.stack([s['energy'] for s in S]),
            forces=torch.cat([s['forces'] for s in S]),
            stress=torch.stack([s['stress'] for s in S]))
        return self._batch

    # -- model -------------------------------------------------------------------
    def atomic_energies(self, D):
        """E_i of every atom of the batch as a function of the pair vectors D [nij,3]
        (eam.py:265-298,495-570; adp.py:315-586 with the per-term squaring)."""
        B = self._batch
        n = B['types'].shape[0]
        z = lambda *shape: torch.zeros(*shape, dtype=self.tdtype, device=self.device)
        r = torch.sqrt(torch.sum(D * D, dim=1) + self.dt.eps)       # universal.py:470-473
        rho, pair, adp = z(n), z(n), z(n)
        for (a, b), sel in B['groups'].items():
            rr, ii = r[sel], B['i'][sel]
            rho = rho.index_add(0, ii, self._rho[(a, b)](rr))
            pair = pair.index_add(0, ii, self._phi[(a, b)](rr))
            if self.kind == 'adp':
                DD = D[sel]
                u, w = self._dip[(a, b)](rr), self._quad[(a, b)](rr)
                mu = z(n, 3).index_add(0, ii, u[:, None] * DD)
                lam = z(n, 3, 3).index_add(0, ii, w[:, None, None] * DD[:, :, None] *
                                           DD[:, None, :])
                diag = torch.diagonal(lam, dim1=1, dim2=2)
                off = lam[:, 0, 1] ** 2 + lam[:, 0, 2] ** 2 + lam[:, 1, 2] ** 2
                tr = diag.sum(dim=1)
                adp = adp + 0.5 * torch.sum(mu * mu, dim=1) + \
                    0.5 * (torch.sum(diag * diag, dim=1) + 2.0 * off) - tr * tr / 6.0
        embed = z(n)
        for a, sel in B['by_type'].items():
            if sel.numel():
                embed = embed.index_add(0, sel, self._embed[a](rho[sel]))
        return embed + 0.5 * pair + adp

    def total_loss(self, want_forces=True, want_stress=True):
        B = self._ensure_batch()
        nb = len(self.structures)
        D = B['D'].detach().requires_grad_(True)
        e_atom = self.atomic_energies(D)
        E = torch.zeros(nb, dtype=self.tdtype, device=self.device).index_add(
            0, B['sid'], e_atom)
        w = self.loss_weights
        loss = losses.energy_loss(B['energy'], E, B['n_atoms'], self.per_atom_energy,
                                  w['energy'])
        parts = {'energy': loss.detach()}
        if want_forces or want_stress:
            g = torch.autograd.grad(E.sum(), D, create_graph=True)[0]
            F, W = PairForce.apply(g, B['nbr'])
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

    # CUDA-graph capture of loss + backward and the replaying train_step: shared with the
    # AtomicNN trainer (same structure: fixed batch, static shapes)
    from tensoralloy_b200.nn.atomic.training import AtomicNNTrainer as _A
    enable_graph = _A.enable_graph
    train_step = _A.train_step
    del _A

    def named_parameters(self):
        """reference variable name -> leaf tensor (trainable and fixed)."""
        return dict(self.fns.named)

    def sync_to_model(self):
        """Write the trained values back into the model's variables (the device tables
        are rebuilt on the next evaluation)."""
        for name, t in self.fns.named.items():
            v = t.detach().cpu().numpy()
            if self.nn._nn.owns(name):
                self.nn.set_variable(name, v)
            else:
                self.nn.set_variable(name, float(v))
