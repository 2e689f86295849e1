"""
oracle/training.py -- TEST INFRASTRUCTURE (see oracle/__init__.py).

CPU restatement of one training-step loss and its parameter gradients for
AtomicNN: energy (per-atom RMSE) + forces RMSE + stress RMSE
(nn/losses.py:69-95,204-456; total = sum, nn/basic.py:626), with the parameter
gradients from torch double-backward -- what the reference gets from TF
second-order autograd (nn/opt.py:132-157).
"""
import numpy as np
import torch

from oracle import atomic as oat
from oracle import neighbor


def loss_and_grads(elements, structures, params, rc, sf=None, acut=None, angular=True,
                   minmax=None, weights=(1.0, 1.0, 1.0), eps=1e-14, grap=None,
                   extra_leaves=None):
    """structures: list of dict(symbols, positions, cell, pbc, energy, forces, stress).
    params[el]: dict(weights=[np], biases=[np or None], activation, use_resnet_dt,
    out_bias).  Returns (loss, parts, {el: ([dW...], [db...])}); with `extra_leaves` (torch
    leaves the descriptor closes over, e.g. the GRAP filter network handed in through
    grap['grid']) their gradients are returned under the key '__extra__'."""
    elements = sorted(elements)
    dtype = torch.float64
    P = {}
    for el in elements:
        p = params[el]
        W = [torch.tensor(w, dtype=dtype, requires_grad=True) for w in p['weights']]
        b = [None if v is None else torch.tensor(v, dtype=dtype, requires_grad=True)
             for v in p['biases']]
        P[el] = (W, b, p)
    E_pred, E_lab, n_at, F_pred, F_lab, S_pred, S_lab = [], [], [], [], [], [], []
    sf = dict(sf or {})
    for s in structures:
        pos = np.asarray(s['positions'], dtype=np.float64)
        cell = np.asarray(s['cell'], dtype=np.float64)
        cut = max(rc, acut) if (angular and acut) else rc
        nl = neighbor.neighbor_list(pos, cell, s['pbc'], cut)
        types = np.array([elements.index(x) for x in s['symbols']])
        R = torch.tensor(pos, dtype=dtype, requires_grad=True)
        h = torch.tensor(cell, dtype=dtype, requires_grad=True)
        if grap is not None:      # GenericRadialAtomicPotential (nn/atomic/grap.py:384-466)
            G = oat.grap_from_dict(grap, elements, types, R, h, nl, rc)
        else:
            G = oat.descriptors(elements, types, R, h, nl[0], nl[1], nl[2], rc,
                                acut if acut else rc, angular, **sf)
        e = torch.zeros((), dtype=dtype)
        for a, el in enumerate(elements):
            sel = torch.nonzero(torch.as_tensor(types == a)).reshape(-1)
            if not sel.numel():
                continue
            x = G[sel]
            if minmax and minmax.get(el) is not None:
                xlo, xhi = [torch.as_tensor(v, dtype=dtype) for v in minmax[el]]
                den = xhi - xlo
                x = torch.where(den == 0, torch.zeros_like(x), (xhi - x) / den)
            W, b, p = P[el]
            ob = b[-1] if p.get('out_bias') is not None else None
            e = e + oat.mlp(x, W, b, p.get('activation', 'softplus'),
                            p.get('use_resnet_dt', False), ob).sum()
        dR, dh = torch.autograd.grad(e, (R, h), create_graph=True)
        F = -dR
        virial = -(F.t() @ R) + dh.t() @ h               # basic.py:306-317
        stress = virial / abs(np.linalg.det(cell))
        voigt = torch.stack([stress[0, 0], stress[1, 1], stress[2, 2],
                             stress[1, 2], stress[0, 2], stress[0, 1]])
        E_pred.append(e)
        E_lab.append(float(s['energy']))
        n_at.append(len(pos))
        F_pred.append(F)
        F_lab.append(torch.tensor(np.asarray(s['forces']), dtype=dtype))
        S_pred.append(voigt)
        S_lab.append(torch.tensor(np.asarray(s['stress']), dtype=dtype))
    n = torch.tensor(n_at, dtype=dtype)

    def rmse(x, y):
        return torch.sqrt(torch.mean((x - y) ** 2) + eps)

    le = rmse(torch.tensor(E_lab, dtype=dtype) / n, torch.stack(E_pred) / n)
    lf = rmse(torch.cat(F_lab), torch.cat(F_pred))
    ls = rmse(torch.stack(S_lab), torch.stack(S_pred))
    loss = weights[0] * le + weights[1] * lf + weights[2] * ls
    leaves, index = [], []
    for el in elements:
        W, b, _ = P[el]
        for k, w in enumerate(W):
            leaves.append(w)
            index.append((el, 'W', k))
        for k, v in enumerate(b):
            if v is not None:
                leaves.append(v)
                index.append((el, 'b', k))
    extra = list(extra_leaves or [])
    grads = torch.autograd.grad(loss, leaves + extra, allow_unused=True)
    out = {el: ({}, {}) for el in elements}
    if extra:
        out['__extra__'] = [None if g is None else g.numpy() for g in grads[len(leaves):]]
        grads = grads[:len(leaves)]
    for (el, kind, k), g in zip(index, grads):
        out[el][0 if kind == 'W' else 1][k] = None if g is None else g.numpy()
    return (float(loss.detach()), {'energy': float(le.detach()),
                                   'forces': float(lf.detach()),
                                   'stress': float(ls.detach())}, out)


def eam_loss_and_grads(kind, elements, structures, fns, leaves_fn, rc,
                       weights=(1.0, 1.0, 1.0), eps=1e-14):
    """Training-step loss of an EAM / ADP model and its parameter gradients by torch
    double-backward on the CPU (reference: TF second-order autograd, nn/opt.py:132-157).
    fns: dict of callables rho(r, key), phi(r, key), embed(rho, el)[, dipole, quadrupole]
    closing over leaf tensors; leaves_fn() -> {name: leaf} AFTER the forward pass (shared
    parameters are created lazily).  Returns (loss, parts, {name: grad})."""
    from oracle import eam as oeam
    elements = sorted(elements)
    dtype = torch.float64
    E_pred, E_lab, n_at, F_pred, F_lab, S_pred, S_lab = [], [], [], [], [], [], []
    for s in structures:
        pos = np.asarray(s['positions'], dtype=np.float64)
        cell = np.asarray(s['cell'], dtype=np.float64)
        nl = neighbor.neighbor_list(pos, cell, s['pbc'], rc)
        types = torch.tensor([elements.index(x) for x in s['symbols']])
        R = torch.tensor(pos, dtype=dtype, requires_grad=True)
        h = torch.tensor(cell, dtype=dtype, requires_grad=True)
        e_atom, _ = oeam.eam_atomic_energies(
            None, kind, elements, types, R, h, torch.from_numpy(nl[0]),
            torch.from_numpy(nl[1]), torch.from_numpy(nl[2]), True, fns=fns)
        e = e_atom.sum()
        dR, dh = torch.autograd.grad(e, (R, h), create_graph=True)
        F = -dR
        virial = -(F.t() @ R) + dh.t() @ h               # basic.py:306-317
        stress = virial / abs(np.linalg.det(cell))
        voigt = torch.stack([stress[0, 0], stress[1, 1], stress[2, 2],
                             stress[1, 2], stress[0, 2], stress[0, 1]])
        E_pred.append(e)
        E_lab.append(float(s['energy']))
        n_at.append(len(pos))
        F_pred.append(F)
        F_lab.append(torch.tensor(np.asarray(s['forces']), dtype=dtype))
        S_pred.append(voigt)
        S_lab.append(torch.tensor(np.asarray(s['stress']), dtype=dtype))
    n = torch.tensor(n_at, dtype=dtype)

    def rmse(x, y):
        return torch.sqrt(torch.mean((x - y) ** 2) + eps)

    le = rmse(torch.tensor(E_lab, dtype=dtype) / n, torch.stack(E_pred) / n)
    lf = rmse(torch.cat(F_lab), torch.cat(F_pred))
    ls = rmse(torch.stack(S_lab), torch.stack(S_pred))
    loss = weights[0] * le + weights[1] * lf + weights[2] * ls
    leaves = leaves_fn()
    names = list(leaves)
    grads = torch.autograd.grad(loss, [leaves[k] for k in names], allow_unused=True)
    return (float(loss.detach()),
            {'energy': float(le.detach()), 'forces': float(lf.detach()),
             'stress': float(ls.detach())},
            {k: (None if g is None else g.numpy()) for k, g in zip(names, grads)})


def td_loss_and_grads(elements, structures, params, rc, minimize, sf=None, acut=None,
                      angular=True, minmax=None, eps=1e-14):
    """Training-step loss of a temperature-dependent AtomicNN (finite_temperature.py:
    211-304,338-366): energy losses for the properties of `minimize` among 'energy' (U),
    'free_energy' (F), 'eentropy' (S), forces / stress from F.  params[el] = dict(H=, S=,
    U=, algo=, special=) of mlp dicts (weights, biases, out_bias as arrays); structures carry
    'etemperature', 'free_energy', 'eentropy'.  Returns (loss, parts, {name: grad}) with the
    reference variable names TD/<El>/<head>/..."""
    elements = sorted(elements)
    dtype = torch.float64
    leaves = {}
    P = {}
    for el in elements:
        P[el] = dict(algo=params[el].get('algo', 'default'), special=params[el].get('special'))
        for head in ('H', 'S', 'U'):
            q = params[el][head]
            base = f"TD/{el}/{head}"
            W, b = [], []
            for k, w in enumerate(q['weights'][:-1]):
                W.append(torch.tensor(w, dtype=dtype, requires_grad=True))
                b.append(torch.tensor(q['biases'][k], dtype=dtype, requires_grad=True))
                leaves[f"{base}/Conv1d{k + 1}/kernel"] = W[-1]
                leaves[f"{base}/Conv1d{k + 1}/bias"] = b[-1]
            W.append(torch.tensor(q['weights'][-1], dtype=dtype, requires_grad=True))
            leaves[f"{base}/Output/kernel"] = W[-1]
            ob = None
            if q.get('out_bias') is not None:
                ob = torch.tensor(q['out_bias'], dtype=dtype, requires_grad=True)
                leaves[f"{base}/Output/bias"] = ob
            P[el][head] = dict(weights=W, biases=b + [None], out_bias=ob,
                               activation=q.get('activation', 'softplus'),
                               use_resnet_dt=q.get('use_resnet_dt', False))
    pred = {'energy': [], 'free_energy': [], 'eentropy': []}
    lab = {'energy': [], 'free_energy': [], 'eentropy': []}
    n_at, F_pred, F_lab, S_pred, S_lab = [], [], [], [], []
    sf = dict(sf or {})
    for s in structures:
        pos = np.asarray(s['positions'], dtype=np.float64)
        cell = np.asarray(s['cell'], dtype=np.float64)
        cut = max(rc, acut) if (angular and acut) else rc
        nl = neighbor.neighbor_list(pos, cell, s['pbc'], cut)
        types = np.array([elements.index(x) for x in s['symbols']])
        R = torch.tensor(pos, dtype=dtype, requires_grad=True)
        h = torch.tensor(cell, dtype=dtype, requires_grad=True)
        G = oat.descriptors(elements, types, R, h, nl[0], nl[1], nl[2], rc,
                            acut if acut else rc, angular, **sf)
        tot = {k: torch.zeros((), dtype=dtype) for k in pred}
        for a, el in enumerate(elements):
            sel = torch.nonzero(torch.as_tensor(types == a)).reshape(-1)
            if not sel.numel():
                continue
            x = G[sel]
            if minmax and minmax.get(el) is not None:
                xlo, xhi = [torch.as_tensor(v, dtype=dtype) for v in minmax[el]]
                den = xhi - xlo
                x = torch.where(den == 0, torch.zeros_like(x), (xhi - x) / den)
            U, S, F = oat.td_heads(x, s['etemperature'], P[el], dtype)
            tot['energy'] = tot['energy'] + U.sum()
            tot['eentropy'] = tot['eentropy'] + S.sum()
            tot['free_energy'] = tot['free_energy'] + F.sum()
        dR, dh = torch.autograd.grad(tot['free_energy'], (R, h), create_graph=True)
        Fo = -dR
        virial = -(Fo.t() @ R) + dh.t() @ h
        stress = virial / abs(np.linalg.det(cell))
        S_pred.append(torch.stack([stress[0, 0], stress[1, 1], stress[2, 2],
                                   stress[1, 2], stress[0, 2], stress[0, 1]]))
        S_lab.append(torch.tensor(np.asarray(s['stress']), dtype=dtype))
        F_pred.append(Fo)
        F_lab.append(torch.tensor(np.asarray(s['forces']), dtype=dtype))
        n_at.append(len(pos))
        for k in pred:
            pred[k].append(tot[k])
            lab[k].append(float(s[k]))
    n = torch.tensor(n_at, dtype=dtype)

    def rmse(x, y):
        return torch.sqrt(torch.mean((x - y) ** 2) + eps)

    loss = torch.zeros((), dtype=dtype)
    parts = {}
    for k in ('energy', 'free_energy', 'eentropy'):
        if k in minimize:
            parts[k] = rmse(torch.tensor(lab[k], dtype=dtype) / n, torch.stack(pred[k]) / n)
            loss = loss + parts[k]
    parts['forces'] = rmse(torch.cat(F_lab), torch.cat(F_pred))
    parts['stress'] = rmse(torch.stack(S_lab), torch.stack(S_pred))
    loss = loss + parts['forces'] + parts['stress']
    names = list(leaves)
    grads = torch.autograd.grad(loss, [leaves[k] for k in names], allow_unused=True)
    return (float(loss.detach()), {k: float(v.detach()) for k, v in parts.items()},
            {k: (None if g is None else g.numpy()) for k, g in zip(names, grads)})
