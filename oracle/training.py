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
                   minmax=None, weights=(1.0, 1.0, 1.0), eps=1e-14, grap=None):
    """structures: list of dict(symbols, positions, cell, pbc, energy, forces, stress).
    params[el]: dict(weights=[np], biases=[np or None], activation, use_resnet_dt,
    out_bias).  Returns (loss, {el: ([dW...], [db...])})."""
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
            G = oat.grap_descriptors(elements, types, R, h, nl[0], nl[1], nl[2], rc,
                                     grap['algorithm'], grap['grid'], grap['moments'],
                                     grap.get('cutoff', 'cosine'))
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
    grads = torch.autograd.grad(loss, leaves, allow_unused=True)
    out = {el: ({}, {}) for el in elements}
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
