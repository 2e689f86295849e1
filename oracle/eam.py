"""
oracle/eam.py -- TEST INFRASTRUCTURE (see oracle/__init__.py).

torch-CPU restatement of the reference EAM / Finnis-Sinclair / ADP energy models
and of the derived outputs of `BasicNN.build`:

  * rij          transformer/universal.py:448-474  D = Rj - Ri + S.h ,
                                                    r = sqrt(D.D + eps)
  * rho / embed  nn/eam/alloy.py:128-196 (alloy: rho of the NEIGHBOUR element),
                 nn/eam/fs.py:146-203    (FS: rho of the ordered pair),
                 nn/eam/eam.py:401-449
  * phi          nn/eam/eam.py:300-362   (0.5 applied after the row sum)
  * dipole/quad  nn/eam/adp.py:315-498   (squared PER K-BODY TERM, a reference
                                          quirk -- SURVEY.md 8(a)-a9)
  * energy       nn/eam/eam.py:265-298
  * forces       nn/basic.py:276-290     F = -dE/dR
  * virial       nn/basic.py:292-331     -F^T R + (dE/dh)^T h ; stress = virial/V
  * voigt        nn/basic.py:333-354     xx yy zz yz xz xy
  * hessian      nn/basic.py:410-421

The padded [T, Nvap, nnl_max] tensors and the virtual atom of the reference
carry zeros only, so the restatement sums over the directed pair list directly.
"""
import numpy as np
import torch

EPS = {torch.float64: 1e-14, torch.float32: 1e-8}   # precision.py:113-114


def pair_geometry(R, cell, i, j, S, periodic=True):
    """universal.py:448-474."""
    D = R[j] - R[i]
    if periodic:
        D = D + S.to(R.dtype) @ cell
    r = torch.sqrt(torch.sum(D * D, dim=-1) + EPS[R.dtype])
    return r, D


def eam_atomic_energies(pot, kind, elements, types, R, cell, i, j, S,
                        periodic=True, fns=None):
    """
    Per-atom energies E_i of an EAM-type model.

    pot      : oracle.potentials.Potential (or None when `fns` given)
    kind     : 'alloy' | 'fs' | 'adp'
    elements : sorted list of element symbols; types[k] indexes into it
    fns      : optional dict of callables overriding pot:
               rho(r, term)->t, phi(r, term)->t, embed(rho, el)->t,
               dipole(r, term), quadrupole(r, term)
    """
    n = R.shape[0]
    dtype = R.dtype
    r, D = pair_geometry(R, cell, i, j, S, periodic)
    ti = types[i]
    tj = types[j]
    fns = fns or {}
    rho_fn = fns.get('rho', getattr(pot, 'rho', None))
    phi_fn = fns.get('phi', getattr(pot, 'phi', None))
    embed_fn = fns.get('embed', getattr(pot, 'embed', None))
    rho = torch.zeros(n, dtype=dtype)
    pair = torch.zeros(n, dtype=dtype)
    adp = torch.zeros(n, dtype=dtype)
    nel = len(elements)
    for a in range(nel):
        for b in range(nel):
            sel = torch.nonzero((ti == a) & (tj == b)).reshape(-1)
            if sel.numel() == 0:
                continue
            rr = r[sel]
            ii = i[sel]
            term = elements[a] + elements[b]
            # ---- rho: alloy.py:162-196 / fs.py:180-203
            if kind == 'fs':
                y = rho_fn(rr, term)
            else:
                y = rho_fn(rr, elements[b])
            rho = rho.index_add(0, ii, y)
            # ---- phi: key is the sorted pair (alloy.py:57-66); 0.5 after sum
            key = ''.join(sorted([elements[a], elements[b]]))
            pair = pair.index_add(0, ii, phi_fn(rr, key))
            # ---- ADP: adp.py:350-498
            if kind == 'adp':
                dip_fn = fns.get('dipole', getattr(pot, 'dipole', None))
                quad_fn = fns.get('quadrupole', getattr(pot, 'quadrupole', None))
                u = dip_fn(rr, key)
                w = quad_fn(rr, key)
                DD = D[sel]
                mu = torch.zeros(n, 3, dtype=dtype).index_add(
                    0, ii, u[:, None] * DD)
                lam = torch.zeros(n, 3, 3, dtype=dtype).index_add(
                    0, ii, w[:, None, None] * DD[:, :, None] * DD[:, None, :])
                e_d = 0.5 * torch.sum(mu * mu, dim=1)
                diag = torch.diagonal(lam, dim1=1, dim2=2)
                off = lam[:, 0, 1] ** 2 + lam[:, 0, 2] ** 2 + lam[:, 1, 2] ** 2
                tr = diag.sum(dim=1)
                e_q = 0.5 * (torch.sum(diag * diag, dim=1) + 2.0 * off) \
                    - tr * tr / 6.0
                adp = adp + e_d + e_q
    embed = torch.zeros(n, dtype=dtype)
    for a in range(nel):
        sel = torch.nonzero(types == a).reshape(-1)
        if sel.numel():
            embed = embed.index_add(0, sel, embed_fn(rho[sel], elements[a]))
    return embed + 0.5 * pair + adp, rho


def evaluate(energy_fn, positions, cell, hessian=False):
    """
    Derived outputs of nn/basic.py:679-787 for a scalar-energy closure
    `energy_fn(R, h) -> (E, E_atom)`.

    Returns a dict of numpy arrays: energy, energy/atom, forces [N,3],
    virial [3,3], stress [6] (Voigt, eV/A^3), (hessian [N,3,N,3]).
    """
    dtype = positions.dtype
    R = positions.clone().requires_grad_(True)
    h = cell.clone().requires_grad_(True)
    E, e_atom = energy_fn(R, h)
    dEdR, dEdh = torch.autograd.grad(E, (R, h), create_graph=hessian,
                                     allow_unused=True)
    if dEdh is None:
        dEdh = torch.zeros_like(h)
    F = -dEdR
    # basic.py:306-317
    right = dEdh.t() @ h
    left = -(F.t() @ R)
    virial = left + right
    volume = torch.abs(torch.det(h.detach()))
    stress = virial / volume
    voigt = torch.stack([stress[0, 0], stress[1, 1], stress[2, 2],
                         stress[1, 2], stress[0, 2], stress[0, 1]])
    out = {
        'energy': E.detach().numpy(),
        'energy/atom': e_atom.detach().numpy(),
        'forces': F.detach().numpy(),
        'virial': virial.detach().numpy(),
        'stress': voigt.detach().numpy(),
    }
    if hessian:
        n = R.shape[0]
        H = torch.zeros(n * 3, n * 3, dtype=dtype)
        g = dEdR.reshape(-1)
        for k in range(n * 3):
            row = torch.autograd.grad(g[k], R, retain_graph=True)[0]
            H[k] = row.reshape(-1)
        out['hessian'] = H.reshape(n, 3, n, 3).numpy()
    return out


def eam_evaluate(pot, kind, elements, symbols, positions, cell, pbc, rc,
                 dtype=torch.float64, hessian=False, nl=None, fns=None):
    """Full oracle call: neighbour list -> E, E_atom, F, virial, stress."""
    from oracle import neighbor
    positions = np.asarray(positions, dtype=np.float64)
    cell = np.asarray(cell, dtype=np.float64).reshape(3, 3)
    if nl is None:
        nl = neighbor.neighbor_list(positions, cell, pbc, rc)
    i, j, S = nl[0], nl[1], nl[2]
    elements = sorted(elements)
    types = torch.tensor([elements.index(s) for s in symbols])
    ti = torch.from_numpy(i)
    tj = torch.from_numpy(j)
    tS = torch.from_numpy(S)
    periodic = bool(np.any(pbc))

    def energy_fn(R, h):
        e_atom, _ = eam_atomic_energies(pot, kind, elements, types, R, h,
                                        ti, tj, tS, periodic, fns=fns)
        return e_atom.sum(), e_atom

    return evaluate(energy_fn, torch.tensor(positions, dtype=dtype),
                    torch.tensor(cell, dtype=dtype), hessian=hessian)
