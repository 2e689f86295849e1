"""
oracle/atomic.py -- TEST INFRASTRUCTURE (see oracle/__init__.py).

torch-CPU restatement of the reference's symmetry-function AtomicNN:
  * cutoffs        nn/cutoff.py:20-85
  * G2             nn/atomic/sf.py:79-119
  * triples        transformer/universal.py:115-233 (symmetric: all j<k pairs of a
                   centre's neighbour row; r_jk from D_ik - D_ij, :213-218)
  * G4             nn/atomic/sf.py:121-182
  * feature order  per centre element c: G2 blocks of the radial terms
                   [cc, c-x1, ...] (utils.py:262-273), then G4 blocks of the angular
                   terms c+(sorted pair), pairs enumerated j<=k (utils.py:274-286);
                   inside a block tau runs over sklearn ParameterGrid order
                   (eta outer / omega inner; beta outer, gamma, zeta inner)
  * min-max        nn/atomic/atomic.py:157-195
  * MLP            nn/convolutional.py:257-290 (1x1 conv = per-atom dense layers,
                   ResNet add when consecutive widths match), nn/utils.py:39-74
  * energy         nn/atomic/atomic.py:270-302
Pinned to test_files/amp_Pd3O2.npz (nn/atomic/tests/test_sf.py:666-691).
"""
import math

import numpy as np
import torch

from oracle.eam import EPS, evaluate


def cosine_cutoff(r, rc):
    z = torch.clamp(r / rc, max=1.0)
    return 0.5 * (torch.cos(z * math.pi) + 1.0)


def polynomial_cutoff(r, rc, gamma=5.0):
    d = torch.clamp(r / rc, max=1.0)
    return 1.0 + gamma * d ** (gamma + 1.0) - (gamma + 1.0) * d ** gamma


def _cut(name):
    return cosine_cutoff if name == 'cosine' else polynomial_cutoff


def radial_grid(eta, omega):
    return [(e, o) for e in eta for o in omega]


def angular_grid(beta, gamma, zeta):
    return [(b, g, z) for b in beta for g in gamma for z in zeta]


def kbody_terms(elements, angular):
    """utils.py:237-290 (symmetric)."""
    elements = sorted(set(elements))
    per = {e: [e + e] for e in elements}
    for a in elements:
        for b in elements:
            if a != b:
                per[a].append(a + b)
    if angular:
        for c in elements:
            for j, a in enumerate(elements):
                for b in elements[j:]:
                    per[c].append(c + ''.join(sorted([a, b])))
    return per


def build_triples(i, n_atoms):
    """All (p, q) row-position pairs with p < q inside each centre's row
    (universal.py:176-232, symmetric).  `i` must be sorted by centre."""
    counts = np.bincount(i, minlength=n_atoms)
    starts = np.concatenate(([0], np.cumsum(counts)[:-1]))
    ps, qs = [], []
    for a in range(n_atoms):
        n = counts[a]
        if n < 2:
            continue
        p, q = np.triu_indices(n, k=1)
        ps.append(p + starts[a])
        qs.append(q + starts[a])
    if not ps:
        return np.zeros(0, np.int64), np.zeros(0, np.int64)
    return np.concatenate(ps), np.concatenate(qs)


def descriptors(elements, types, R, cell, i, j, S, rc, acut=None, angular=True,
                eta=(0.05, 4.0, 20.0, 80.0), omega=(0.0,), beta=(0.005,),
                gamma=(1.0, -1.0), zeta=(1.0, 4.0), cutoff='cosine',
                ang_list=None):
    """Returns G [n_atoms, D] (row = the centre's own feature layout).
    (i, j, S) = radial list (rc); ang_list = (i, j, S) for acut if different."""
    elements = sorted(elements)
    n = R.shape[0]
    nel = len(elements)
    dtype = R.dtype
    fcut = _cut(cutoff)
    rgrid = radial_grid(eta, omega)
    agrid = angular_grid(beta, gamma, zeta)
    n_r, n_a = len(rgrid), len(agrid)
    D = nel * n_r + (nel * (nel + 1) // 2 * n_a if angular else 0)
    ti = torch.as_tensor(i)
    tj = torch.as_tensor(j)
    tS = torch.as_tensor(S).to(dtype)
    Dij = R[tj] - R[ti] + tS @ cell
    rij = torch.sqrt((Dij * Dij).sum(-1) + EPS[dtype])
    G = torch.zeros(n, D, dtype=dtype)
    types_t = torch.as_tensor(types)
    ci, cj = types_t[ti], types_t[tj]
    # radial term index inside kbody_terms_for_element[centre]
    tidx = torch.where(ci == cj, torch.zeros_like(ci), cj - (cj > ci).long() + 1)
    fc = fcut(rij, rc)
    for tau, (e, o) in enumerate(rgrid):
        v = torch.exp(-e * (rij - o) ** 2 / rc ** 2) * fc
        col = tidx * n_r + tau
        G = G.index_put((ti, col), v, accumulate=True)
    if not angular:
        return G
    acut = acut if acut is not None else rc
    if ang_list is not None:
        ai, aj, aS = ang_list
        tai, taj = torch.as_tensor(ai), torch.as_tensor(aj)
        Da = R[taj] - R[tai] + torch.as_tensor(aS).to(dtype) @ cell
    else:
        ai, tai, taj, Da = i, ti, tj, Dij
    p, q = build_triples(np.asarray(ai), n)
    p, q = torch.as_tensor(p), torch.as_tensor(q)
    centre = tai[p]
    D1, D2 = Da[p], Da[q]
    D3 = D2 - D1
    r1 = torch.sqrt((D1 * D1).sum(-1) + EPS[dtype])
    r2 = torch.sqrt((D2 * D2).sum(-1) + EPS[dtype])
    r3 = torch.sqrt((D3 * D3).sum(-1) + EPS[dtype])
    sj, sk = types_t[taj[p]], types_t[taj[q]]
    lo, hi = torch.minimum(sj, sk), torch.maximum(sj, sk)
    # index of the sorted pair (lo <= hi) in the j<=k enumeration
    pair_idx = lo * nel - lo * (lo - 1) // 2 + (hi - lo)
    z = (r1 * r1 + r2 * r2 + r3 * r3) / acut ** 2
    upper = r1 * r1 + r2 * r2 - r3 * r3
    lower = 2.0 * r1 * r2
    theta = torch.where(lower == 0, torch.zeros_like(lower), upper / lower)
    fc3 = fcut(r1, acut) * (fcut(r2, acut) * fcut(r3, acut))
    base = nel * n_r
    for tau, (b, g, zt) in enumerate(agrid):
        outer = 2.0 ** (1.0 - zt)
        v = torch.pow(1.0 + g * theta, zt) * (torch.exp(-b * z) * fc3) * outer
        col = base + pair_idx * n_a + tau
        G = G.index_put((centre, col), v, accumulate=True)
    return G


def grap_radial(algorithm, prm, rij, rc):
    """f_tau(r) of the closed-form GRAP algorithms (grap.py:121-209; generic.py:15-30,
    87-100, 120-168), without cutoff."""
    if algorithm == 'sf':
        return torch.exp(-prm[0] * (rij - prm[1]) ** 2 / rc ** 2)
    if algorithm == 'morse':
        d, g, r0 = prm
        return d * (torch.exp(-2.0 * g * (rij - r0)) - 2.0 * torch.exp(-g * (rij - r0)))
    if algorithm == 'density':
        a, b, re = prm
        return a * torch.exp(-b * (rij / re - 1.0))
    if algorithm == 'pexp':
        rl, pl = prm
        return torch.exp(-(rij / rl) ** pl)
    raise ValueError(algorithm)


def grap_descriptors(elements, types, R, cell, i, j, S, rc, algorithm, grid, moments,
                     cutoff='cosine'):
    """Legacy-mode GenericRadialAtomicPotential descriptors (nn/atomic/grap.py:384-466,
    algorithms :121-234, generic.py:15-30,87-100,120-168).  grid: list of parameter
    tuples in the algorithm's key order (sf: eta, omega; morse: D, gamma, r0; density:
    A, beta, re; pexp: rl, pl).  Layout per radial term: [tau][moment]."""
    elements = sorted(elements)
    n, nel = R.shape[0], len(elements)
    dtype = R.dtype
    fcut = _cut(cutoff)
    moments = sorted(set(moments))
    n_r, n_m = len(grid), len(moments)
    ti, tj = torch.as_tensor(i), torch.as_tensor(j)
    Dij = R[tj] - R[ti] + torch.as_tensor(S).to(dtype) @ cell
    rij = torch.sqrt((Dij * Dij).sum(-1) + EPS[dtype])
    types_t = torch.as_tensor(types)
    ci, cj = types_t[ti], types_t[tj]
    tidx = torch.where(ci == cj, torch.zeros_like(ci), cj - (cj > ci).long() + 1)
    fc = fcut(rij, rc)
    u = Dij / rij[:, None]                       # div_no_nan: r > 0 for real pairs
    cols = []
    for tau, prm in enumerate(grid):
        v = grap_radial(algorithm, prm, rij, rc)
        w = v * fc
        for m in moments:
            # accumulate per (centre, term): index = centre * nel + term
            key = ti * nel + tidx
            if m == 0:
                acc = torch.zeros(n * nel, dtype=dtype).index_add(0, key, w)
                val = acc
            elif m == 1:
                acc = torch.zeros(n * nel, 3, dtype=dtype).index_add(0, key, w[:, None] * u)
                val = (acc * acc).sum(-1)
            else:
                uu = u[:, :, None] * u[:, None, :]
                acc = torch.zeros(n * nel, 3, 3, dtype=dtype).index_add(
                    0, key, w[:, None, None] * uu)
                val = (acc * acc).sum((-1, -2))
            cols.append((tau, moments.index(m), val.reshape(n, nel)))
    G = torch.zeros(n, nel * n_r * n_m, dtype=dtype)
    for tau, mi, val in cols:
        for term in range(nel):
            G[:, (term * n_r + tau) * n_m + mi] = val[:, term]
    return G


# -- GRAP new mode (grap.py:470-680) -------------------------------------------------
_AB = [(0, 0), (0, 1), (0, 2), (1, 1), (1, 2), (2, 2)]                    # grap.py:505-506
_ABC = [(0, 0, 0), (0, 0, 1), (0, 0, 2), (0, 1, 1), (0, 1, 2), (0, 2, 2),
        (1, 1, 1), (1, 1, 2), (1, 2, 2), (2, 2, 2)]                          # grap.py:510-512


def grap_multiplicity_tensor(max_moment, symmetric=False):
    """T_dm over the UNIQUE Cartesian index tuples (grap.py:470-496)."""
    d = {0: 1, 1: 4, 2: 10}.get(max_moment, 20)
    T = np.zeros((d, min(max_moment, 3) + 1))
    T[0, 0] = 1.0
    if max_moment >= 1:
        T[1:4, 1] = 1.0
    if max_moment >= 2:
        T[4:10, 2] = 1, 2, 2, 1, 2, 1
        if symmetric:
            T[0, 2] = -1.0 / 3.0
    if max_moment >= 3:
        T[10:20, 3] = 1, 3, 3, 3, 6, 3, 1, 3, 3, 1
        if symmetric:
            T[1:4, 3] = -3.0 / 5.0
    return T


def grap_moment_coeff(u, max_moment):
    """M_d per pair over the unique index tuples (grap.py:498-535): 1 | u_a | u_a u_b
    (a <= b) | u_a u_b u_c (a <= b <= c).  u: [P, 3] unit vectors."""
    rows = [torch.ones_like(u[:, 0])]
    if max_moment >= 1:
        rows += [u[:, a] for a in range(3)]
    if max_moment >= 2:
        rows += [u[:, a] * u[:, b] for a, b in _AB]
    if max_moment >= 3:
        rows += [u[:, a] * u[:, b] * u[:, c] for a, b, c in _ABC]
    return torch.stack(rows, 0)


def grap_T_dm_full(max_moment):
    """`get_T_dm` (grap.py:574-594): ones over the FULL 3^m index ranges."""
    dims = [1, 4, 13, 40, 121, 364]
    T = np.zeros((dims[max_moment], max_moment + 1))
    lo = 0
    for m in range(max_moment + 1):
        T[lo:lo + 3 ** m, m] = 1.0
        lo += 3 ** m
    return T


def grap_moment_tensor_full(u, max_moment):
    """`get_moment_tensor` (grap.py:537-572): all 3^m products, row-major."""
    rows = [torch.ones_like(u[:, :1]).T]
    cur = rows[0]
    for m in range(1, max_moment + 1):
        cur = (cur[:, None, :] * u.T[None, :, :]).reshape(-1, u.shape[0])
        rows.append(cur)
    return torch.cat(rows, 0)


def grap_descriptors_new_mode(elements, types, R, cell, i, j, S, rc, algorithm, grid,
                              max_moment, cutoff='cosine', symmetric=False):
    """`apply_model` of the new mode (grap.py:596-680) for the closed-form algorithms:
    P = sum_j H_k M_d, S = P^2, Q = S . T_dm, G[m=0] = sign(P_0) sqrt(Q_0 + 1e-16),
    G[m>0] = Q_m; layout per term [tau][m = 0..max_moment]."""
    elements = sorted(elements)
    n, nel = R.shape[0], len(elements)
    dtype = R.dtype
    fcut = _cut(cutoff)
    ti, tj = torch.as_tensor(i), torch.as_tensor(j)
    Dij = R[tj] - R[ti] + torch.as_tensor(S).to(dtype) @ cell
    rij = torch.sqrt((Dij * Dij).sum(-1) + EPS[dtype])
    types_t = torch.as_tensor(types)
    ci, cj = types_t[ti], types_t[tj]
    tidx = torch.where(ci == cj, torch.zeros_like(ci), cj - (cj > ci).long() + 1)
    fc = fcut(rij, rc)
    u = Dij / rij[:, None]
    if max_moment > 3:
        M, T = grap_moment_tensor_full(u, max_moment), grap_T_dm_full(max_moment)
    else:
        M, T = grap_moment_coeff(u, max_moment), grap_multiplicity_tensor(max_moment, symmetric)
    T = torch.as_tensor(T, dtype=dtype)
    if algorithm == 'nn':
        # NNAlgorithm (grap.py:211-234, 619-646): one filter network r -> K values shared by
        # every (centre, neighbour) element pair, no output bias; grid = dict(weights,
        # biases, activation, use_resnet_dt[, h_abck_modifier, rcov])
        W = [torch.as_tensor(w, dtype=dtype) for w in grid['weights']]
        b = [None if v is None else torch.as_tensor(v, dtype=dtype) for v in grid['biases']]
        # grap.py:621-632: h_abck_modifier 1 / 2 feed r / r_cov or exp(-r / r_cov) of the CENTRE
        # element; grid['rcov'] = {element: covalent radius} (ase.data.covalent_radii there)
        mod = int(grid.get('h_abck_modifier', 0))
        h_in = rij
        if mod:
            rcov = torch.as_tensor([grid['rcov'][e] for e in elements], dtype=dtype)[ci]
            h_in = rij / rcov if mod == 1 else torch.exp(-(rij / rcov))
        H = mlp(h_in[:, None], W, b, grid.get('activation', 'softplus'),
                grid.get('use_resnet_dt', True), None, all_outputs=True) * fc[:, None]
    else:
        H = torch.stack([grap_radial(algorithm, prm, rij, rc) * fc for prm in grid], 1)  # [P, K]
    key = ti * nel + tidx
    HM = H[:, :, None] * M.T[:, None, :]                                              # [P, K, D]
    P = torch.zeros(n * nel, H.shape[1], M.shape[0], dtype=dtype).index_add(0, key, HM)
    Q = (P * P) @ T                                                                   # [.., K, m]
    G0 = torch.sqrt(Q[..., :1] + 1e-16) * torch.sign(P[..., :1])
    G = torch.cat([G0, Q[..., 1:]], -1)
    return G.reshape(n, nel * H.shape[1] * T.shape[1])


def grap_from_dict(grap, elements, types, R, h, nl, rc):
    """grap: dict(algorithm, grid, moments, cutoff='cosine', new_mode=False,
    symmetric=False) -> descriptors of the legacy or the new formulation."""
    if grap.get('new_mode', False):
        return grap_descriptors_new_mode(elements, types, R, h, nl[0], nl[1], nl[2], rc,
                                         grap['algorithm'], grap['grid'],
                                         max(grap['moments']), grap.get('cutoff', 'cosine'),
                                         grap.get('symmetric', False))
    return grap_descriptors(elements, types, R, h, nl[0], nl[1], nl[2], rc,
                            grap['algorithm'], grap['grid'], grap['moments'],
                            grap.get('cutoff', 'cosine'))


def activation(name):
    name = name.lower()
    if name == 'softplus':
        return torch.nn.functional.softplus
    if name == 'tanh':
        return torch.tanh
    if name == 'relu':
        return torch.relu
    if name == 'leaky_relu':
        return lambda x: torch.nn.functional.leaky_relu(x, 0.2)   # tf default alpha
    if name == 'sigmoid':
        return torch.sigmoid
    if name == 'softsign':
        return torch.nn.functional.softsign
    if name == 'elu':
        return torch.nn.functional.elu
    if name == 'squareplus':
        return lambda x: 0.5 * (x + torch.sqrt(x * x + 4.0))
    raise ValueError(name)


def mlp(x, weights, biases, act, use_resnet_dt=False, out_bias=None, all_outputs=False):
    """convolutional.py:257-290.  weights[k]: [in, out]; last = output layer."""
    fn = activation(act)
    h = x
    nh = len(weights) - 1
    for k in range(nh):
        y = fn(h @ weights[k] + biases[k])
        if k and use_resnet_dt and weights[k].shape[1] == weights[k - 1].shape[1]:
            h = y + h
        else:
            h = y
    out = h @ weights[nh]
    if out_bias is not None:
        out = out + out_bias
    return out if all_outputs else out[:, 0]


def atomic_evaluate(elements, symbols, positions, cell, pbc, rc, params, sf=None,
                    acut=None, angular=True, dtype=torch.float64, hessian=False,
                    minmax=None, grap=None):
    """Full oracle call for AtomicNN + SymmetryFunction.
    params[el] = dict(weights=[...], biases=[...], out_bias=float or None,
                      activation=str, use_resnet_dt=bool)
    minmax[el] = (xlo, xhi) arrays or None."""
    from oracle import neighbor
    elements = sorted(elements)
    sf = dict(sf or {})
    positions = np.asarray(positions, dtype=np.float64)
    cell = np.asarray(cell, dtype=np.float64).reshape(3, 3)
    nl = neighbor.neighbor_list(positions, cell, pbc, rc)
    ang = None
    acut_eff = acut if acut is not None else rc
    if angular and abs(acut_eff - rc) > 5e-3:
        ang = neighbor.neighbor_list(positions, cell, pbc, acut_eff)[:3]
    types = np.array([elements.index(s) for s in symbols])

    def energy_fn(R, h):
        if grap is not None:
            G = grap_from_dict(grap, elements, types, R, h, nl, rc)
        else:
            G = descriptors(elements, types, R, h, nl[0], nl[1], nl[2], rc, acut_eff,
                            angular, ang_list=ang, **sf)
        e_atom = torch.zeros(R.shape[0], dtype=R.dtype)
        for a, el in enumerate(elements):
            sel = torch.nonzero(torch.as_tensor(types == a)).reshape(-1)
            if not sel.numel():
                continue
            x = G[sel]
            if minmax and minmax.get(el) is not None:
                xlo, xhi = [torch.as_tensor(v, dtype=R.dtype) for v in minmax[el]]
                den = xhi - xlo
                x = torch.where(den == 0, torch.zeros_like(x), (xhi - x) / den)
            p = params[el]
            W = [torch.as_tensor(w, dtype=R.dtype) for w in p['weights']]
            b = [None if v is None else torch.as_tensor(v, dtype=R.dtype)
                 for v in p['biases']]
            ob = p.get('out_bias')
            ob = None if ob is None else torch.as_tensor(ob, dtype=R.dtype)
            y = mlp(x, W, b, p.get('activation', 'softplus'),
                    p.get('use_resnet_dt', False), ob)
            e_atom = e_atom.index_add(0, sel, y)
        return e_atom.sum(), e_atom

    return evaluate(energy_fn, torch.tensor(positions, dtype=dtype),
                    torch.tensor(cell, dtype=dtype), hessian=hessian)


def td_heads(x, T, p, dtype):
    """nn/atomic/finite_temperature.py:211-304 for the atoms of ONE element.
    x [n, D] descriptors (already min-max scaled), T scalar electron temperature,
    p = dict(H=..., S=..., U=...) of mlp parameter dicts + 'algo' + 'special'.
    Returns per-atom (U, S, F = U - T S)."""
    def net(q, inp, squeeze=True):
        W = [torch.as_tensor(w, dtype=dtype) for w in q['weights']]
        b = [None if v is None else torch.as_tensor(v, dtype=dtype) for v in q['biases']]
        ob = q.get('out_bias')
        ob = None if ob is None else torch.as_tensor(ob, dtype=dtype)
        fn = activation(q.get('activation', 'softplus'))
        h = inp
        nh = len(W) - 1
        for k in range(nh):
            y = fn(h @ W[k] + b[k])
            if k and q.get('use_resnet_dt', False) and W[k].shape[1] == W[k - 1].shape[1]:
                h = y + h
            else:
                h = y
        out = h @ W[nh]
        if ob is not None:
            out = out + ob
        return out[:, 0] if squeeze else out
    H = net(p['H'], x, squeeze=False)                       # :243-256 (linear output layer)
    t = torch.full((x.shape[0], 1), float(T), dtype=dtype)
    Ht = torch.cat([H, t], dim=1)                            # :92-118
    Tv = t[:, 0]
    if p.get('special') == 'Be':                             # special/beryllium.py:23-77
        t2 = Tv * Tv
        ft = torch.relu(1.0 - 1.45 * Tv) ** 2
        base = -0.5718444 * t2 * ft + 0.83744317 * Tv + (-0.2110962) * (1.0 - ft)
        S = base * torch.nn.functional.softplus(net(p['S'], Ht))
    else:
        S = net(p['S'], Ht)                                  # :120-166
        if p.get('algo', 'default') == 'Sommerfeld':
            S = S * Tv
    U = net(p['U'], Ht)                                      # :168-209
    return U, S, U - Tv * S                                  # :296-301


def td_atomic_evaluate(elements, symbols, positions, cell, pbc, rc, params, etemperature,
                       sf=None, acut=None, angular=True, dtype=torch.float64, minmax=None):
    """Oracle call for TemperatureDependentAtomicNN: forces / stress derive from the
    FREE energy (basic.py:190-202); also returns 'energy' (U), 'eentropy' (S),
    'free_energy'."""
    from oracle import neighbor
    elements = sorted(elements)
    sf = dict(sf or {})
    positions = np.asarray(positions, dtype=np.float64)
    cell = np.asarray(cell, dtype=np.float64).reshape(3, 3)
    nl = neighbor.neighbor_list(positions, cell, pbc, rc)
    acut_eff = acut if acut is not None else rc
    ang = None
    if angular and abs(acut_eff - rc) > 5e-3:
        ang = neighbor.neighbor_list(positions, cell, pbc, acut_eff)[:3]
    types = np.array([elements.index(s) for s in symbols])
    keep = {}

    def energy_fn(R, h):
        G = descriptors(elements, types, R, h, nl[0], nl[1], nl[2], rc, acut_eff,
                        angular, ang_list=ang, **sf)
        n = R.shape[0]
        U = torch.zeros(n, dtype=R.dtype)
        S = torch.zeros(n, dtype=R.dtype)
        F = torch.zeros(n, dtype=R.dtype)
        for a, el in enumerate(elements):
            sel = torch.nonzero(torch.as_tensor(types == a)).reshape(-1)
            if not sel.numel():
                continue
            x = G[sel]
            if minmax and minmax.get(el) is not None:
                xlo, xhi = [torch.as_tensor(v, dtype=R.dtype) for v in minmax[el]]
                den = xhi - xlo
                x = torch.where(den == 0, torch.zeros_like(x), (xhi - x) / den)
            u, s, f = td_heads(x, etemperature, params[el], R.dtype)
            U = U.index_add(0, sel, u)
            S = S.index_add(0, sel, s)
            F = F.index_add(0, sel, f)
        keep['U'], keep['S'] = U.detach(), S.detach()
        return F.sum(), F

    out = evaluate(energy_fn, torch.tensor(positions, dtype=dtype),
                   torch.tensor(cell, dtype=dtype))
    out['free_energy'] = out['energy']
    out['free_energy/atom'] = out['energy/atom']
    out['energy'] = keep['U'].sum().numpy()
    out['energy/atom'] = keep['U'].numpy()
    out['eentropy'] = keep['S'].sum().numpy()
    out['eentropy/atom'] = keep['S'].numpy()
    return out
