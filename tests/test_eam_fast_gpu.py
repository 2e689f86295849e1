"""The lane-split single-element EAM kernels (csrc/eam_fast.cuh) and MD-valid list reuse
(skin + cutoff mask + displacement check) against the oracle on FRESH lists.

Tolerances (BASELINE.json north_star): float64 1e-10 eV/atom, 1e-8 eV/A (forces), virial
1e-8 eV per atom; float32 1e-5 relative."""
import os

import numpy as np
import pytest
import torch

from oracle import eam as oeam
from oracle import potentials as opot
from tensoralloy_b200 import _lib
from tensoralloy_b200.atoms import bulk_fcc
from tensoralloy_b200.nn.eam.potentials import get_potential

pytestmark = pytest.mark.gpu
RC = 6.5


def _model(name='zjw04'):
    pot = get_potential(name)
    return _lib.EamModel(_lib.EAM_ALLOY, 1, [pot.rho('Ni')], [pot.phi('NiNi')],
                         [pot.embed('Ni')])


def _eval(model, nl, n, precision=0):
    e = torch.zeros(1, dtype=torch.float64, device='cuda')
    ea = torch.zeros(n, dtype=torch.float64, device='cuda')
    f = torch.zeros((n, 3), dtype=torch.float64, device='cuda')
    v = torch.zeros(9, dtype=torch.float64, device='cuda')
    model.eval(nl, precision, e, ea, f, v)
    torch.cuda.synchronize()
    return e.item(), ea.cpu().numpy(), f.cpu().numpy(), v.cpu().numpy().reshape(3, 3)


def _oracle(pos, cell, pbc=(1, 1, 1), name='zjw04'):
    n = len(pos)
    return oeam.eam_evaluate(opot.get_potential(name), 'alloy', ['Ni'], ['Ni'] * n, pos, cell,
                             list(pbc), RC)


def _check64(res, ref, n):
    e, ea, f, v = res
    assert abs(e - ref['energy']) / n < 1e-10
    assert np.abs(ea - ref['energy/atom']).max() < 1e-10
    assert np.abs(f - ref['forces']).max() < 1e-8
    assert np.abs(v - ref['virial']).max() / n < 1e-8


def _check32(res, ref, n):
    e, ea, f, v = res
    assert abs(e - ref['energy']) <= 1e-5 * abs(ref['energy'])
    fscale = np.abs(ref['forces']).max()
    assert np.abs(f - ref['forces']).max() <= 1e-5 * fscale
    # the virial is a sum of pair terms g (x) D of both signs that nearly cancels at the
    # equilibrium density (|W| ~ 25 eV against sum |g.D| ~ |E| ~ 1e3 eV for 256 Ni atoms): the
    # rounding error of a float32 evaluation scales with the terms, not with their sum
    vscale = max(np.abs(ref['virial']).max(), abs(ref['energy']))
    assert np.abs(v - ref['virial']).max() <= 1e-5 * vscale


@pytest.fixture
def env_guard():
    saved = {k: os.environ.get(k) for k in ('TAB_EAMZ_L',)}
    yield
    for k, v in saved.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v


@pytest.mark.parametrize('cells,pbc', [((4, 4, 4), (1, 1, 1)), ((2, 2, 2), (1, 1, 1)),
                                       ((3, 3, 3), (1, 1, 0)), ((5, 4, 3), (0, 0, 0))])
def test_lane_split_variants(env_guard, cells, pbc):
    """Every lane count equals the oracle (multi-image cells, mixed and no periodicity
    included)."""
    atoms = bulk_fcc('Ni', 3.52, cells)
    rng = np.random.default_rng(611)
    pos = atoms.positions + rng.normal(scale=0.05, size=atoms.positions.shape)
    cell = np.asarray(atoms.cell)
    n = len(pos)
    ref = _oracle(pos, cell, pbc)
    model = _model()
    d_pos = torch.tensor(pos, dtype=torch.float64, device='cuda')
    # non-periodic directions need a frame that holds the atoms
    frame = cell.copy()
    org = np.zeros(3)
    for k in range(3):
        if not pbc[k]:
            frame[k] = frame[k] * 1.2 + np.eye(3)[k]
            org[k] = -1.0
    for L in (0, 1, 2, 4, 8):       # 0 = the thread-per-atom kernels of round 1
        os.environ['TAB_EAMZ_L'] = str(L)
        nl = _lib.NeighborList()
        nl.build_dd(d_pos, None, n, frame, org, pbc, RC)
        _check64(_eval(model, nl, n, 0), ref, n)
        _check32(_eval(model, nl, n, 1), ref, n)


def test_zjw04xc_embedding_through_fast_kernels(env_guard):
    atoms = bulk_fcc('Ni', 3.52, (4, 4, 4))
    rng = np.random.default_rng(3)
    pos = atoms.positions + rng.normal(scale=0.08, size=atoms.positions.shape)
    n = len(pos)
    ref = _oracle(pos, atoms.cell, name='zjw04xc')
    model = _model('zjw04xc')
    nl = _lib.NeighborList()
    nl.build(torch.tensor(pos, dtype=torch.float64, device='cuda'), None, atoms.cell,
             [1, 1, 1], RC)
    _check64(_eval(model, nl, n, 0), ref, n)
    _check32(_eval(model, nl, n, 1), ref, n)


@pytest.mark.parametrize('cells', [(4, 4, 4), (2, 2, 2)])
def test_skin_random_walk_equals_fresh_lists(cells):
    """20 MD-like steps on lists with a 0.5 A skin: tab_nbr_update + evaluation equals the
    oracle on a FRESH list at every step (the reference rebuilds per call, universal.py:58);
    the displacement check triggers the rebuilds."""
    atoms = bulk_fcc('Ni', 3.52, cells)
    rng = np.random.default_rng(12)
    pos = atoms.positions + rng.normal(scale=0.05, size=atoms.positions.shape)
    cell = np.asarray(atoms.cell)
    n = len(pos)
    model = _model()
    nl = _lib.NeighborList()
    nl.set_skin(0.5)
    rebuilds = 0
    for step in range(20):
        d_pos = torch.tensor(pos, dtype=torch.float64, device='cuda')
        rebuilt = nl.step(d_pos, None, cell, [1, 1, 1], RC)
        rebuilds += int(rebuilt)
        ref = _oracle(pos, cell)
        _check64(_eval(model, nl, n, 0), ref, n)
        _check32(_eval(model, nl, n, 1), ref, n)
        disp, skin = nl.max_displacement()
        assert skin == 0.5 and 2.0 * disp <= skin
        pos = pos + rng.normal(scale=0.03, size=pos.shape)
    # the walk moves atoms by ~0.03 sqrt(3 k): a rebuild every few steps, not every step
    assert 2 <= rebuilds <= 12


def test_skin_generic_kernels_and_adp():
    """The r < rc mask of the multi-species EAM and the ADP kernels: lists with a skin give
    the result of exact lists (same positions)."""
    from oracle import eam as oe
    atoms = bulk_fcc('Ni', 3.6, (3, 3, 3))
    rng = np.random.default_rng(7)
    sym = ['Mo' if x < 0.4 else 'Ni' for x in rng.random(len(atoms))]
    pos = atoms.positions + rng.normal(scale=0.08, size=atoms.positions.shape)
    n = len(pos)
    els = ['Mo', 'Ni']
    pot = get_potential('zjw04')
    rho = [pot.rho(f'{a}{b}') for a in els for b in els]
    phi = [pot.phi(''.join(sorted([a, b]))) for a in els for b in els]
    model = _lib.EamModel(_lib.EAM_ALLOY, 2, rho, phi, [pot.embed(a) for a in els])
    d_pos = torch.tensor(pos, dtype=torch.float64, device='cuda')
    d_t = torch.tensor([els.index(s) for s in sym], dtype=torch.int32, device='cuda')
    ref = oe.eam_evaluate(opot.get_potential('zjw04'), 'alloy', els, sym, pos, atoms.cell,
                          [1, 1, 1], 6.0)
    for skin in (0.0, 0.4):
        nl = _lib.NeighborList()
        nl.set_skin(skin)
        nl.build(d_pos, d_t, atoms.cell, [1, 1, 1], 6.0)
        _check64(_eval(model, nl, n, 0), ref, n)


def test_skin_lists_refused_by_exporting_consumers():
    atoms = bulk_fcc('Ni', 3.52, (3, 3, 3))
    nl = _lib.NeighborList()
    nl.set_skin(0.3)
    nl.build(torch.tensor(atoms.positions, dtype=torch.float64, device='cuda'), None,
             atoms.cell, [1, 1, 1], RC)
    with pytest.raises(_lib.TabError, match='skin'):
        nl.export()


def test_tile_build_with_skin_32k():
    """Large-system (tile) builder with a skin: E / F / virial equal the exact-list result of
    the same positions, before and after a displacement below skin / 2."""
    atoms = bulk_fcc('Ni', 3.52, (20, 20, 20))
    rng = np.random.default_rng(5)
    pos = atoms.positions + rng.normal(scale=0.05, size=atoms.positions.shape)
    n = len(pos)
    model = _model()
    d_pos = torch.tensor(pos, dtype=torch.float64, device='cuda')
    exact = _lib.NeighborList()
    exact.build(d_pos, None, atoms.cell, [1, 1, 1], RC)
    skin = _lib.NeighborList()
    skin.set_skin(0.4)
    skin.build(d_pos, None, atoms.cell, [1, 1, 1], RC)
    assert skin.sizes()[0] > exact.sizes()[0]
    a = _eval(model, exact, n, 0)
    b = _eval(model, skin, n, 0)
    assert abs(a[0] - b[0]) / n < 1e-12 and np.abs(a[2] - b[2]).max() < 1e-10
    assert np.abs(a[3] - b[3]).max() / n < 1e-10
    pos2 = pos + rng.uniform(-0.1, 0.1, size=pos.shape)      # |dR| <= 0.174 < 0.2
    d_pos2 = torch.tensor(pos2, dtype=torch.float64, device='cuda')
    assert skin.step(d_pos2, None, atoms.cell, [1, 1, 1], RC) is False
    exact.build(d_pos2, None, atoms.cell, [1, 1, 1], RC)
    a = _eval(model, exact, n, 0)
    b = _eval(model, skin, n, 0)
    assert abs(a[0] - b[0]) / n < 1e-12 and np.abs(a[2] - b[2]).max() < 1e-10
    assert np.abs(a[3] - b[3]).max() / n < 1e-10
    a32 = _eval(model, exact, n, 1)
    b32 = _eval(model, skin, n, 1)
    assert np.abs(a32[2] - b32[2]).max() < 1e-5 * np.abs(a32[2]).max()
    pos3 = pos2 + 0.25                                      # rigid shift: every atom moved 0.43
    assert skin.step(torch.tensor(pos3, dtype=torch.float64, device='cuda'), None, atoms.cell,
                     [1, 1, 1], RC) is True


def test_pipelined_rebuild_decision_matches_fresh_lists():
    """`NeighborList.step(..., max_step=...)`: the rebuild decision runs one step behind the
    device (previous displacement reading + the bound of one step) instead of blocking it.
    Every evaluation still equals the one on freshly built lists, and the lists are rebuilt no
    later than the blocking rule would (here: one step earlier at most)."""
    import torch
    from tensoralloy_b200 import _lib
    from tensoralloy_b200.atoms import fcc_positions
    from tensoralloy_b200.nn.eam.potentials import get_potential
    pos, cell = fcc_positions(3.52, 14, 14, 14)
    rng = np.random.default_rng(5)
    pos = pos + rng.normal(scale=0.05, size=pos.shape)
    n = len(pos)
    pot = get_potential('zjw04')
    model = _lib.EamModel(_lib.EAM_ALLOY, 1, [pot.rho('Ni')], [pot.phi('NiNi')],
                          [pot.embed('Ni')])
    skin, rc = 0.3, 6.5
    vel = rng.normal(size=(n, 3))
    vstep = 0.5 * skin / 6.3
    vel *= vstep / np.linalg.norm(vel, axis=1).max()
    d_pos = torch.tensor(pos, device='cuda')
    d_vel = torch.tensor(vel, device='cuda')
    runs = {}
    for mode in ('blocking', 'pipelined'):
        nbr = _lib.NeighborList()
        nbr.set_skin(skin)
        p = d_pos.clone()
        nbr.build(p, None, cell, [1, 1, 1], rc)
        e = torch.zeros(1, dtype=torch.float64, device='cuda')
        f = torch.zeros((n, 3), dtype=torch.float64, device='cuda')
        v = torch.zeros(9, dtype=torch.float64, device='cuda')
        rebuilt, out = [], []
        for k in range(20):
            p.add_(d_vel)
            rebuilt.append(nbr.step(p, None, cell, [1, 1, 1], rc,
                                    max_step=vstep if mode == 'pipelined' else None))
            model.eval(nbr, _lib.PRECISION_HIGH, energy=e, forces=f, virial=v)
            out.append((e.item(), f.clone()))
        runs[mode] = (rebuilt, out, p)
    # fresh lists at every step: the reference's semantics
    fresh = _lib.NeighborList()
    p = d_pos.clone()
    e = torch.zeros(1, dtype=torch.float64, device='cuda')
    f = torch.zeros((n, 3), dtype=torch.float64, device='cuda')
    v = torch.zeros(9, dtype=torch.float64, device='cuda')
    for k in range(20):
        p.add_(d_vel)
        fresh.build(p, None, cell, [1, 1, 1], rc)
        model.eval(fresh, _lib.PRECISION_HIGH, energy=e, forces=f, virial=v)
        for mode in runs:
            ek, fk = runs[mode][1][k]
            assert abs(ek - e.item()) < 1e-10 * n, (mode, k)
            assert (fk - f).abs().max().item() < 1e-8, (mode, k)
    rb, rp = runs['blocking'][0], runs['pipelined'][0]
    assert sum(rb) >= 2 and sum(rp) >= sum(rb)
    first_b, first_p = rb.index(True), rp.index(True)
    assert first_b - 1 <= first_p <= first_b          # never later, at most one step earlier
