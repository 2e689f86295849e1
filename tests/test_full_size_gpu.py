"""The headline configuration at FULL size (BASELINE.json configs[1]: EAM Ni fcc 63^3 cells =
1 000 188 atoms, zjw04, rc 6.5 A, rattle 0.05 A), checked through size-independent properties
and a sampled comparison with the oracle (the oracle itself needs ~0.1 s per atom
environment, so it cannot run the whole structure):

  * list:     sum of the per-atom counts = nij, nij even (every pair is listed twice), counts
              within the bounds the fcc shells allow, the same number of pairs for a permuted
              atom order;
  * physics:  sum of forces = 0, virial symmetric, energy / forces / virial invariant under a
              rigid translation with re-wrapping and under a permutation of the atom order,
              lists with a 0.3 A skin == lists at exactly rc, float32 within 1e-5 relative;
  * oracle:   E_i and F_i of 32 sampled atoms from their 2 rc environments, 1e-10 eV /
              1e-8 eV/A (the same check bench.py prints for 256 atoms).

All calls go through the C ABI (ctypes, tensoralloy_b200/_lib.py)."""
import numpy as np
import pytest

import bench
from tensoralloy_b200 import _lib
from tensoralloy_b200.nn.eam.potentials import get_potential

pytestmark = pytest.mark.gpu

CELLS = 63


@pytest.fixture(scope='module')
def system():
    import torch
    pos, cell = bench.make_lattice(CELLS)
    pot = get_potential('zjw04')
    model = _lib.EamModel(_lib.EAM_ALLOY, 1, [pot.rho('Ni')], [pot.phi('NiNi')],
                          [pot.embed('Ni')])
    return dict(torch=torch, pos=pos, cell=cell, model=model)


def _evaluate(system, pos, skin=0.0, precision=None, eatom=False):
    torch = system['torch']
    precision = _lib.PRECISION_HIGH if precision is None else precision
    n = len(pos)
    d_pos = torch.from_numpy(np.ascontiguousarray(pos)).cuda()
    nbr = _lib.NeighborList()
    if skin > 0:
        nbr.set_skin(skin)
    nbr.build(d_pos, None, system['cell'], [1, 1, 1], bench.RC)
    e = torch.zeros(1, dtype=torch.float64, device='cuda')
    f = torch.zeros((n, 3), dtype=torch.float64, device='cuda')
    v = torch.zeros(9, dtype=torch.float64, device='cuda')
    ea = torch.zeros(n, dtype=torch.float64, device='cuda') if eatom else None
    system['model'].eval(nbr, precision, energy=e, eatom=ea, forces=f, virial=v)
    torch.cuda.synchronize()
    return dict(nbr=nbr, E=float(e.item()), F=f, V=v.cpu().numpy().reshape(3, 3), ea=ea)


def test_list_and_conservation_laws(system):
    torch = system['torch']
    n = len(system['pos'])
    assert n == 1000188
    r = _evaluate(system, system['pos'], eatom=True)
    nij = r['nbr'].sizes()[0]
    counts = r['nbr'].counts()
    assert int(counts.sum().item()) == nij and nij % 2 == 0
    # fcc with a 0.05 A rattle: 86 neighbours up to the 5th shell, part of the 48 atoms of the
    # shell at 6.585 A
    assert 86 <= int(counts.min().item()) and int(counts.max().item()) <= 134
    # Newton's third law and the symmetric virial
    fsum = r['F'].sum(dim=0).abs().max().item()
    fmax = r['F'].abs().max().item()
    assert fsum < 1e-9 * n ** 0.5 * max(fmax, 1.0)
    V = r['V']
    assert np.abs(V - V.T).max() < 1e-9 * np.abs(V).max()
    # sum of the per-atom energies = the total (two reduction paths)
    assert abs(float(r['ea'].sum().item()) - r['E']) < 1e-12 * abs(r['E'])
    # E / atom of rattled fcc Ni sits just above the perfect lattice's -4.44999667 eV
    assert -4.4500 < r['E'] / n < -4.40
    system['ref'] = r


def test_sampled_atoms_against_the_oracle(system):
    r = system.get('ref') or _evaluate(system, system['pos'], eatom=True)
    rng = np.random.default_rng(20261018)
    idx = np.sort(rng.choice(len(system['pos']), size=32, replace=False))
    sel = system['torch'].as_tensor(idx, device='cuda')
    de, df = bench.sampled_check(system['pos'], system['cell'], idx,
                                 r['ea'][sel].cpu().numpy(), r['F'][sel].cpu().numpy(),
                                 per_call=32)
    assert de < 1e-10 and df < 1e-8, (de, df)


def test_translation_and_permutation_invariance(system):
    r0 = system.get('ref') or _evaluate(system, system['pos'])
    pos = system['pos']
    L = np.diag(system['cell'])
    # rigid translation, atoms re-wrapped into the cell: other cells, other tiles, other images
    shifted = np.mod(pos + np.array([1.2345, -7.7, 111.1]), L)
    r1 = _evaluate(system, shifted)
    assert abs(r1['E'] - r0['E']) < 1e-11 * abs(r0['E'])
    assert (r1['F'] - r0['F']).abs().max().item() < 1e-9
    assert np.abs(r1['V'] - r0['V']).max() < 1e-10 * np.abs(r0['V']).max()
    # permutation of the caller's atom order: forces come back in the caller's order
    perm = np.random.default_rng(3).permutation(len(pos))
    r2 = _evaluate(system, pos[perm])
    assert abs(r2['E'] - r0['E']) < 1e-11 * abs(r0['E'])
    d_perm = system['torch'].as_tensor(perm, device='cuda')
    assert (r2['F'] - r0['F'][d_perm]).abs().max().item() < 1e-9
    assert r2['nbr'].sizes()[0] == r0['nbr'].sizes()[0]


def test_skin_lists_and_float32(system):
    r0 = system.get('ref') or _evaluate(system, system['pos'])
    n = len(system['pos'])
    rs = _evaluate(system, system['pos'], skin=0.3)
    assert rs['nbr'].sizes()[0] > r0['nbr'].sizes()[0]          # the skin entries are listed
    assert abs(rs['E'] - r0['E']) < 1e-12 * abs(r0['E'])        # ... and contribute exactly 0
    assert (rs['F'] - r0['F']).abs().max().item() < 1e-10
    r32 = _evaluate(system, system['pos'], precision=_lib.PRECISION_MEDIUM)
    fscale = r0['F'].abs().max().item()
    assert abs(r32['E'] - r0['E']) < 1e-5 * abs(r0['E'])
    err = (r32['F'] - r0['F']).abs()
    # north_star's float32 bound (1e-5 relative) for the typical component, ten times that for
    # the worst of the 3e6 components (float32 records quantise positions to 2^-21 A)
    assert err.pow(2).mean().sqrt().item() < 1e-5 * fscale
    assert err.max().item() < 1e-4 * fscale
    assert n == len(r32['F'])
