"""EAM E / F / virial on the GPU vs the oracle (torch float64 autograd).
Tolerances from BASELINE.json north_star: 1e-10 eV/atom, 1e-8 eV/A (float64);
1e-5 relative (float32)."""
import numpy as np
import pytest
import torch

from oracle import eam as oeam
from oracle import potentials as opot
from tensoralloy_b200 import _lib
from tensoralloy_b200.atoms import bulk_fcc
from tensoralloy_b200.nn.eam.potentials import get_potential

pytestmark = pytest.mark.gpu


def gpu_eam(pot_name, elements, symbols, pos, cell, pbc, rc, precision=0,
            kind=_lib.EAM_ALLOY):
    elements = sorted(elements)
    pot = get_potential(pot_name)
    n_el = len(elements)
    rho = [pot.rho(f'{a}{b}') for a in elements for b in elements]
    phi = [pot.phi(''.join(sorted([a, b]))) for a in elements for b in elements]
    embed = [pot.embed(a) for a in elements]
    model = _lib.EamModel(kind, n_el, rho, phi, embed)
    nl = _lib.NeighborList()
    n = len(pos)
    d_pos = torch.tensor(pos, dtype=torch.float64, device='cuda').contiguous()
    d_t = torch.tensor([elements.index(s) for s in symbols], dtype=torch.int32,
                       device='cuda')
    nl.build(d_pos, d_t, cell, pbc, rc)
    e = torch.zeros(1, dtype=torch.float64, device='cuda')
    ea = torch.zeros(n, dtype=torch.float64, device='cuda')
    f = torch.zeros((n, 3), dtype=torch.float64, device='cuda')
    v = torch.zeros(9, dtype=torch.float64, device='cuda')
    model.eval(nl, precision, e, ea, f, v)
    torch.cuda.synchronize()
    return (e.cpu().numpy()[0], ea.cpu().numpy(), f.cpu().numpy(),
            v.cpu().numpy().reshape(3, 3))


def compare(pot_name, elements, symbols, pos, cell, pbc, rc):
    ref = oeam.eam_evaluate(opot.get_potential(pot_name), 'alloy', elements,
                            symbols, pos, cell, pbc, rc)
    e, ea, f, v = gpu_eam(pot_name, elements, symbols, pos, cell, pbc, rc)
    n = len(pos)
    assert abs(e - ref['energy']) / n < 1e-10
    assert np.abs(ea - ref['energy/atom']).max() < 1e-10
    assert np.abs(f - ref['forces']).max() < 1e-8
    assert np.abs(v - ref['virial']).max() / n < 1e-8
    # float32 'medium'
    e32, _, f32, v32 = gpu_eam(pot_name, elements, symbols, pos, cell, pbc, rc,
                               precision=1)
    assert abs(e32 - ref['energy']) <= 1e-5 * abs(ref['energy'])
    fscale = max(np.abs(ref['forces']).max(), 1e-3)
    assert np.abs(f32 - ref['forces']).max() <= 2e-5 * fscale + 1e-5
    return ref


def test_ni_fcc_256():
    atoms = bulk_fcc('Ni', 3.52, (4, 4, 4))
    sym = atoms.get_chemical_symbols()
    ref = compare('zjw04', ['Ni'], sym, atoms.positions, atoms.cell, [1, 1, 1], 6.5)
    # fcc E/atom = -4.44999667 eV at rc 6.5
    assert abs(ref['energy'] / 256 + 4.44999667) < 1e-7
    rng = np.random.default_rng(611)
    pos = atoms.positions + rng.normal(scale=0.05, size=atoms.positions.shape)
    compare('zjw04', ['Ni'], sym, pos, atoms.cell, [1, 1, 1], 6.5)
    compare('zjw04', ['Ni'], sym, pos, atoms.cell, [1, 1, 1], 6.0)
    compare('zjw04xc', ['Ni'], sym, pos, atoms.cell, [1, 1, 1], 6.5)


def test_mo_ni_alloy():
    atoms = bulk_fcc('Ni', 3.6, (3, 3, 3))
    rng = np.random.default_rng(7)
    sym = ['Mo' if x < 0.4 else 'Ni' for x in rng.random(len(atoms))]
    pos = atoms.positions + rng.normal(scale=0.08, size=atoms.positions.shape)
    compare('zjw04', ['Mo', 'Ni'], sym, pos, atoms.cell, [1, 1, 1], 6.0)
    compare('zjw04xcp', ['Mo', 'Ni'], sym, pos, atoms.cell, [1, 1, 1], 6.0)


def test_small_cell_images():
    atoms = bulk_fcc('Ni', 3.52, (2, 2, 2))
    rng = np.random.default_rng(1)
    pos = atoms.positions + rng.normal(scale=0.03, size=atoms.positions.shape)
    compare('zjw04', ['Ni'], atoms.get_chemical_symbols(), pos, atoms.cell,
            [1, 1, 1], 6.5)
