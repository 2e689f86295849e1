"""Pins the oracle (CPU restatement of the reference) to the reference's own
golden vectors; fixtures copied by tests/golden/make_golden.py."""
import os

import numpy as np
import torch

from oracle import eam as oeam
from oracle import neighbor as onl
from oracle import potentials as opot

GOLD = os.path.join(os.path.dirname(__file__), 'golden')


def _grid(n, d):
    # the reference tabulates with np.arange(0, n*d, d) (alloy.py:239-243)
    return torch.tensor(np.arange(0.0, n * d, d)[:n])


def test_zjw04_ni_tables_written_by_the_reference():
    g = np.load(os.path.join(GOLD, 'zjw04_Ni_setfl.npz'))
    pot = opot.get_potential('zjw04')
    r = _grid(int(g['nr']), float(g['dr']))
    rho = _grid(int(g['nrho']), float(g['drho']))
    assert np.abs(pot.rho(r, 'Ni').numpy() - g['rho_Ni']).max() < 1e-12
    assert np.abs((pot.phi(r, 'NiNi') * r).numpy() - g['rphi_NiNi'])[1:].max() < 1e-12
    assert np.abs(pot.embed(rho, 'Ni').numpy() - g['F_Ni']).max() < 1e-12


def test_zhou_alcu_tables():
    # nn/eam/tests/test_eam_alloy_nn.py:138-165 (assert_array_equal = 1e-12)
    g = np.load(os.path.join(GOLD, 'Zhou_AlCu_setfl.npz'))
    pot = opot.get_potential('zjw04')
    r = _grid(int(g['nr']), float(g['dr']))
    rho = _grid(int(g['nrho']), float(g['drho']))
    for el in ('Al', 'Cu'):
        assert np.abs(pot.rho(r, el).numpy() - g[f'rho_{el}']).max() < 1e-12
        assert np.abs(pot.embed(rho, el).numpy() - g[f'F_{el}']).max() < 1e-12
        assert np.abs((pot.phi(r, el + el) * r).numpy() - g[f'rphi_{el}{el}'])[1:].max() < 1e-12
    mixed = (pot.phi(r, 'AlCu') * r).numpy()
    assert np.abs(mixed - g['rphi_CuAl'])[1:].max() < 1e-12


def test_independent_moni_zhou04_tables():
    """MoNi_Zhou04.eam.alloy (NIST repository, written by Zhou's own Zhou04_create_v2.f --
    NOT by the reference; the reference compares against it through LAMMPS,
    nn/eam/tests/test_eam_alloy_nn.py:207-217): pins the Mo parameter set and the Mo-Ni
    mixing rule.  The Fortran generator clamps r below ~0.5 r_e and uses its own continuation
    of F above 1.15 rho_e, so the comparison covers r > 1.8 A and the first two branches of F."""
    g = np.load(os.path.join(GOLD, 'MoNi_Zhou04_setfl.npz'))
    pot = opot.get_potential('zjw04')
    nr, dr = int(g['nr']), float(g['dr'])
    r = torch.tensor(np.arange(nr) * dr)
    rho = torch.tensor(np.arange(int(g['nrho'])) * float(g['drho']))
    sel = (r > 1.8).numpy()
    rs = r[torch.as_tensor(sel)]
    for el in ('Mo', 'Ni'):
        assert np.abs(pot.rho(rs, el).numpy() - g[f'rho_{el}'][sel]).max() < 1e-12
        lim = (rho < 1.15 * pot.params[el]['rho_e']).numpy()
        F = pot.embed(rho, el).numpy()
        # the file carries F(0) = 1e-6 for Mo; everything else agrees to rounding
        assert np.abs(F - g[f'F_{el}'])[lim].max() < 2e-6
        assert np.abs(F - g[f'F_{el}'])[lim][1:].max() < 2e-6
    for key, term in (('rphi_MoMo', 'MoMo'), ('rphi_NiMo', 'MoNi'), ('rphi_NiNi', 'NiNi')):
        assert np.abs((pot.phi(rs, term) * rs).numpy() - g[key][sel]).max() < 1e-12
    assert np.abs(g['rphi_MoMo'][sel]).max() > 10.0        # not a comparison of zeros


def test_sutton_chen_ag_funcfl_table():
    """Ag.funcfl.eam, the golden of nn/eam/potentials/tests/test_sutton90.py:47-57 (LAMMPS
    funcfl: F(rho), Z(r) with phi = 27.2 * 0.529 Z^2 / r eV, rho(r)): pins AgSutton90."""
    g = np.load(os.path.join(GOLD, 'Ag_funcfl.npz'))
    pot = opot.get_potential('sutton90')
    r = torch.tensor(np.arange(len(g['Z'])) * float(g['dr']))
    rho = torch.tensor(np.arange(len(g['F'])) * float(g['drho']))
    s = slice(20, None)                       # r >= 0.2 A (the table diverges at r -> 0)
    assert np.abs(pot.rho(r[s], 'Ag').numpy() / g['rho'][s] - 1.0).max() < 1e-12
    assert np.abs(pot.embed(rho[1:], 'Ag').numpy() - g['F'][1:]).max() < 1e-12
    phi = 27.2 * 0.529 * g['Z'][s] ** 2 / r[s].numpy()
    assert np.abs(pot.phi(r[s], 'AgAg').numpy() / phi - 1.0).max() < 1e-12


def test_mendelev_al_fe_fs_table():
    """Mendelev_Al_Fe.fs.eam: the reference exports its AlFeMsah11 functions to setfl and
    compares them with this file at 1e-8 (nn/eam/tests/test_eam_fs_nn.py:60-83).  Pins the
    msah11 restatement: piecewise phi (screened Coulomb / exp bridge / polynomial tails),
    truncated-power densities, both embedding functions."""
    g = np.load(os.path.join(GOLD, 'Mendelev_AlFe_fs.npz'))
    pot = opot.get_potential('msah11')
    r = torch.tensor(np.arange(len(g['rho_AlAl'])) * float(g['dr']))
    rho = torch.tensor(np.arange(len(g['F_Al'])) * float(g['drho']))
    for el in ('Al', 'Fe'):
        assert np.abs(pot.embed(rho, el).numpy() - g['F_' + el]).max() < 1e-11
    for key in ('AlAl', 'AlFe', 'FeAl', 'FeFe'):
        assert np.abs(pot.rho(r, key).numpy() - g['rho_' + key]).max() < 1e-11
    for key, term in (('AlAl', 'AlAl'), ('FeAl', 'AlFe'), ('FeFe', 'FeFe')):
        assert np.abs((pot.phi(r, term) * r).numpy() - g['rphi_' + key])[1:].max() < 1e-10
    assert np.abs(g['rphi_FeFe']).max() > 100.0


def test_neighbor_oracle_known_counts():
    from tensoralloy_b200.atoms import bulk_fcc
    atoms = bulk_fcc('Ni', 3.52, (4, 4, 4))
    i, j, S, d, D = onl.neighbor_list(atoms.positions, atoms.cell, [1, 1, 1], 6.5)
    assert len(i) == 22016                       # 86 x 256
    assert np.all(np.bincount(i) == 86)
    i6 = onl.neighbor_list(atoms.positions, atoms.cell, [1, 1, 1], 6.0)[0]
    assert len(i6) == 19968                      # 78 x 256
    # KD-tree search == brute force, incl. multi-image small cells
    small = bulk_fcc('Ni', 3.52, (1, 1, 1))
    a = onl.neighbor_list(small.positions, small.cell, [1, 1, 1], 6.5)
    b = onl.neighbor_list_brute(small.positions, small.cell, [1, 1, 1], 6.5)
    for x, y in zip(a[:3], b[:3]):
        assert np.array_equal(x, y)
    assert np.abs(a[3] - b[3]).max() == 0.0


def test_eam_ni_known_answers():
    from tensoralloy_b200.atoms import bulk_fcc
    atoms = bulk_fcc('Ni', 3.52, (4, 4, 4))
    ref = oeam.eam_evaluate(opot.get_potential('zjw04'), 'alloy', ['Ni'],
                            atoms.get_chemical_symbols(), atoms.positions,
                            atoms.cell, [1, 1, 1], 6.5)
    # SURVEY.md 8(c) probe: fcc E/atom = -4.44999667 eV at rc 6.5
    assert abs(ref['energy'] / 256 + 4.44999667) < 1e-7
    assert np.abs(ref['forces']).max() < 1e-12          # perfect lattice
    # virial == sum_p dE/dD_p (x) D_p is symmetric and isotropic for fcc
    s = ref['stress']
    assert np.abs(s[:3] - s[0]).max() < 1e-14 and np.abs(s[3:]).max() < 1e-14


def test_forces_are_minus_gradient_by_finite_differences():
    from tensoralloy_b200.atoms import bulk_fcc
    atoms = bulk_fcc('Ni', 3.52, (2, 2, 2))
    rng = np.random.default_rng(0)
    pos = atoms.positions + rng.normal(scale=0.05, size=atoms.positions.shape)
    pot = opot.get_potential('zjw04')
    sym = atoms.get_chemical_symbols()
    nl = onl.neighbor_list(pos, atoms.cell, [1, 1, 1], 6.5)
    ref = oeam.eam_evaluate(pot, 'alloy', ['Ni'], sym, pos, atoms.cell, [1, 1, 1], 6.5,
                            nl=nl)
    h = 1e-5
    for (a, c) in ((0, 0), (5, 1), (17, 2)):
        p1, p2 = pos.copy(), pos.copy()
        p1[a, c] += h
        p2[a, c] -= h
        e1 = oeam.eam_evaluate(pot, 'alloy', ['Ni'], sym, p1, atoms.cell, [1, 1, 1],
                               6.5, nl=nl)['energy']
        e2 = oeam.eam_evaluate(pot, 'alloy', ['Ni'], sym, p2, atoms.cell, [1, 1, 1],
                               6.5, nl=nl)['energy']
        assert abs(-(e1 - e2) / (2 * h) - ref['forces'][a, c]) < 1e-6
    # stress by strain finite difference (checks the virial sign / units)
    eps = 1e-6
    cell = np.asarray(atoms.cell)
    vol = abs(np.linalg.det(cell))

    def energy_at(strain):
        F = np.eye(3) + strain
        return oeam.eam_evaluate(pot, 'alloy', ['Ni'], sym, pos @ F, cell @ F,
                                 [1, 1, 1], 6.5, nl=nl)['energy']
    st = np.zeros((3, 3))
    st[0, 0] = eps
    sxx = (energy_at(st) - energy_at(-st)) / (2 * eps) / vol
    assert abs(sxx - ref['stress'][0]) < 1e-7
