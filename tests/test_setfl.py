"""setfl tables: reader/writer round trip on CPU, tabulated-spline potentials on
the GPU vs the oracle spline and vs the analytic potential they tabulate."""
import os

import numpy as np
import pytest

from tensoralloy_b200.io.lammps import (SetFL, Spline, read_adp_setfl,
                                        read_eam_alloy_setfl, write_setfl)

GOLD = os.path.join(os.path.dirname(__file__), 'golden')


def _setfl_from_golden():
    """Rebuild the reference-written zjw04 Ni table (test_files/lammps/
    zjw04_Ni.alloy.eam, copied into zjw04_Ni_setfl.npz) as a SetFL object."""
    g = np.load(os.path.join(GOLD, 'zjw04_Ni_setfl.npz'))
    nr, dr, nrho, drho = int(g['nr']), float(g['dr']), int(g['nrho']), float(g['drho'])
    r = np.linspace(0.0, nr * dr, nr, endpoint=False)
    phi = g['rphi_NiNi'].copy()
    phi[1:] /= r[1:]
    sp = lambda n, d, y: Spline(0.0, 0.0, np.linspace(0, n * d, n, endpoint=False), y, True)
    return SetFL(elements=['Ni'], rho={'Ni': sp(nr, dr, g['rho_Ni'])},
                 phi={'NiNi': sp(nr, dr, phi)}, embed={'Ni': sp(nrho, drho, g['F_Ni'])},
                 dipole={}, quadrupole={}, nr=nr, dr=dr, nrho=nrho, drho=drho,
                 rcut=float(g['rc']), atomic_masses=[58.6934], lattice_constants=[3.52],
                 lattice_types=['fcc'])


def test_setfl_round_trip(tmp_path):
    s = _setfl_from_golden()
    f = tmp_path / 'Ni.alloy.eam'
    write_setfl(str(f), s, comments=("a", "b", "c"))
    t = read_eam_alloy_setfl(str(f))
    assert t.elements == ['Ni'] and t.nr == s.nr and t.nrho == s.nrho
    assert abs(t.dr - s.dr) < 1e-18 and abs(t.rcut - s.rcut) < 1e-12
    assert np.abs(t.rho['Ni'].y - s.rho['Ni'].y).max() < 1e-15
    assert np.abs(t.embed['Ni'].y - s.embed['Ni'].y).max() < 1e-15
    rel = np.abs(t.phi['NiNi'].y[1:] - s.phi['NiNi'].y[1:]) / (1 + np.abs(s.phi['NiNi'].y[1:]))
    assert rel.max() < 1e-13
    # ADP layout: dipole / quadrupole blocks stored raw
    s.dipole = {'NiNi': Spline(0, 0, s.rho['Ni'].x, np.sin(s.rho['Ni'].x), True)}
    s.quadrupole = {'NiNi': Spline(0, 0, s.rho['Ni'].x, np.cos(s.rho['Ni'].x), True)}
    f2 = tmp_path / 'Ni.adp'
    write_setfl(str(f2), s, is_adp=True)
    u = read_adp_setfl(str(f2))
    assert np.abs(u.dipole['NiNi'].y - s.dipole['NiNi'].y).max() < 1e-15
    assert np.abs(u.quadrupole['NiNi'].y - s.quadrupole['NiNi'].y).max() < 1e-15
    c = s.rho['Ni'].coefficients()
    assert c.shape == (s.nr - 1, 4) and np.abs(c[:, 0] - s.rho['Ni'].y[:-1]).max() == 0.0


@pytest.mark.gpu
def test_spline_potential_matches_oracle_and_analytic(tmp_path):
    from oracle import eam as oeam
    from oracle import potentials as opot
    from tensoralloy_b200.atoms import bulk_fcc
    from tensoralloy_b200.calculator import TensorAlloyCalculator
    from tensoralloy_b200.nn.eam import EamAlloyNN
    from tensoralloy_b200.precision import precision_scope
    from tensoralloy_b200.transformer import UniversalTransformer
    s = _setfl_from_golden()
    f = tmp_path / 'Ni.alloy.eam'
    write_setfl(str(f), s)
    atoms = bulk_fcc('Ni', 3.52, (3, 3, 3))
    atoms.positions += np.random.default_rng(1).normal(scale=0.05,
                                                       size=atoms.positions.shape)
    rc = 5.9            # the table ends at nr * dr = 6.0
    with precision_scope('high'):
        nn = EamAlloyNN(['Ni'], custom_potentials=f"spline@{f}",
                        export_properties=['energy', 'forces', 'stress', 'hessian'])
        nn.attach_transformer(UniversalTransformer(['Ni'], rcut=rc))
        calc = TensorAlloyCalculator(nn)
        calc.calculate(atoms, ['energy', 'forces', 'stress'])
        e, fo, st = calc.results['energy'], calc.get_forces(atoms), calc.get_stress(atoms)
        nn2 = EamAlloyNN(['Ni'], custom_potentials='zjw04',
                         export_properties=['energy', 'forces', 'stress'])
        nn2.attach_transformer(UniversalTransformer(['Ni'], rcut=rc))
        c2 = TensorAlloyCalculator(nn2)
        e2, f2 = c2.get_potential_energy(atoms), c2.get_forces(atoms)
        small = bulk_fcc('Ni', 3.52, (2, 2, 2))
        small.positions += np.random.default_rng(2).normal(scale=0.03,
                                                           size=small.positions.shape)
        H = calc.get_hessian(small)
    tables = {'rho': {'Ni': (0.0, s.dr, s.rho['Ni'].y)},
              'phi': {'NiNi': (0.0, s.dr, s.phi['NiNi'].y)},
              'embed': {'Ni': (0.0, s.drho, s.embed['Ni'].y)}}
    pot = opot.SplineTable(tables)
    ref = oeam.eam_evaluate(pot, 'alloy', ['Ni'], atoms.get_chemical_symbols(),
                            atoms.positions, atoms.cell, [1, 1, 1], rc)
    n = len(atoms)
    assert abs(e - ref['energy']) / n < 1e-10
    assert np.abs(fo - ref['forces']).max() < 1e-8
    assert np.abs(st - ref['stress']).max() < 1e-9
    # the table (dr = 5e-4) reproduces the analytic potential it was written from
    assert abs(e - e2) / n < 1e-7 and np.abs(fo - f2).max() < 1e-5
    refh = oeam.eam_evaluate(pot, 'alloy', ['Ni'], small.get_chemical_symbols(),
                             small.positions, small.cell, [1, 1, 1], rc, hessian=True)
    m = len(small)
    assert np.abs(H - refh['hessian'].reshape(3 * m, 3 * m)).max() < 1e-7
