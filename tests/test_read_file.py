"""`io.read.read_file` / `io.units.get_conversion_units` against the reference's reader
tests (tensoralloy/io/tests/test_read.py:19-96, test_units.py:19-48) on the same files
(fixtures copied by tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

from tensoralloy_b200.io.read import (Dataset, cell_to_cellpar, cellpar_to_cell,
                                      read_file)
from tensoralloy_b200.io.units import get_conversion_units

GOLD = os.path.join(os.path.dirname(__file__), 'golden')
# ase.units (CODATA 2014)
KCAL, MOL, HARTREE, GPA = 2.611447418269555e+22, 6.022140857e+23, 27.211386024367243, \
    1.0 / 160.21766208


def test_read_xyz():
    # test_read.py:19-29
    db = read_file(os.path.join(GOLD, 'B28_2frames.xyz'), num_examples=2)
    atoms = db[1]
    assert len(db) == 2
    assert abs(atoms.positions[1, 1] - 10.65007390) < 1e-7
    assert abs(atoms.info['energy'] - (-78.51063520)) < 1e-7
    assert atoms.cell.sum() > 1e-8 and np.allclose(atoms.cell, np.eye(3) * 20.0)
    assert np.array_equal(atoms.info['forces'], np.zeros((28, 3)))     # local minima
    assert db.metadata['extxyz'] is False and db.metadata['max_occurs'] == {'B': 28}
    assert len(read_file(os.path.join(GOLD, 'B28_2frames.xyz'), num_examples=1)) == 1


def test_read_extxyz():
    # test_read.py:32-52
    db = read_file(os.path.join(GOLD, 'examples.extxyz'))
    atoms = db[1]
    assert len(db) == 2 and len(atoms) == 21
    assert abs(atoms.info['forces'][0, 2] - 2.49790655) < 1e-6
    assert abs(atoms.info['energy'] - (-17637.613286)) < 1e-6
    assert db.metadata['max_occurs'] == {'C': 10, 'H': 8, 'O': 4}
    assert db.metadata['periodic'] is False and db.metadata['stress'] is False
    assert db.max_occurs['C'] == 10 and not db.has_stress
    # fmax filter (read.py:119-121)
    fmax = max(np.abs(a.info['forces']).max() for a in db)
    assert len(read_file(os.path.join(GOLD, 'examples.extxyz'), fmax=fmax * 0.999)) == 1


def test_read_snap_stress_in_kbar():
    # test_read.py:55-72
    db = read_file(os.path.join(GOLD, 'snap_Ni_id11.extxyz'), units={"stress": "kbar"})
    atoms = db[0]
    assert abs(atoms.info['stress'][0] - (-0.01388831152640921)) < 1e-8
    assert atoms.info['stress'].shape == (6,) and db.has_stress
    assert atoms.info['source'] == "Ni.AIMD.0"
    assert np.allclose(atoms.info['weights'], [1.0, 1.0, 0.0])
    assert atoms.pbc.all() and db.has_periodic_structures
    assert abs(atoms.cell[1, 0] - (-1.239893)) < 1e-9          # rows = lattice vectors


def test_read_stepmax_xyz():
    # test_read.py:75-82
    db = read_file(os.path.join(GOLD, 'Pu8.stepmax.xyz'), num_examples=1, file_type='stepmax')
    atoms = db[0]
    cellpars = cell_to_cellpar(atoms.cell)
    assert abs(atoms.positions[0, 2] - 3.301309) < 1e-7
    assert abs(cellpars[2] - 6.7942123514756485) < 1e-6
    assert abs(cellpars[4] - 79.74117275500237) < 1e-9 and abs(cellpars[3] - 90.0) < 1e-9
    assert abs(atoms.info['energy'] - (-32.4 * 27.211386024367243)) < 1e-6
    assert atoms.pbc.all() and np.array_equal(atoms.info['forces'], np.zeros((8, 3)))
    # triclinic round trip of the cell-parameter helpers
    par = [3.0, 4.0, 5.0, 80.0, 95.0, 110.0]
    assert np.allclose(cell_to_cellpar(cellpar_to_cell(par)), par, atol=1e-12)
    assert np.allclose(cellpar_to_cell([2.0, 3.0, 4.0, 90, 90, 90]), np.diag([2.0, 3.0, 4.0]),
                       atol=0)


def test_read_electron_temperature_and_entropy():
    # test_read.py:85-96; the 3x3 stress of this file becomes the Voigt vector
    db = read_file(os.path.join(GOLD, 'Be_liquid_4000K_1frame.extxyz'), num_examples=3)
    atoms = db[0]
    assert abs(atoms.info['etemperature'] - 0.34469373) < 1e-6
    assert abs(atoms.info['eentropy'] - 14.939843166274024) < 1e-6
    assert abs(atoms.info['stress'][0] - (-0.38507292526334685)) < 1e-12
    assert abs(atoms.info['stress'][3] - 1.0410837221973275e-08) < 1e-16    # yz
    assert len(atoms) == 128 and 'free_energy' in atoms.info


def test_unit_conversion():
    # test_units.py:19-48
    to_eV, _, to_s = get_conversion_units({'energy': 'kcal/mol*Hartree/eV',
                                           'stress': '0.1*GPa'})
    assert abs(to_eV - KCAL / MOL * HARTREE) < 1e-12 and abs(to_s - 0.1 * GPA) < 1e-15
    assert abs(get_conversion_units({'stress': 'kbar'})[2] - 0.1 * GPA) < 1e-15
    assert abs(get_conversion_units({'stress': 'eV/Angstrom**3'})[2] - 1.0) < 1e-15
    assert get_conversion_units(None) == (1.0, 1.0, 1.0)
    db = read_file(os.path.join(GOLD, 'examples.extxyz'), units={'energy': 'kcal/mol'})
    assert abs(db[1].info['energy'] - (-17637.613286 * KCAL / MOL)) < 1e-6
    with pytest.raises(ValueError, match="unknown unit"):
        get_conversion_units({'energy': '__import__("os")'})


def test_dataset_feeds_a_trainer_and_errors():
    db = read_file(os.path.join(GOLD, 'Be_liquid_4000K_1frame.extxyz'))

    class Sink:
        def __init__(self):
            self.rows = []

        def add_structure(self, atoms, energy, forces, stress):
            self.rows.append((atoms, energy, forces, stress))

    sink = db.fill(Sink())
    assert len(sink.rows) == 1 and sink.rows[0][2].shape == (128, 3)
    assert sink.rows[0][3].shape == (6,) and isinstance(db, Dataset)
    with pytest.raises(ValueError, match="Unknown file type"):
        read_file('x.cif')
    with pytest.raises(NotImplementedError):
        read_file('x.db')
    with pytest.raises(ValueError, match="stepmax"):
        read_file(os.path.join(GOLD, 'B28_2frames.xyz'), file_type='stepmax')


def test_read_vasp_xml():
    # io/tests/test_vasp.py:21-28 (kB / eV of ase.units, CODATA 2014)
    from tensoralloy_b200 import atoms_utils
    from tensoralloy_b200.io.vasp import read_vasp_xml
    kB = 8.6173303e-05
    atoms = next(read_vasp_xml(os.path.join(GOLD, 'Be_hcp_4000K_vasprun.xml.gz'), index=0))
    assert abs(atoms_utils.get_electron_temperature(atoms) - 4000.0 * kB) < 1e-6
    assert abs(atoms_utils.get_electron_entropy(atoms) - 0.2210591) < 1e-6
    info = atoms.info
    assert info['forces'].shape == (len(atoms), 3) and info['stress'].shape == (6,)
    assert set(atoms.get_chemical_symbols()) == {'Be'} and atoms.pbc.all()
    # E(sigma -> 0) lies between the free energy F and the internal energy U = F + sigma S
    hot = next(read_vasp_xml(os.path.join(GOLD, 'Be_hcp_4000K_vasprun.xml.gz'), index=0,
                             finite_temperature=True))
    U, F = hot.info['energy'], hot.info['free_energy']
    assert abs(U - (F + 0.2210591 * 4000.0 * kB)) < 1e-5
    assert F < info['energy'] < U
    assert abs(info['energy'] - 0.5 * (U + F)) < 1e-3        # first-order smearing correction
    # positions: scaled coordinates times the cell of the same step
    assert np.all(np.linalg.solve(atoms.cell.T, atoms.positions.T).T < 1.0 + 1e-9)


def test_read_vasp_md_xml():
    # io/tests/test_vasp.py:31-39
    from tensoralloy_b200.atoms_utils import get_kinetic_energy
    from tensoralloy_b200.io.vasp import read_vasp_xml
    path = os.path.join(GOLD, 'Be_md_vasprun.xml.gz')
    trajectory = list(read_vasp_xml(path, index=slice(0, 10), finite_temperature=True))
    assert len(trajectory) == 10
    assert abs(get_kinetic_energy(trajectory[4]) - 48.64234933) < 1e-8
    assert len(list(read_vasp_xml(path, index=[0, 3]))) == 2
    last = next(read_vasp_xml(path))                           # index = -1
    assert np.array_equal(last.positions, trajectory[9].positions)


def test_dataset_from_vasp_trajectory():
    from tensoralloy_b200.io.vasp import read_vasp_xml
    db = Dataset.from_images(read_vasp_xml(os.path.join(GOLD, 'Be_md_vasprun.xml.gz'),
                                           index=slice(0, 3), finite_temperature=True))
    assert len(db) == 3 and db.has_stress and db.has_periodic_structures
    assert list(db.max_occurs) == ['Be'] and db.max_occurs['Be'] == len(db[0])
    e, f, s = db.labels(1)
    assert f.shape == (len(db[1]), 3) and s.shape == (6,) and np.isfinite(e)
    assert db[1].info['etemperature'] > 0.0
