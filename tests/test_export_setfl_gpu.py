"""`export_to_setfl` (alloy.py:198-381, fs.py, adp.py:588-794): the model's
functions tabulated by the GPU evaluators (`tab_eam_tabulate`) and written in LAMMPS setfl
layout.  Mirrors the reference's own golden tests:
  nn/eam/tests/test_eam_alloy_nn.py:138-165   zjw04 Al-Cu  vs Zhou_AlCu.alloy.eam  (1e-12)
  nn/eam/tests/test_eam_fs_nn.py:60-83        msah11 Al-Fe vs Mendelev_Al_Fe.fs.eam (1e-8)"""
import os

import numpy as np
import pytest

from tensoralloy_b200.io import lammps as io
from tensoralloy_b200.nn.eam import AdpNN, EamAlloyNN, EamFsNN
from tensoralloy_b200.transformer import UniversalTransformer

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), 'golden')


def test_alloy_export_matches_zhou_alcu_table(tmp_path):
    g = np.load(os.path.join(GOLD, 'Zhou_AlCu_setfl.npz'))
    nn = EamAlloyNN(['Al', 'Cu'], custom_potentials='zjw04')
    nn.attach_transformer(UniversalTransformer(['Al', 'Cu'], rcut=6.0))
    path = str(tmp_path / 'AlCu.alloy.eam')
    nn.export_to_setfl(path, nr=int(g['nr']), dr=float(g['dr']), nrho=int(g['nrho']),
                       drho=float(g['drho']),
                       lattice_constants={'Al': 4.05, 'Cu': 3.615},
                       lattice_types={'Al': 'fcc', 'Cu': 'fcc'})
    t = io.read_eam_alloy_setfl(path)
    assert t.elements == ['Al', 'Cu'] and t.nr == 2000 and abs(t.rcut - 6.0) < 1e-12
    r = np.arange(t.nr) * t.dr
    for el in ('Al', 'Cu'):
        assert np.abs(t.embed[el].y - g[f'F_{el}']).max() < 1e-11
        assert np.abs(t.rho[el].y - g[f'rho_{el}']).max() < 1e-11
    for key, gold in (('AlAl', 'rphi_AlAl'), ('AlCu', 'rphi_CuAl'), ('CuCu', 'rphi_CuCu')):
        assert np.abs(t.phi[key].y[1:] * r[1:] - g[gold][1:]).max() < 1e-10, key


def _tokens(path, skip):
    with open(path) as fp:
        lines = fp.read().split('\n')
    out = []
    for line in lines[skip:]:
        out.extend(line.split())
    return lines, out


def test_fs_export_matches_mendelev_table(tmp_path):
    g = np.load(os.path.join(GOLD, 'Mendelev_AlFe_fs.npz'))
    nn = EamFsNN(['Al', 'Fe'], custom_potentials='msah11')
    nn.attach_transformer(UniversalTransformer(['Al', 'Fe'], rcut=6.5))
    path = str(tmp_path / 'AlFe.fs.eam')
    nr = nrho = 10000
    nn.export_to_setfl(path, nr=nr, dr=0.00065, nrho=nrho, drho=0.03,
                       lattice_constants={'Al': 4.04527, 'Fe': 2.855312},
                       lattice_types={'Al': 'fcc', 'Fe': 'bcc'})
    lines, tok = _tokens(path, 5)
    assert lines[3].split() == ['2', 'Al', 'Fe']
    assert lines[4].split()[0] == '10000' and abs(float(lines[4].split()[4]) - 6.5) < 1e-12
    pos = 0
    for el, z in (('Al', '13'), ('Fe', '26')):
        assert tok[pos] == z and tok[pos + 3] == ('fcc' if el == 'Al' else 'bcc')
        pos += 4
        F = np.array(tok[pos:pos + nrho], dtype=np.float64)
        pos += nrho
        assert np.abs(F[::10] - g['F_' + el]).max() < 1e-10
        for other in ('Al', 'Fe'):
            rho = np.array(tok[pos:pos + nr], dtype=np.float64)
            pos += nr
            assert np.abs(rho[::10] - g[f'rho_{el}{other}']).max() < 1e-10
    for key in ('AlAl', 'FeAl', 'FeFe'):
        y = np.array(tok[pos:pos + nr], dtype=np.float64)
        pos += nr
        assert np.abs(y[::10] - g['rphi_' + key]).max() < 1e-8, key      # the reference's delta
    assert pos == len(tok)


def test_adp_export_round_trips_through_the_reader(tmp_path):
    cp = {'Ni': {'rho': 'zjw04', 'embed': 'zjw04'},
          'NiNi': {'phi': 'zjw04', 'dipole': 'mishinh', 'quadrupole': 'mishinh'}}
    nn = AdpNN(['Ni'], custom_potentials=cp)
    nn.attach_transformer(UniversalTransformer(['Ni'], rcut=6.0))
    path = str(tmp_path / 'Ni.adp')
    nn.export_to_setfl(path, nr=600, dr=0.01, nrho=500, drho=0.1,
                       lattice_constants={'Ni': 3.52})
    t = io.read_adp_setfl(path)
    assert t.elements == ['Ni'] and set(t.dipole) == {'NiNi'} == set(t.quadrupole)
    model = nn._device_model()
    r = np.arange(600) * 0.01
    for which, table in (('rho', t.rho['Ni']), ('dipole', t.dipole['NiNi']),
                         ('quadrupole', t.quadrupole['NiNi'])):
        y = model.tabulate(which, 0, r)[0]
        assert np.abs(table.y - y).max() <= 1e-15 * max(1.0, np.abs(y).max())
    assert np.abs(t.dipole['NiNi'].y).max() > 1e-6
