"""PhononCalculator (analysis/phonon.py:283-592): force constants in
phonopy layout vs the reference's golden Ni_fc2.npy, acoustic sum rule, Gamma-point
modes, and the dynamical-matrix frequencies against a direct supercell diagonalisation."""
import os

import numpy as np
import pytest

from tensoralloy_b200.analysis import PhononCalculator
from tensoralloy_b200.analysis.phonon import VaspToCm, VaspToTHz, get_masses
from tensoralloy_b200.atoms import Atoms, bulk_fcc
from tensoralloy_b200.nn.eam import EamAlloyNN
from tensoralloy_b200.precision import precision_scope
from tensoralloy_b200.transformer import UniversalTransformer

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), 'golden')


def _calc():
    nn = EamAlloyNN(['Ni'], custom_potentials='zjw04',
                    export_properties=['energy', 'forces', 'hessian'])
    nn.attach_transformer(UniversalTransformer(['Ni'], rcut=6.5))
    return PhononCalculator(nn)


def test_force_constants_match_reference_golden():
    with precision_scope('high'):
        calc = _calc()
        cubic = bulk_fcc('Ni', 3.52, (1, 1, 1))
        fc, sup = calc.get_force_constants(cubic, (2, 2, 2))
    fc2 = np.load(os.path.join(GOLD, 'Ni_fc2.npy'))      # nn/constraint/tests/test_fc2.py
    assert fc.shape == fc2.shape == (32, 32, 3, 3)
    assert np.abs(fc - fc2).max() < 3e-5                  # float32 golden
    # acoustic sum rule and symmetry of the force constants
    assert np.abs(fc.sum(axis=1)).max() < 1e-9
    assert np.abs(fc - fc.transpose(1, 0, 3, 2)).max() < 1e-10


def test_gamma_modes_and_reference_wavenumber_convention():
    with precision_scope('high'):
        calc = _calc()
        atoms = bulk_fcc('Ni', 3.52, (2, 2, 2))
        wn, modes = calc.get_frequencies_and_normal_modes(atoms)
    assert wn.shape == (96,) and modes.shape == (96, 96)
    assert np.abs(wn[:3]).max() < 1e-6 and wn[3] > 1.0       # three acoustic zero modes
    # reference convention (phonon.py:318-319): eigenvalue * VaspToCm
    H = calc.get_hessian(atoms)
    m = get_masses(atoms)
    assert abs(m[0] - 58.6934) < 1e-10
    ev = np.linalg.eigvalsh(H / m[0])
    assert np.abs(wn - ev * VaspToCm).max() < 1e-6 * np.abs(wn).max()


def test_dynamical_matrix_frequencies():
    with precision_scope('high'):
        calc = _calc()
        a = 3.52
        prim = Atoms(['Ni'], [[0.0, 0.0, 0.0]],
                     [[0.0, a / 2, a / 2], [a / 2, 0.0, a / 2], [a / 2, a / 2, 0.0]], True)
        # q commensurate with a 4x4x4 supercell: the dynamical matrix is exact and must
        # reproduce the spectrum of the supercell Hessian
        q = np.array([[0, 0, 0], [0.5, 0.0, 0.5], [0.25, 0.0, 0.0], [0.5, 0.5, 0.5]], float)
        f = calc.get_phonon_frequencies(prim, q, supercell=(4, 4, 4))
        sup = prim * (4, 4, 4)
        H = calc.get_hessian(sup)
    assert f.shape == (4, 3)
    assert np.abs(f[0]).max() < 1e-4                       # Gamma
    ev = np.linalg.eigvalsh(H / 58.6934)
    all_f = np.sort(np.sign(ev) * np.sqrt(np.abs(ev)) * VaspToTHz)
    for row in f[1:]:
        for x in row:                                      # every band value is a supercell mode
            assert np.abs(all_f - x).min() < 1e-6, (x,)
    # X point of fcc Ni: two transverse + one longitudinal branch, in the measured range
    # (experiment: 6.3 / 8.6 THz)
    fx = np.sort(f[1])
    assert abs(fx[0] - fx[1]) < 1e-6 and 5.0 < fx[0] < 8.0 and 7.5 < fx[2] < 11.0
