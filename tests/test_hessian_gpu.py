"""Analytic EAM Hessian (csrc/hessian.cu) vs the oracle's autograd Hessian and the
reference's golden test_files/crystals/Ni_fc2.npy (nn/constraint/tests/
test_fc2.py:30-54, generated in float32 -> tolerance 1e-5)."""
import os

import numpy as np
import pytest

from oracle import eam as oeam
from oracle import potentials as opot
from tensoralloy_b200.atoms import Atoms, bulk_fcc, bulk_hcp
from tensoralloy_b200.calculator import TensorAlloyCalculator
from tensoralloy_b200.nn.eam import EamAlloyNN
from tensoralloy_b200.precision import precision_scope
from tensoralloy_b200.transformer import UniversalTransformer

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), 'golden')


def gpu_hessian(atoms, elements, pot, rc):
    with precision_scope('high'):
        nn = EamAlloyNN(elements, custom_potentials=pot,
                        export_properties=['energy', 'forces', 'hessian'])
        nn.attach_transformer(UniversalTransformer(elements, rcut=rc))
        calc = TensorAlloyCalculator(nn)
        calc.calculate(atoms, properties=['energy', 'forces', 'hessian'])
        raw = calc.results['hessian']
        assert raw.shape == (len(atoms) + 1, 3, len(atoms) + 1, 3) or len(elements) > 1
        return calc.get_hessian(atoms), calc


def test_ni_fc2_golden_and_oracle():
    atoms = bulk_fcc('Ni', 3.52, (2, 2, 2))        # bulk("Ni", cubic=True) * [2,2,2]
    H, calc = gpu_hessian(atoms, ['Ni'], 'zjw04', 6.5)
    n = len(atoms)
    fc2 = np.load(os.path.join(GOLD, 'Ni_fc2.npy'))          # [N, N, 3, 3]
    H4 = H.reshape(n, 3, n, 3).transpose(0, 2, 1, 3)
    # the golden was produced by the reference in float32 ('medium', eps 1e-8):
    # its entries (up to ~10 eV/A^2) carry ~1e-6 relative float32 noise
    assert np.abs(H4 - fc2).max() < 3e-5
    assert np.abs(H4 - fc2).max() / np.abs(fc2).max() < 3e-6
    vap = calc.transformer.get_vap_transformer(atoms)
    assert np.abs(vap.reverse_map_hessian(calc.results['hessian'],
                                          phonopy_format=True) - fc2).max() < 3e-5
    ref = oeam.eam_evaluate(opot.get_potential('zjw04'), 'alloy', ['Ni'],
                            atoms.get_chemical_symbols(), atoms.positions, atoms.cell,
                            [1, 1, 1], 6.5, hessian=True)
    assert np.abs(H - ref['hessian'].reshape(3 * n, 3 * n)).max() < 1e-8


@pytest.mark.parametrize("pot", ['Be/1', 'zjw04xc'])
def test_be_hcp_rattled(pot):
    atoms = bulk_hcp('Be', 2.2644, 3.5673, (3, 3, 2))
    rng = np.random.default_rng(4)
    atoms.positions += rng.normal(scale=0.01, size=atoms.positions.shape)
    H, _ = gpu_hessian(atoms, ['Be'], pot, 5.0)
    n = len(atoms)
    ref = oeam.eam_evaluate(opot.get_potential(pot), 'alloy', ['Be'],
                            atoms.get_chemical_symbols(), atoms.positions, atoms.cell,
                            [1, 1, 1], 5.0, hessian=True)
    scale = np.abs(ref['hessian']).max()
    assert np.abs(H - ref['hessian'].reshape(3 * n, 3 * n)).max() < 1e-8 * max(scale, 1)


def test_config5_be_5x5x3_properties():
    """Full-size config 5 (150 atoms): symmetry, acoustic sum rule and agreement
    with central finite differences of the GPU forces."""
    atoms = bulk_hcp('Be', 2.2644, 3.5673, (5, 5, 3))
    assert len(atoms) == 150
    rng = np.random.default_rng(8)
    atoms.positions += rng.normal(scale=0.01, size=atoms.positions.shape)
    H, calc = gpu_hessian(atoms, ['Be'], 'Be/1', 5.0)
    assert np.abs(H - H.T).max() < 1e-10
    assert np.abs(H.reshape(150, 3, 150, 3).sum(axis=2)).max() < 1e-9
    h = 1e-4
    for (a, c) in ((0, 0), (77, 2)):
        p = atoms.copy()
        p.positions[a, c] += h
        m = atoms.copy()
        m.positions[a, c] -= h
        with precision_scope('high'):
            fp = calc.get_forces(p)
            fm = calc.get_forces(m)
        col = -(fp - fm).reshape(-1) / (2 * h)
        assert np.abs(col - H[:, 3 * a + c]).max() < 1e-5


def test_binary_alloy_hessian():
    base = bulk_fcc('Ni', 3.6, (2, 2, 2))
    rng = np.random.default_rng(2)
    sym = ['Mo' if x < 0.4 else 'Ni' for x in rng.random(len(base))]
    atoms = Atoms(sym, base.positions + rng.normal(scale=0.05, size=base.positions.shape),
                  base.cell, True)
    H, _ = gpu_hessian(atoms, ['Mo', 'Ni'], 'zjw04', 5.5)
    n = len(atoms)
    ref = oeam.eam_evaluate(opot.get_potential('zjw04'), 'alloy', ['Mo', 'Ni'], sym,
                            atoms.positions, atoms.cell, [1, 1, 1], 5.5, hessian=True)
    assert np.abs(H - ref['hessian'].reshape(3 * n, 3 * n)).max() < 1e-8


def test_adp_hessian_from_analytic_forces():
    """AdpNN has no closed-form second-derivative kernel (csrc/hessian.cu refuses ADP): the
    calculator serves `hessian` by fourth-order differences of the ANALYTIC forces on fixed
    lists (BasicNN._hessian_from_forces) -- against the oracle's autograd Hessian of the ADP
    energy (tf.hessians of the reference works for every model, basic.py:410-421)."""
    from tensoralloy_b200.nn.eam import AdpNN
    om = opot.get_potential('mishinh')
    oz = opot.get_potential('zjw04')
    fns = {'rho': oz.rho, 'phi': oz.phi, 'embed': oz.embed,
           'dipole': om.dipole, 'quadrupole': om.quadrupole}
    cp = {'Ni': {'rho': 'zjw04', 'embed': 'zjw04'},
          'NiNi': {'phi': 'zjw04', 'dipole': 'mishinh', 'quadrupole': 'mishinh'}}
    atoms = bulk_fcc('Ni', 3.52, (2, 2, 2))
    atoms.positions += np.random.default_rng(11).normal(scale=0.05, size=atoms.positions.shape)
    with precision_scope('high'):
        nn = AdpNN(['Ni'], custom_potentials=cp,
                   export_properties=['energy', 'forces', 'hessian'])
        nn.attach_transformer(UniversalTransformer(['Ni'], rcut=5.0))
        calc = TensorAlloyCalculator(nn)
        H = calc.get_hessian(atoms)
    n = len(atoms)
    ref = oeam.eam_evaluate(None, 'adp', ['Ni'], atoms.get_chemical_symbols(), atoms.positions,
                            atoms.cell, [1, 1, 1], 5.0, hessian=True, fns=fns)
    Href = ref['hessian'].reshape(3 * n, 3 * n)
    assert np.abs(H - H.T).max() < 1e-12
    assert np.abs(H - Href).max() < 1e-7 * max(np.abs(Href).max(), 1.0)


def test_atomic_nn_hessian_from_analytic_forces():
    """AtomicNN (G2 + G4 + MLP) Hessian through the calculator vs the oracle's autograd
    Hessian, Be hcp 3x3x2 (the reference's phonon workflow runs on NN potentials,
    analysis/phonon.py:520-534)."""
    from oracle import atomic as oat
    from tensoralloy_b200.nn.atomic import AtomicNN, SymmetryFunction
    atoms = bulk_hcp('Be', 2.2644, 3.5673, (3, 3, 2))
    atoms.positions += np.random.default_rng(4).normal(scale=0.02, size=atoms.positions.shape)
    elements = ['Be']
    with precision_scope('high'):
        clf = UniversalTransformer(elements, rcut=4.5, angular=True)
        nn = AtomicNN(elements, SymmetryFunction(elements), minmax_scale=False,
                      export_properties=('energy', 'forces', 'hessian'))
        nn.attach_transformer(clf)
        nn.initialize_variables(seed=611)
        key = "Atomic/Be/Output/kernel"
        nn.set_variable(key, nn.get_variable(key) * 0.02)
        calc = TensorAlloyCalculator(nn)
        H = calc.get_hessian(atoms)
    params = {'Be': nn.mlp_params('Be')}
    sfd = nn.descriptor.as_dict()
    sf = {k: tuple(sfd[k]) for k in ('eta', 'omega', 'beta', 'gamma', 'zeta')}
    sf['cutoff'] = sfd['cutoff_function']
    ref = oat.atomic_evaluate(elements, atoms.get_chemical_symbols(), atoms.positions,
                              atoms.cell, atoms.pbc, 4.5, params, sf=sf, angular=True,
                              minmax={'Be': None}, hessian=True)
    n = len(atoms)
    Href = ref['hessian'].reshape(3 * n, 3 * n)
    assert np.abs(H - Href).max() < 1e-7 * max(np.abs(Href).max(), 1.0)
    assert np.abs(H.reshape(n, 3, n, 3).sum(axis=2)).max() < 1e-7      # acoustic sum rule
