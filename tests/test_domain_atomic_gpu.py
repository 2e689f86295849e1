"""Slab decomposition of AtomicNN (2 rc halo, inner-halo rows recomputed, one position
exchange; tensoralloy_b200/domain_atomic.py): all ranks run in one process on one GPU and
must reproduce the single-domain E / per-atom E / forces / virial."""
import numpy as np
import pytest

from tensoralloy_b200.atoms import Atoms, bulk_fcc
from tensoralloy_b200.calculator import TensorAlloyCalculator
from tensoralloy_b200.domain_atomic import AtomicSlabLayout, run_loopback
from tensoralloy_b200.nn.atomic import AtomicNN, SymmetryFunction
from tensoralloy_b200.precision import precision_scope
from tensoralloy_b200.transformer import UniversalTransformer

pytestmark = pytest.mark.gpu


def _model(elements, rc, acut, angular):
    nn = AtomicNN(elements, SymmetryFunction(elements), minmax_scale=False,
                  hidden_sizes=[32, 16], export_properties=('energy', 'forces', 'stress'))
    nn.attach_transformer(UniversalTransformer(elements, rcut=rc, acut=acut, angular=angular))
    nn.initialize_variables(seed=7)
    for el in nn.elements:
        key = f"Atomic/{el}/Output/kernel"
        nn.set_variable(key, nn.get_variable(key) * 0.02)
    return nn


@pytest.mark.parametrize("world", [1, 2, 3])
def test_atomic_nn_slab_decomposition_matches_single_domain(world):
    rng = np.random.default_rng(9)
    base = bulk_fcc('Ni', 3.6, (8, 3, 3))
    sym = ['Mo' if x < 0.4 else 'Ni' for x in rng.random(len(base))]
    pos = base.positions + rng.normal(scale=0.1, size=base.positions.shape)
    pos[:, 0] += 1.234                 # atoms on both sides of the periodic x boundary
    atoms = Atoms(sym, pos, base.cell, True)
    elements, rc, acut = ['Mo', 'Ni'], 4.6, 4.0
    with precision_scope('high'):
        nn = _model(elements, rc, acut, True)
        calc = TensorAlloyCalculator(nn)
        calc.calculate(atoms, properties=['energy', 'forces', 'stress'])
        e0, f0 = calc.results['energy'], calc.get_forces(atoms)
        w0 = calc.results['virial']
        ea0 = calc.get_atomic(atoms) if hasattr(calc, 'get_atomic') else None
        types = nn.transformer.get_types(atoms)
        e, f, w, ea = run_loopback(nn._device_model(), pos, types, np.asarray(atoms.cell),
                                   max(rc, acut), world, precision=0)
    n = len(atoms)
    assert abs(e - e0) / n < 1e-12
    assert np.abs(f - f0).max() < 1e-10
    assert np.abs(w - w0).max() / n < 1e-11
    assert np.abs(f0).max() > 1e-2
    if ea0 is not None:
        assert np.abs(ea - ea0).max() < 1e-11


@pytest.mark.parametrize("world", [2, 3])
def test_adp_and_eam_recompute_decomposition(world):
    """ADP (two species: per-term moments) and EAM through the same 2 rc / recompute path
    (tab_eam_eval_dd): no F' or moment exchange."""
    from tensoralloy_b200.nn.eam import AdpNN, EamAlloyNN
    rng = np.random.default_rng(21)
    base = bulk_fcc('Ni', 3.6, (11, 3, 3))
    sym = ['Mo' if x < 0.45 else 'Ni' for x in rng.random(len(base))]
    pos = base.positions + rng.normal(scale=0.08, size=base.positions.shape)
    pos[:, 0] -= 0.777
    atoms = Atoms(sym, pos, base.cell, True)
    rc = 6.0
    cp = {'Mo': {'rho': 'zjw04', 'embed': 'zjw04'}, 'Ni': {'rho': 'zjw04', 'embed': 'zjw04'},
          'MoMo': {'phi': 'zjw04', 'dipole': 'mishinh', 'quadrupole': 'mishinh'},
          'MoNi': {'phi': 'zjw04', 'dipole': 'mishinh', 'quadrupole': 'mishinh'},
          'NiNi': {'phi': 'zjw04', 'dipole': 'mishinh', 'quadrupole': 'mishinh'}}
    with precision_scope('high'):
        for nn in (AdpNN(['Mo', 'Ni'], custom_potentials=cp,
                         export_properties=('energy', 'forces', 'stress')),
                   EamAlloyNN(['Mo', 'Ni'], custom_potentials='zjw04',
                              export_properties=('energy', 'forces', 'stress'))):
            nn.attach_transformer(UniversalTransformer(['Mo', 'Ni'], rcut=rc))
            calc = TensorAlloyCalculator(nn)
            calc.calculate(atoms, properties=['energy', 'forces', 'stress'])
            e0, f0, w0 = calc.results['energy'], calc.get_forces(atoms), calc.results['virial']
            types = nn.transformer.get_types(atoms)
            e, f, w, _ = run_loopback(nn._device_model(), pos, types, np.asarray(atoms.cell),
                                      rc, world, precision=0)
            n = len(atoms)
            assert abs(e - e0) / n < 1e-12, type(nn).__name__
            assert np.abs(f - f0).max() < 1e-10, type(nn).__name__
            assert np.abs(w - w0).max() / n < 1e-11, type(nn).__name__
            assert np.abs(f0).max() > 1e-2


def test_layout_rejects_too_narrow_slabs():
    with pytest.raises(ValueError):
        AtomicSlabLayout(20.0, 4, 0, 4.6)      # width 5 < 2 rc
