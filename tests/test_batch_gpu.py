"""Batches of structures in ONE neighbour handle (tab_nbr_build_batch; BASELINE configs 2
and 4: "batched E/F/stress inference", "structure-parallel batch"):
  * the batch lists equal the ASE restatement per structure, bit-exact after canonical sort
    (the reference's batch path still builds each list with ASE, universal.py:58);
  * E / F / stress of every structure of a batch equal the single-structure call
    (same kernels, same summation order inside an atom row up to the candidate order:
    1e-12) and the oracle (1e-10 eV/atom, 1e-8 eV/A);
  * EAM (one and two species), ADP and AtomicNN G2+G4 models, float64 and float32."""
import os

import numpy as np
import pytest
import torch

from oracle import eam as oeam
from oracle import neighbor as onl
from oracle import potentials as opot
from tensoralloy_b200 import _lib
from tensoralloy_b200.atoms import Atoms, bulk_fcc
from tensoralloy_b200.calculator import TensorAlloyCalculator
from tensoralloy_b200.nn.atomic import AtomicNN, SymmetryFunction
from tensoralloy_b200.nn.eam import AdpNN, EamAlloyNN
from tensoralloy_b200.precision import precision_scope
from tensoralloy_b200.transformer import UniversalTransformer

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), 'golden')


def _rattled(el, a, reps, seed, scale=0.05, symbols=None):
    atoms = bulk_fcc(el, a, reps)
    rng = np.random.default_rng(seed)
    pos = atoms.positions + rng.normal(scale=scale, size=atoms.positions.shape)
    return Atoms(symbols or atoms.get_chemical_symbols(), pos, atoms.cell, True)


def test_batch_lists_match_ase_restatement():
    rng = np.random.default_rng(7)
    tri = np.array([[9.0, 0.0, 0.0], [2.5, 8.0, 0.0], [1.0, -1.5, 10.0]])
    a3 = bulk_fcc('Ni', 3.52, (3, 3, 3))
    far = a3.positions + rng.normal(scale=0.05, size=a3.positions.shape)
    far += rng.integers(-2, 3, size=far.shape) @ a3.cell       # atoms outside the cell
    structs = [
        (_rattled('Ni', 3.52, (4, 4, 4), 611).positions, bulk_fcc('Ni', 3.52, (4, 4, 4)).cell,
         [1, 1, 1]),
        (bulk_fcc('Ni', 3.52, (2, 2, 2)).positions, bulk_fcc('Ni', 3.52, (2, 2, 2)).cell,
         [1, 1, 1]),                                            # rc > L/2: several images
        (bulk_fcc('Ni', 3.52, (1, 1, 1)).positions, bulk_fcc('Ni', 3.52, (1, 1, 1)).cell,
         [1, 1, 1]),                                            # rc spans two images
        (rng.random((120, 3)) @ tri, tri, [1, 1, 1]),
        (rng.random((90, 3)) @ tri, tri, [1, 1, 0]),
        (rng.random((70, 3)) @ tri, tri, [0, 0, 0]),
        (far, a3.cell, [1, 1, 1]),
    ]
    rc = 6.5
    offsets = np.concatenate(([0], np.cumsum([len(p) for p, _, _ in structs]))).astype(np.int32)
    pos = np.concatenate([p for p, _, _ in structs])
    nl = _lib.NeighborList()
    nl.build_batch(torch.tensor(pos, dtype=torch.float64, device='cuda'), None, offsets,
                   np.stack([np.asarray(c, dtype=float).reshape(3, 3) for _, c, _ in structs]),
                   np.array([p for _, _, p in structs]), rc)
    gi, gj, gS = (t.cpu().numpy() for t in nl.export())
    counts = nl.counts().cpu().numpy()
    total = 0
    for s, (p, cell, pbc) in enumerate(structs):
        ri, rj, rS, _, _ = onl.neighbor_list_brute(p, cell, pbc, rc)
        lo, hi = offsets[s], offsets[s + 1]
        sel = (gi >= lo) & (gi < hi)
        assert ((gj[sel] >= lo) & (gj[sel] < hi)).all()        # no cross-structure pairs
        bi, bj, bS = onl.canonical_sort(gi[sel] - lo, gj[sel] - lo, gS[sel])
        assert len(bi) == len(ri), (s, len(bi), len(ri))
        np.testing.assert_array_equal(bi, ri)
        np.testing.assert_array_equal(bj, rj)
        np.testing.assert_array_equal(bS, rS)
        np.testing.assert_array_equal(counts[lo:hi], np.bincount(ri, minlength=len(p)))
        total += len(ri)
    assert nl.sizes()[0] == total


def _single_vs_batch(nn_factory, clf_factory, images, tol_e=1e-12, tol_f=1e-11,
                     oracle=None):
    with precision_scope('high'):
        nn = nn_factory()
        nn.attach_transformer(clf_factory())
        calc = TensorAlloyCalculator(nn)
        batch = calc.calculate_batch(images, properties=('energy', 'forces', 'stress'))
        assert len(batch) == len(images)
        for s, atoms in enumerate(images):
            calc.calculate(atoms, properties=['energy', 'forces', 'stress'])
            e, f, st = calc.results['energy'], calc.get_forces(atoms), calc.get_stress(atoms)
            ea = calc.get_atomic(atoms) if hasattr(calc, 'get_atomic') else None
            n = len(atoms)
            b = batch[s]
            assert abs(b['energy'] - e) / n < tol_e, (s, b['energy'], e)
            assert np.abs(b['forces'] - f).max() < tol_f, s
            assert np.abs(b['stress'] - st).max() < tol_f, s
            if ea is not None:
                assert np.abs(b['energy/atom'] - ea).max() < tol_e * 10
            if oracle is not None:
                ref = oracle(atoms)
                assert abs(b['energy'] - ref['energy']) / n < 1e-10
                assert np.abs(b['forces'] - ref['forces']).max() < 1e-8
                assert np.abs(b['stress'] - ref['stress']).max() < 1e-8
    with precision_scope('medium'):
        calc32 = TensorAlloyCalculator(nn)
        b32 = calc32.calculate_batch(images, properties=('energy', 'forces'))
        for s in range(len(images)):
            assert abs(b32[s]['energy'] - batch[s]['energy']) <= \
                2e-5 * max(abs(batch[s]['energy']), 1.0)
            fs = max(np.abs(batch[s]['forces']).max(), 1e-2)
            assert np.abs(b32[s]['forces'] - batch[s]['forces']).max() <= 1e-3 * fs
    return batch


def test_eam_ni_batch_of_unequal_structures():
    images = [_rattled('Ni', 3.52, (4, 4, 4), 611), _rattled('Ni', 3.55, (3, 3, 3), 2),
              _rattled('Ni', 3.50, (2, 2, 2), 3), _rattled('Ni', 3.60, (3, 3, 3), 4),
              _rattled('Ni', 3.52, (5, 5, 5), 5)]          # 256, 108, 32, 108, 500 atoms
    pot = opot.get_potential('zjw04')

    def oracle(atoms):
        return oeam.eam_evaluate(pot, 'alloy', ['Ni'], atoms.get_chemical_symbols(),
                                 atoms.positions, atoms.cell, [1, 1, 1], 6.5)
    _single_vs_batch(lambda: EamAlloyNN(['Ni'], custom_potentials='zjw04',
                                        export_properties=('energy', 'forces', 'stress')),
                     lambda: UniversalTransformer(['Ni'], rcut=6.5), images, oracle=oracle)


def test_eam_two_species_and_adp_batches():
    rng = np.random.default_rng(13)
    images = []
    for k in range(4):
        base = _rattled('Ni', 3.55 + 0.03 * k, (3, 3, 3), 20 + k, scale=0.08)
        sym = ['Mo' if x < 0.45 else 'Ni' for x in rng.random(len(base))]
        images.append(Atoms(sym, base.positions, base.cell, True))
    _single_vs_batch(lambda: EamAlloyNN(['Mo', 'Ni'], custom_potentials='zjw04',
                                        export_properties=('energy', 'forces', 'stress')),
                     lambda: UniversalTransformer(['Mo', 'Ni'], rcut=6.0), images)
    cp2 = {'Mo': {'rho': 'zjw04', 'embed': 'zjw04'},
           'Ni': {'rho': 'zjw04', 'embed': 'zjw04'},
           'MoMo': {'phi': 'zjw04', 'dipole': 'mishinh', 'quadrupole': 'mishinh'},
           'MoNi': {'phi': 'zjw04', 'dipole': 'mishinh', 'quadrupole': 'mishinh'},
           'NiNi': {'phi': 'zjw04', 'dipole': 'mishinh', 'quadrupole': 'mishinh'}}
    _single_vs_batch(lambda: AdpNN(['Mo', 'Ni'], custom_potentials=cp2,
                                   export_properties=('energy', 'forces', 'stress')),
                     lambda: UniversalTransformer(['Mo', 'Ni'], rcut=6.0), images)


def test_atomic_nn_batch_be_liquid():
    d = np.load(os.path.join(GOLD, 'Be_liquid_4000K.npz'))
    images = [Atoms(list(d['symbols']), d['positions'][k], d['cells'][k], True)
              for k in (2, 1, 0, 2)]
    rng = np.random.default_rng(1)
    images[3] = Atoms(list(d['symbols']),
                      d['positions'][2] + rng.normal(scale=0.05, size=(128, 3)),
                      d['cells'][2], True)

    def factory():
        return AtomicNN(['Be'], SymmetryFunction(['Be']), minmax_scale=False,
                        export_properties=('energy', 'forces', 'stress'))

    def clf():
        return UniversalTransformer(['Be'], rcut=5.0, acut=5.0, angular=True)

    batch = _single_vs_batch(_scaled(factory), clf, images)
    assert np.abs(batch[0]['forces']).max() > 0.05


def _scaled(factory):
    """Bring random he_normal weights to a physical energy scale (test_atomic_gpu)."""
    def wrapped():
        nn = factory()
        orig_attach = nn.attach_transformer

        def attach(clf):
            orig_attach(clf)
            nn.initialize_variables(seed=611)
            for el in nn.elements:
                key = f"Atomic/{el}/Output/kernel"
                nn.set_variable(key, nn.get_variable(key) * 0.02)
        nn.attach_transformer = attach
        return nn
    return wrapped


def test_atomic_nn_two_species_batch():
    rng = np.random.default_rng(5)
    images = []
    for k in range(3):
        base = bulk_fcc('Ni', 3.5 + 0.1 * k, (2, 2, 2) if k else (3, 3, 3))
        sym = ['Mo' if x < 0.4 else 'Ni' for x in rng.random(len(base))]
        images.append(Atoms(sym, base.positions + rng.normal(scale=0.1,
                                                              size=base.positions.shape),
                            base.cell, True))
    _single_vs_batch(
        _scaled(lambda: AtomicNN(['Mo', 'Ni'], SymmetryFunction(['Mo', 'Ni']),
                                 minmax_scale=False, hidden_sizes=[32, 32],
                                 export_properties=('energy', 'forces', 'stress'))),
        lambda: UniversalTransformer(['Mo', 'Ni'], rcut=4.6, acut=4.0, angular=True), images)


def test_batch_edge_cases_and_errors():
    """Ragged / degenerate inputs: a one-structure batch, an isolated atom (no neighbours),
    a non-periodic cluster next to periodic crystals; error behaviour of the reference
    (unsupported element -> ValueError, universal.py:280-282) and of the C ABI."""
    with precision_scope('high'):
        nn = EamAlloyNN(['Ni'], custom_potentials='zjw04',
                        export_properties=('energy', 'forces', 'stress'))
        nn.attach_transformer(UniversalTransformer(['Ni'], rcut=6.5))
        calc = TensorAlloyCalculator(nn)
        crystal = _rattled('Ni', 3.52, (3, 3, 3), 31)
        lonely = Atoms(['Ni'], [[5.0, 5.0, 5.0]], np.diag([30.0, 30.0, 30.0]), True)
        rng = np.random.default_rng(3)
        cluster = Atoms(['Ni'] * 13, 2.2 * rng.normal(size=(13, 3)), np.zeros((3, 3)), False)
        images = [crystal, lonely, cluster]
        res = calc.calculate_batch(images, properties=('energy', 'forces', 'stress'))
        for s, atoms in enumerate(images):
            calc.calculate(atoms, properties=['energy', 'forces'])
            assert abs(res[s]['energy'] - calc.results['energy']) < 1e-11 * max(1, len(atoms))
            assert np.abs(res[s]['forces'] - calc.get_forces(atoms)).max() < 1e-11
        assert np.abs(res[1]['forces']).max() == 0.0          # no neighbours at all
        one = calc.calculate_batch([crystal], properties=('energy', 'forces'))
        assert abs(one[0]['energy'] - res[0]['energy']) < 1e-11
        with pytest.raises(ValueError):
            calc.calculate_batch([crystal, Atoms(['Cu'], [[0, 0, 0]], np.eye(3) * 9, True)])
        with pytest.raises(KeyError):
            calc.calculate_batch([crystal], properties=('energy', 'dipole_moment'))
    # C ABI: empty structure, update on a batch handle
    nl = _lib.NeighborList()
    pos = torch.tensor(crystal.positions, dtype=torch.float64, device='cuda')
    with pytest.raises(_lib.TabError):
        nl.build_batch(pos, None, np.array([0, 0, len(crystal)], dtype=np.int32),
                       np.stack([crystal.cell, crystal.cell]), np.ones((2, 3), bool), 6.5)
    nl.build_batch(pos, None, np.array([0, len(crystal)], dtype=np.int32),
                   crystal.cell[None], np.ones((1, 3), bool), 6.5)
    with pytest.raises(_lib.TabError):
        nl.update(pos)


def test_finite_temperature_batch_matches_single_calls():
    from tensoralloy_b200.nn.atomic import BeNN, TemperatureDependentAtomicNN
    d = np.load(os.path.join(GOLD, 'Be_liquid_4000K.npz'))
    images = []
    for k, temp in ((2, 0.34469373), (1, 0.2), (2, 0.05)):
        a = Atoms(list(d['symbols']), d['positions'][k], d['cells'][k], True)
        a.info['etemperature'] = temp
        images.append(a)
    props = ('energy', 'forces', 'stress', 'eentropy', 'free_energy')
    with precision_scope('high'):
        for cls in (TemperatureDependentAtomicNN, BeNN):
            nn = cls(['Be'], SymmetryFunction(['Be']), hidden_sizes=[16, 16],
                     minmax_scale=False, export_properties=props,
                     finite_temperature=dict(layers=[16, 8]))
            nn.attach_transformer(UniversalTransformer(['Be'], rcut=5.0, angular=True))
            nn.initialize_variables(seed=4)
            for head in ('U', 'S'):
                key = f"TD/Be/{head}/Output/kernel"
                nn.set_variable(key, nn.get_variable(key) * 0.2)
            key = "TD/Be/H/Conv1d1/kernel"
            nn.set_variable(key, nn.get_variable(key) * 0.2)
            calc = TensorAlloyCalculator(nn)
            batch = calc.calculate_batch(images, properties=props)
            for s, atoms in enumerate(images):
                calc.calculate(atoms, properties=list(props))
                for key in ('energy', 'eentropy', 'free_energy'):
                    assert abs(batch[s][key] - calc.results[key]) < 1e-10, (cls.__name__, key)
                assert np.abs(batch[s]['forces'] - calc.get_forces(atoms)).max() < 1e-11
                assert np.abs(batch[s]['stress'] - calc.get_stress(atoms)).max() < 1e-11
            assert abs(batch[0]['eentropy'] - batch[2]['eentropy']) > 1e-6   # T matters
