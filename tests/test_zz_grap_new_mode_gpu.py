"""GenericRadialAtomicPotential, NEW mode (nn/atomic/grap.py:470-680: multiplicity tensor
T_dm, moment coefficients M_dnac, m = 0 entry sign(P) sqrt(P^2 + 1e-16), moment 3, traceless
`symmetric=True` option) on the GPU vs the oracle's independent restatement
(`oracle.atomic.grap_descriptors_new_mode`).  Tolerances: BASELINE.json north_star
(1e-10 eV/atom, 1e-8 eV/A float64; 1e-5 relative float32)."""
import os

import numpy as np
import pytest
import torch

from oracle import atomic as oat
from oracle import neighbor as onl
from oracle import training as otr
from tensoralloy_b200.atoms import Atoms, bulk_fcc
from tensoralloy_b200.calculator import TensorAlloyCalculator
from tensoralloy_b200.nn.atomic import AtomicNN, GenericRadialAtomicPotential
from tensoralloy_b200.nn.atomic.training import AtomicNNTrainer
from tensoralloy_b200.precision import precision_scope
from tensoralloy_b200.transformer import UniversalTransformer

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), 'golden')


def _compare(atoms, elements, rc, algorithm, parameters, max_moment, symmetric,
             cutoff='cosine', seed=611, medium=True):
    with precision_scope('high'):
        clf = UniversalTransformer(elements, rcut=rc, angular=False)
        desc = GenericRadialAtomicPotential(elements, algorithm=algorithm,
                                            parameters=parameters, moment_tensors=max_moment,
                                            cutoff_function=cutoff, symmetric=symmetric,
                                            legacy_mode=False)
        nn = AtomicNN(elements, desc, export_properties=('energy', 'forces', 'stress'),
                      minmax_scale=False)
        nn.attach_transformer(clf)
        nn.initialize_variables(seed=seed)
        for el in nn.elements:
            key = f"Atomic/{el}/Output/kernel"
            nn.set_variable(key, nn.get_variable(key) * 0.05)
        calc = TensorAlloyCalculator(nn)
        calc.calculate(atoms, properties=['energy', 'forces', 'stress'])
        e, f, s = calc.results['energy'], calc.get_forces(atoms), calc.get_stress(atoms)
        g = nn.get_descriptors(clf.get_constant_features(atoms))
    params = {el: nn.mlp_params(el) for el in nn.elements}
    grap = dict(algorithm=algorithm, grid=desc.radial_sets(), moments=desc.moments(),
                cutoff=cutoff, new_mode=True, symmetric=symmetric)
    ref = oat.atomic_evaluate(elements, atoms.get_chemical_symbols(), atoms.positions,
                              atoms.cell, atoms.pbc, rc, params, angular=False, grap=grap)
    nl = onl.neighbor_list(atoms.positions, atoms.cell, atoms.pbc, rc)
    types = np.array([sorted(elements).index(x) for x in atoms.get_chemical_symbols()])
    G = oat.grap_descriptors_new_mode(
        elements, types, torch.tensor(atoms.positions), torch.tensor(np.asarray(atoms.cell)),
        nl[0], nl[1], nl[2], rc, algorithm, desc.radial_sets(), max_moment, cutoff,
        symmetric).numpy()
    n = len(atoms)
    assert g.shape == G.shape == (n, len(elements) * len(desc.grid) * (max_moment + 1))
    print(algorithm, max_moment, symmetric, 'dG', np.abs(g - G).max(),
          'dE/N', abs(e - ref['energy']) / n, 'dF', np.abs(f - ref['forces']).max(),
          'Fmax', np.abs(ref['forces']).max())
    assert np.abs(g - G).max() < 1e-10 * max(1.0, np.abs(G).max())
    assert abs(e - ref['energy']) / n < 1e-10
    assert np.abs(f - ref['forces']).max() < 1e-8
    assert np.abs(s - ref['stress']).max() < 1e-8
    if not medium:
        return ref, G
    with precision_scope('medium'):
        calc32 = TensorAlloyCalculator(nn)
        calc32.calculate(atoms, properties=['energy', 'forces'])
        e32, f32 = calc32.results['energy'], calc32.get_forces(atoms)
    assert abs(e32 - ref['energy']) <= 2e-5 * max(abs(ref['energy']), 1.0)
    assert np.abs(f32 - ref['forces']).max() <= 1e-3 * max(np.abs(ref['forces']).max(), 1e-2)
    return ref, G


def test_grap_new_mode_moments_and_traceless_form():
    d = np.load(os.path.join(GOLD, 'Be_liquid_4000K.npz'))
    be = Atoms(list(d['symbols']), d['positions'][2], d['cells'][2], True)
    # moment 0 only: the signed-square-root entry alone
    _compare(be, ['Be'], 5.0, 'sf', dict(eta=[0.05, 4.0, 20.0], omega=[0.0, 0.5, 1.0]), 0, False)
    # morse sums change sign between parameter sets: sign(P) matters
    ref, G = _compare(be, ['Be'], 5.0, 'morse',
                      dict(D=[0.05, 0.1, 0.1], gamma=[1.2, 0.8, 1.0], r0=[2.2, 2.6, 3.4]),
                      2, True, cutoff='polynomial', medium=False)   # cancelling sums: float64 only
    assert (G[:, 0::3] < 0).any() and (G[:, 0::3] > 0).any()
    assert np.abs(ref['forces']).max() > 1e-3
    # moment 3, plain and traceless
    _compare(be, ['Be'], 5.0, 'pexp', dict(rl=[1.0, 1.5, 2.0, 2.5], pl=[1.0, 2.0, 3.0, 2.5]),
             3, False)
    _compare(be, ['Be'], 5.0, 'density', dict(A=[1.0, 2.0], beta=[3.0, 5.0], re=[2.2, 2.4]),
             3, True)
    # two species, mixed periodicity
    pd3o2 = Atoms('Pd3O2', pbc=[True, True, False],
                  cell=[[7.78, 0., 0.], [0., 5.50129076, 0.], [0., 0., 15.37532269]],
                  positions=[[3.89, 0., 8.37532269], [0., 2.75064538, 8.37532269],
                             [3.89, 2.75064538, 8.37532269], [5.835, 1.37532269, 8.5],
                             [5.835, 7.12596807, 8.]])
    _compare(pd3o2, ['O', 'Pd'], 6.5, 'pexp', dict(rl=[1.5, 2.5], pl=[2.0, 3.0]), 3, True)


def test_grap_new_mode_parameter_gradients_match_oracle():
    """Training step with moment 3 and the traceless form: the force op and the JVP kernel
    carry the third-order and trace terms (loss = energy + forces + stress)."""
    rng = np.random.default_rng(5)
    structs = []
    for k in range(2):
        base = bulk_fcc('Ni', 3.3 + 0.4 * rng.random(), (2, 2, 2))
        sym = ['Mo' if x < 0.4 else 'Ni' for x in rng.random(len(base))]
        pos = base.positions + rng.normal(scale=0.1, size=base.positions.shape)
        structs.append(dict(atoms=Atoms(sym, pos, base.cell, True), symbols=sym, positions=pos,
                            cell=base.cell, pbc=[1, 1, 1],
                            energy=-4.0 * len(base) + rng.normal(),
                            forces=rng.normal(scale=0.3, size=pos.shape),
                            stress=rng.normal(scale=0.01, size=6)))
    elements = ['Mo', 'Ni']
    rc = 4.5
    with precision_scope('high'):
        desc = GenericRadialAtomicPotential(elements, algorithm='pexp',
                                            parameters=dict(rl=[1.5, 2.5], pl=[2.0, 3.0]),
                                            moment_tensors=3, symmetric=True,
                                            legacy_mode=False)
        nn = AtomicNN(elements, desc, hidden_sizes=[16, 16], activation='softplus',
                      minmax_scale=False, minimize_properties=('energy', 'forces', 'stress'),
                      export_properties=('energy', 'forces', 'stress'))
        nn.attach_transformer(UniversalTransformer(elements, rcut=rc, angular=False))
        nn.initialize_variables(seed=3)
        for el in elements:
            key = f"Atomic/{el}/Output/kernel"
            nn.set_variable(key, nn.get_variable(key) * 0.05)
        tr = AtomicNNTrainer(nn)
        for s in structs:
            tr.add_structure(s['atoms'], s['energy'], s['forces'], s['stress'])
        loss, parts = tr.gradients()
        params = {el: nn.mlp_params(el) for el in elements}
        grap = dict(algorithm='pexp', grid=desc.radial_sets(), moments=desc.moments(),
                    cutoff='cosine', new_mode=True, symmetric=True)
        ref_loss, ref_parts, ref_g = otr.loss_and_grads(elements, structs, params, rc,
                                                        angular=False, grap=grap)
        assert abs(loss.item() - ref_loss) < 1e-9 * max(1.0, abs(ref_loss))
        for key in ('energy', 'forces', 'stress'):
            assert abs(parts[key].item() - ref_parts[key]) < 1e-9
        for el in elements:
            L = tr.layers[el]
            for k, w in enumerate(L['W']):
                g = w.grad.cpu().numpy()
                r = ref_g[el][0][k]
                assert np.abs(g - r).max() < 1e-8 * max(1.0, np.abs(r).max()), (el, k)
