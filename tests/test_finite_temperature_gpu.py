"""TemperatureDependentAtomicNN / BeNN on the GPU vs the oracle
restatement of nn/atomic/finite_temperature.py:92-304 and special/beryllium.py:23-77.
The reference holds no numeric golden for these heads (its tests only build the graph,
nn/atomic/tests/test_finite_temperature.py), so parity is oracle <-> GPU on the Be liquid
frames that carry `etemperature` (test_files/Be_liquid_4000K_TS.extxyz).
Tolerances: 1e-10 eV/atom, 1e-8 eV/A (float64); 1e-5 relative (float32)."""
import os

import numpy as np
import pytest

from oracle import atomic as oat
from tensoralloy_b200.atoms import Atoms, bulk_fcc
from tensoralloy_b200.calculator import TensorAlloyCalculator
from tensoralloy_b200.nn.atomic import SymmetryFunction
from tensoralloy_b200.nn.atomic.finite_temperature import (BeNN,
                                                           TemperatureDependentAtomicNN)
from tensoralloy_b200.precision import precision_scope
from tensoralloy_b200.transformer import UniversalTransformer

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), 'golden')


def _compare(cls, atoms, elements, rc, etemp, nn_kwargs=None, ft=None, angular=True,
             seed=611):
    atoms.info['etemperature'] = etemp
    nn_kwargs = nn_kwargs or {}
    with precision_scope('high'):
        clf = UniversalTransformer(elements, rcut=rc, angular=angular)
        nn = cls(elements, SymmetryFunction(elements),
                 export_properties=('energy', 'forces', 'stress', 'eentropy',
                                    'free_energy'),
                 finite_temperature=ft or dict(activation='softplus', layers=[32, 16]),
                 **nn_kwargs)
        nn.attach_transformer(clf)
        nn.initialize_variables(seed=seed)
        for el in nn.elements:       # physical scale (see test_atomic_gpu._compare)
            for head in ('U', 'S'):
                key = f"TD/{el}/{head}/Output/kernel"
                nn.set_variable(key, nn.get_variable(key) * 0.2)
            key = f"TD/{el}/H/Conv1d1/kernel"
            nn.set_variable(key, nn.get_variable(key) * 0.2)
        if nn_kwargs.get('minmax_scale', True):
            rng = np.random.default_rng(seed + 1)
            dim = nn._dim()
            for el in nn.elements:
                lo = rng.random(dim) * 0.1
                nn.set_variable(f"TD/{el}/xlo", lo.reshape(1, 1, -1))
                nn.set_variable(f"TD/{el}/xhi",
                                (lo + 1.0 + 5 * rng.random(dim)).reshape(1, 1, -1))
        assert nn.is_finite_temperature and nn.variational_energy == 'free_energy'
        calc = TensorAlloyCalculator(nn)
        calc.calculate(atoms, properties=['energy', 'forces', 'stress', 'eentropy',
                                          'free_energy'])
        res = dict(calc.results)
        f, s = calc.get_forces(atoms), calc.get_stress(atoms)
    sfd = nn.descriptor.as_dict()
    osf = {k: tuple(sfd[k]) for k in ('eta', 'omega', 'beta', 'gamma', 'zeta')}
    osf['cutoff'] = sfd['cutoff_function']
    params = {el: nn.td_params(el) for el in nn.elements}
    minmax = {el: nn.minmax(el) for el in nn.elements}
    ref = oat.td_atomic_evaluate(elements, atoms.get_chemical_symbols(), atoms.positions,
                                 atoms.cell, atoms.pbc, rc, params, etemp, sf=osf,
                                 angular=angular, minmax=minmax)
    n = len(atoms)
    print('U/N', ref['energy'] / n, 'S/N', ref['eentropy'] / n, 'F/N',
          ref['free_energy'] / n, 'dU', abs(res['energy'] - ref['energy']) / n, 'dS',
          abs(res['eentropy'] - ref['eentropy']) / n, 'dF',
          abs(res['free_energy'] - ref['free_energy']) / n, 'dforces',
          np.abs(f - ref['forces']).max(), 'Fmax', np.abs(ref['forces']).max())
    assert abs(res['energy'] - ref['energy']) / n < 1e-10
    assert abs(res['eentropy'] - ref['eentropy']) / n < 1e-10
    assert abs(res['free_energy'] - ref['free_energy']) / n < 1e-10
    assert np.abs(f - ref['forces']).max() < 1e-8
    assert np.abs(s - ref['stress']).max() < 1e-8
    assert np.abs(ref['forces']).max() > 1e-3      # the comparison is not vacuous
    # forces are the gradient of the FREE energy, not of U (basic.py:190-202)
    assert abs(ref['eentropy']) > 1e-3
    with precision_scope('medium'):
        c32 = TensorAlloyCalculator(nn)
        c32.calculate(atoms, properties=['energy', 'forces', 'free_energy'])
        f32 = c32.get_forces(atoms)
        assert abs(c32.results['free_energy'] - ref['free_energy']) <= \
            2e-5 * max(abs(ref['free_energy']), 1.0)
        assert np.abs(f32 - ref['forces']).max() <= 1e-3 * max(np.abs(ref['forces']).max(),
                                                               1e-2)
    return ref


def _be(frame):
    d = np.load(os.path.join(GOLD, 'Be_liquid_4000K.npz'))
    return Atoms(list(d['symbols']), d['positions'][frame], d['cells'][frame], True)


def test_td_atomic_nn_be_liquid():
    # etemperature of the fixture frames (extxyz header): 0.34469373 eV
    _compare(TemperatureDependentAtomicNN, _be(2), ['Be'], 5.0, 0.34469373,
             nn_kwargs=dict(minmax_scale=False, hidden_sizes=[32, 32]))
    _compare(TemperatureDependentAtomicNN, _be(1), ['Be'], 5.0, 0.34469373,
             nn_kwargs=dict(minmax_scale=True, use_resnet_dt=True, hidden_sizes=[32, 32],
                            activation='tanh'),
             ft=dict(activation='tanh', layers=[32, 32, 8], algo='Sommerfeld'))


def test_benn_special_entropy():
    _compare(BeNN, _be(2), ['Be'], 5.0, 0.34469373,
             nn_kwargs=dict(minmax_scale=False, hidden_sizes=[16, 16]))


def test_td_two_elements():
    base = bulk_fcc('Ni', 3.6, (2, 2, 2))
    rng = np.random.default_rng(5)
    sym = ['Mo' if x < 0.4 else 'Ni' for x in rng.random(len(base))]
    atoms = Atoms(sym, base.positions + rng.normal(scale=0.1, size=base.positions.shape),
                  base.cell, True)
    _compare(TemperatureDependentAtomicNN, atoms, ['Mo', 'Ni'], 4.6, 0.17, angular=False,
             nn_kwargs=dict(atomic_static_energy={'Mo': -1.5, 'Ni': -0.7},
                            hidden_sizes=[16]))
