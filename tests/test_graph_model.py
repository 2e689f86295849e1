"""Reading a frozen .pb exported by the reference (nn/basic.py:1017-1153): the
shipped legacy-layout file test_files/models/Mo.zhou04.pb."""
import json
import os

import numpy as np
import pytest

from tensoralloy_b200.io.graph_model import load_graph_model

GOLD = os.path.join(os.path.dirname(__file__), 'golden')
PB = os.path.join(GOLD, 'Mo.zhou04.pb')


def test_parse_reference_pb():
    m = load_graph_model(PB)
    assert m.precision == 'high'
    assert m.nn.__class__.__name__ == 'EamAlloyNN'
    assert m.nn.transformer.as_dict() == {
        'class': 'UniversalTransformer', 'elements': ['Mo'], 'rcut': 6.5, 'acut': None,
        'angular': False, 'periodic': True, 'symmetric': True,
        'use_computed_dists': True}
    assert m.nn.potentials == {'Mo': {'rho': 'zjw04', 'embed': 'zjw04'},
                               'MoMo': {'phi': 'zjw04'}}
    data = json.load(open(os.path.join(os.path.dirname(GOLD), '..', 'tensoralloy_b200',
                                       'data', 'zjw04.json')))['zjw04']['Mo']
    for key, val in data.items():
        assert m.nn.get_variable(f'EAM/Shared/Mo/{key}') == val
    assert 'energy' in m.predict_properties and 'forces' in m.predict_properties


@pytest.mark.gpu
def test_calculator_from_pb_matches_model_object():
    from tensoralloy_b200.atoms import Atoms
    from tensoralloy_b200.calculator import TensorAlloyCalculator
    from tensoralloy_b200.nn.eam import EamAlloyNN
    from tensoralloy_b200.precision import precision_scope
    from tensoralloy_b200.transformer import UniversalTransformer
    # bcc Mo 3x3x3, rattled
    a = 3.147
    base = np.array([[0, 0, 0], [.5, .5, .5]]) * a
    atoms = Atoms(['Mo'] * 2, base, np.eye(3) * a, True).repeat((3, 3, 3))
    atoms.positions += np.random.default_rng(0).normal(scale=0.05,
                                                       size=atoms.positions.shape)
    calc = TensorAlloyCalculator(PB)
    e = calc.get_potential_energy(atoms)
    f = calc.get_forces(atoms)
    s = calc.get_stress(atoms)
    assert e.dtype == np.float64        # 'high' precision read from the file
    with precision_scope('high'):
        nn = EamAlloyNN(['Mo'], custom_potentials='zjw04',
                        export_properties=['energy', 'forces', 'stress'])
        nn.attach_transformer(UniversalTransformer(['Mo'], rcut=6.5))
        ref = TensorAlloyCalculator(nn)
        assert abs(ref.get_potential_energy(atoms) - e) < 1e-12
        assert np.abs(ref.get_forces(atoms) - f).max() < 1e-12
        assert np.abs(ref.get_stress(atoms) - s).max() < 1e-14
    assert calc.get_model_timestamp().startswith('2020-07-27')
    atomic = calc.get_property('atomic', atoms)
    assert abs(atomic.sum() - e) < 1e-9
