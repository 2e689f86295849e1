"""Model-file layouts on the drop-in boundary: the LAMMPS-native `.npz`
of `AtomicNN.export_to_lammps_native` (tensoralloy/nn/atomic/atomic.py:304-480) and the
frozen-`.pb` parameter container of `BasicNN.export` (nn/basic.py:1017-1153).
CPU: keys / dtypes / shapes / round trips and the min-max fold against the oracle MLP.
GPU (marked): an exported and re-loaded model evaluates identically."""
import os

import numpy as np
import pytest
import torch

from oracle import atomic as oat
from tensoralloy_b200.atoms import Atoms
from tensoralloy_b200.io import native
from tensoralloy_b200.nn.atomic import AtomicNN, GenericRadialAtomicPotential
from tensoralloy_b200.precision import precision_scope
from tensoralloy_b200.transformer import UniversalTransformer

GOLD = os.path.join(os.path.dirname(__file__), 'golden')


def _grap_model(elements=('O', 'Pd'), minmax=False, algorithm='pexp', moments=(0, 1, 2),
                method='pair', seed=7, hidden=(16, 16, 8), resnet=True, static=True):
    params = {'pexp': dict(rl=[1.5, 2.5, 3.0], pl=[2.0, 3.0, 1.0]),
              'morse': dict(D=[0.5, 1.0], gamma=[1.2, 0.8], r0=[2.2, 2.6]),
              'density': dict(A=[1.0, 2.0], beta=[3.0, 5.0], re=[2.2, 2.4]),
              'sf': dict(eta=[0.05, 4.0], omega=[0.0, 1.0])}[algorithm]
    elements = sorted(elements)
    desc = GenericRadialAtomicPotential(elements, algorithm=algorithm, parameters=params,
                                        param_space_method=method,
                                        moment_tensors=list(moments),
                                        cutoff_function='polynomial')
    nn = AtomicNN(elements, desc, hidden_sizes=list(hidden), activation='softplus',
                  minmax_scale=minmax, use_resnet_dt=resnet,
                  use_atomic_static_energy=static,
                  atomic_static_energy={e: -1.0 - k for k, e in enumerate(elements)},
                  export_properties=('energy', 'forces', 'stress'))
    nn.attach_transformer(UniversalTransformer(elements, rcut=5.5, angular=False))
    nn.initialize_variables(seed=seed)
    rng = np.random.default_rng(seed)
    for e in elements:
        for k in range(1, len(hidden) + 1):
            key = f"Atomic/{e}/Conv1d{k}/bias"
            nn.set_variable(key, rng.normal(size=nn.get_variable(key).shape) * 0.1)
        if minmax:
            dim = desc.dimension(False)
            nn.set_variable(f"Atomic/{e}/xlo", rng.uniform(-0.5, 0.0, size=(1, 1, dim)))
            nn.set_variable(f"Atomic/{e}/xhi", rng.uniform(1.0, 3.0, size=(1, 1, dim)))
    return nn


def test_npz_keys_match_reference_layout(tmp_path):
    nn = _grap_model()
    path = str(tmp_path / 'model.npz')
    nn.export_to_lammps_native(path)
    z = np.load(path)
    # atomic.py:368-478, key for key
    expect = {"rmax", "nelt", "masses", "numbers", "tdnp", "precision", "use_fnn",
              "descriptor::method", "descriptor::rl", "descriptor::pl", "nlayers",
              "max_moment", "actfn", "fctype", "layer_sizes", "use_resnet_dt",
              "apply_output_bias", "is_T_symmetric"}
    for i in range(2):
        for j in range(4):
            expect |= {f"weights_{i}_{j}", f"biases_{i}_{j}"}
    assert set(z.files) == expect
    assert z["rmax"].dtype == np.float64 and float(z["rmax"]) == 5.5
    assert int(z["nelt"]) == 2 and int(z["precision"]) == 64 and int(z["tdnp"]) == 0
    # 'O' is padded with a zero code, 'Pd' takes two (atomic.py:359-366)
    assert z["numbers"].dtype == np.int32
    assert z["numbers"].tolist() == [ord('O'), 0, ord('P'), ord('d')]
    assert np.allclose(z["masses"], [15.999, 106.42])
    assert int(z["descriptor::method"]) == 0 and int(z["fctype"]) == 1
    assert int(z["actfn"]) == 1 and int(z["max_moment"]) == 2
    assert z["layer_sizes"].dtype == np.int32
    assert z["layer_sizes"].tolist() == [16, 16, 8, 1] and int(z["nlayers"]) == 4
    dim = nn.descriptor.dimension(False)
    assert z["weights_0_0"].shape == (dim, 16) and z["biases_0_0"].shape == (16,)
    assert z["weights_1_3"].shape == (8,) and z["biases_1_3"].shape == ()
    assert float(z["biases_0_3"]) == -1.0 and float(z["biases_1_3"]) == -2.0
    # float32 variant
    nn.export_to_lammps_native(path, dtype=np.float32)
    z = np.load(path)
    assert int(z["precision"]) == 32 and z["weights_0_1"].dtype == np.float32
    assert z["masses"].dtype == np.float32 and z["layer_sizes"].dtype == np.int32


@pytest.mark.parametrize("algorithm,method,code", [('pexp', 'pair', 0), ('morse', 'pair', 1),
                                                   ('density', 'cross', 2), ('sf', 'cross', 3)])
def test_npz_round_trip(tmp_path, algorithm, method, code):
    nn = _grap_model(algorithm=algorithm, method=method, moments=(0, 1, 2), resnet=False,
                     static=False, elements=('Be',))
    path = str(tmp_path / 'm.npz')
    nn.export_to_lammps_native(path)
    z = np.load(path)
    assert int(z["descriptor::method"]) == code
    # 'cross' grids are written as pairs (Algorithm.as_dict(convert_to_pairs=True),
    # grap.py:104-112): one entry per grid row for every key
    for key in native.METHOD_KEYS[algorithm]:
        assert z[f"descriptor::{key}"].shape == (len(nn.descriptor.grid),)
    assert "biases_0_3" not in z.files and int(z["apply_output_bias"]) == 0
    back, precision = native.read_lammps_native(path)
    assert precision == 'high'
    assert back.elements == ['Be'] and back.transformer.rcut == 5.5
    assert not back.transformer.angular
    assert back.descriptor.algorithm == algorithm
    assert back.descriptor.radial_sets() == nn.descriptor.radial_sets()
    # only max_moment is stored: the file means moments 0..max
    assert list(back.descriptor.moment_tensors) == [0, 1, 2]
    assert back.descriptor.cutoff_function == 'polynomial'
    assert back.hidden_sizes == nn.hidden_sizes and back.activation == 'softplus'
    assert not back.use_resnet_dt and not back.use_atomic_static_energy


def test_npz_second_export_is_identical(tmp_path):
    nn = _grap_model()
    a, b = str(tmp_path / 'a.npz'), str(tmp_path / 'b.npz')
    nn.export_to_lammps_native(a)
    back, _ = native.read_lammps_native(a)
    back.export_to_lammps_native(b)
    za, zb = np.load(a), np.load(b)
    assert set(za.files) == set(zb.files)
    for k in za.files:
        assert za[k].dtype == zb[k].dtype and za[k].shape == zb[k].shape
        assert np.array_equal(za[k], zb[k]), k


def test_npz_folds_minmax_scaling_exactly(tmp_path):
    """The reference drops xlo / xhi in this file; here the affine map is folded into the
    first layer.  Oracle MLP on raw descriptors with the scaling == oracle MLP with the
    file's layers."""
    nn = _grap_model(minmax=True)
    hi = nn.get_variable("Atomic/O/xhi").copy()
    hi[..., 2] = nn.get_variable("Atomic/O/xlo")[..., 2]      # a zero range: x' = 0 there
    nn.set_variable("Atomic/O/xhi", hi)
    path = str(tmp_path / 'm.npz')
    nn.export_to_lammps_native(path)
    back, _ = native.read_lammps_native(path)
    rng = np.random.default_rng(3)
    for e in nn.elements:
        p, q = nn.mlp_params(e), back.mlp_params(e)
        assert q['xlo'] is None
        x = torch.tensor(rng.uniform(0.0, 2.0, size=(9, len(p['xlo']))))
        # oracle/atomic.py:280-282 == atomic.py:199 (div_no_nan)
        den = torch.tensor(p['xhi'] - p['xlo'])
        xs = torch.where(den == 0, torch.zeros_like(x), (torch.tensor(p['xhi']) - x) / den)
        y0 = oat.mlp(xs, [torch.tensor(w) for w in p['weights']],
                     [None if b is None else torch.tensor(b) for b in p['biases']],
                     p['activation'], p['use_resnet_dt'], torch.tensor(p['out_bias']))
        y1 = oat.mlp(x, [torch.tensor(w) for w in q['weights']],
                     [None if b is None else torch.tensor(b) for b in q['biases']],
                     q['activation'], q['use_resnet_dt'], torch.tensor(q['out_bias']))
        assert float((y0 - y1).abs().max()) < 1e-12


def test_npz_error_behaviour(tmp_path):
    from tensoralloy_b200.nn.atomic import SymmetryFunction
    nn = AtomicNN(['Be'], SymmetryFunction(['Be']))
    nn.attach_transformer(UniversalTransformer(['Be'], rcut=5.0, angular=False))
    with pytest.raises(ValueError, match="GenericRadialAtomicPotential is required"):
        nn.export_to_lammps_native(str(tmp_path / 'x.npz'))         # atomic.py:312-314
    desc = GenericRadialAtomicPotential(['Al', 'Be'], 'sf', dict(eta=[1.0], omega=[0.0]))
    nn = AtomicNN(['Al', 'Be'], desc, hidden_sizes={'Al': [8, 8], 'Be': [8, 4]},
                  minmax_scale=False)
    nn.attach_transformer(UniversalTransformer(['Al', 'Be'], rcut=5.0, angular=False))
    nn.initialize_variables()
    with pytest.raises(ValueError, match="Layer sizes of all elements"):   # atomic.py:318-320
        nn.export_to_lammps_native(str(tmp_path / 'x.npz'))
    good = _grap_model(elements=('Be',), resnet=False)
    data = native.lammps_native_dict(good)
    data["use_fnn"] = np.int32(1)                      # claims a filter network, has none
    np.savez(str(tmp_path / 'fnn.npz'), **data)
    with pytest.raises(KeyError):
        native.read_lammps_native(str(tmp_path / 'fnn.npz'))
    with pytest.raises(ValueError, match="max_moment"):
        _grap_model(elements=('Be',), moments=(0, 2)).export_to_lammps_native(
            str(tmp_path / 'z.npz'))


@pytest.mark.gpu
def test_npz_model_evaluates_like_the_source_model(tmp_path):
    from tensoralloy_b200.calculator import TensorAlloyCalculator
    from tensoralloy_b200.precision import precision_scope
    d = np.load(os.path.join(GOLD, 'Be_liquid_4000K.npz'))
    be = Atoms(list(d['symbols']), d['positions'][1], d['cells'][1], True)
    nn = _grap_model(elements=('Be',), minmax=True, algorithm='morse', moments=(0, 1, 2))
    path = str(tmp_path / 'be.npz')
    nn.export_to_lammps_native(path)
    with precision_scope('high'):
        a = TensorAlloyCalculator(nn)
        a.calculate(be, properties=['energy', 'forces', 'stress'])
    b = TensorAlloyCalculator(path)
    assert b.elements == ['Be'] and 'stress' in b.predict_properties
    b.calculate(be, properties=['energy', 'forces', 'stress'])
    assert abs(a.results['energy'] - b.results['energy']) / len(be) < 1e-10
    assert np.abs(a.results['forces'] - b.results['forces']).max() < 1e-8
    assert np.abs(a.results['stress'] - b.results['stress']).max() < 1e-8
    assert np.abs(a.results['forces']).max() > 1e-3


# --- frozen .pb parameter container: BasicNN.export (basic.py:1017-1153) ----------------
def _same_variables(a, b):
    assert set(a) == set(b)
    for k in a:
        assert np.array_equal(np.asarray(a[k]).reshape(-1), np.asarray(b[k]).reshape(-1)), k


def test_pb_export_round_trip_atomic(tmp_path):
    from tensoralloy_b200.io.graph_model import load_graph_model, parse_graph_def, const_value
    from tensoralloy_b200.nn.atomic import SymmetryFunction
    from tensoralloy_b200.utils import ModeKeys
    els = ['O', 'Pd']
    nn = AtomicNN(els, SymmetryFunction(els, eta=[0.1, 2.0], omega=[0.0, 1.5],
                                        beta=[0.005], gamma=[1.0, -1.0], zeta=[1.0, 4.0],
                                        cutoff_function='polynomial'),
                  hidden_sizes={'O': [12, 6], 'Pd': [10]}, activation='tanh',
                  atomic_static_energy={'O': -3.0, 'Pd': -5.0},
                  export_properties=('energy', 'forces', 'stress', 'hessian'))
    nn.attach_transformer(UniversalTransformer(els, rcut=6.0, acut=5.0, angular=True))
    nn.initialize_variables(seed=3)
    path = str(tmp_path / 'atomic.pb')
    with precision_scope('high'):       # the file carries the precision in force
        nn.export(path)
    g = parse_graph_def(path)
    consts = {n.name: n for n in g.node}
    # the reference's metadata nodes (basic.py:1075-1092)
    for key in ('Transformer/params', 'Metadata/timestamp', 'Metadata/precision',
                'Metadata/variational_energy', 'Metadata/is_finite_temperature',
                'Metadata/api', 'Metadata/ops', 'Atomic/O/Conv1d2/kernel',
                'Atomic/Pd/Output/bias', 'Atomic/Pd/xlo'):
        assert key in consts, key
    assert const_value(consts['Atomic/O/Conv1d1/kernel']).dtype == np.float64
    assert const_value(consts['Metadata/variational_energy']) == b'energy'
    loaded = load_graph_model(path)
    back = loaded.nn
    assert loaded.precision == 'high' and loaded.api_version == '1.1'
    assert set(loaded.predict_properties) == {'energy', 'forces', 'stress', 'hessian'}
    assert loaded.ops['forces'] == 'Output/Forces/forces:0'
    assert back.as_dict() == nn.as_dict()
    assert back.transformer.as_dict() == nn.transformer.as_dict()
    _same_variables(nn.variables, back.variables)
    # 'medium' files hold float32 constants and load as 'medium'
    nn.export(path, precision='medium')
    loaded = load_graph_model(path)
    assert loaded.precision == 'medium'
    k = 'Atomic/O/Conv1d1/kernel'
    assert np.array_equal(loaded.nn.variables[k], nn.variables[k].astype(np.float32))
    with pytest.raises(ValueError, match="transformer must be attached"):
        AtomicNN(els).export(path)
    # mode NATIVE -> the npz layout
    grap = _grap_model(elements=('Be',))
    grap.export(str(tmp_path / 'be.npz'), mode=ModeKeys.NATIVE)
    assert int(np.load(str(tmp_path / 'be.npz'))['nelt']) == 1


def test_pb_export_round_trip_eam(tmp_path):
    from tensoralloy_b200.io.graph_model import load_graph_model, parse_graph_def
    from tensoralloy_b200.nn.eam import AdpNN, EamAlloyNN
    nn = EamAlloyNN(['Al', 'Cu'], custom_potentials={
        'Al': {'rho': 'zjw04', 'embed': 'nn'}, 'Cu': {'rho': 'zjw04', 'embed': 'zjw04'},
        'AlAl': {'phi': 'zjw04'}, 'AlCu': {'phi': 'nn'}, 'CuCu': {'phi': 'zjw04'}},
        hidden_sizes=[8, 4])
    nn.attach_transformer(UniversalTransformer(['Al', 'Cu'], rcut=6.0))
    nn.initialize_variables(seed=11)
    nn.set_variable('EAM/Shared/Al/r_eq', 2.9)          # a "trained" shared variable
    path = str(tmp_path / 'eam.pb')
    nn.export(path, precision='high')
    names = {n.name for n in parse_graph_def(path).node}
    assert 'EAM/Shared/Al/r_eq' in names and 'EAM/Shared/Cu/F0' in names
    assert 'EAM/Shared/Ni/F0' not in names              # only the model's own elements
    back = load_graph_model(path).nn
    assert isinstance(back, EamAlloyNN)
    assert back.as_dict() == nn.as_dict()
    assert back.get_variable('EAM/Shared/Al/r_eq') == 2.9
    assert back.get_variable('EAM/Shared/Cu/r_eq') == nn.get_variable('EAM/Shared/Cu/r_eq')
    assert len(nn.variables) > 0
    _same_variables(nn.variables, back.variables)
    adp = AdpNN(['Mo', 'Ni'])
    adp.attach_transformer(UniversalTransformer(['Mo', 'Ni'], rcut=6.5))
    adp.export(path)
    back = load_graph_model(path).nn
    assert isinstance(back, AdpNN) and back.as_dict() == adp.as_dict()


def test_pb_export_round_trip_temperature_dependent(tmp_path):
    from tensoralloy_b200.io.graph_model import load_graph_model
    from tensoralloy_b200.nn.atomic import SymmetryFunction
    from tensoralloy_b200.nn.atomic.finite_temperature import TemperatureDependentAtomicNN
    nn = TemperatureDependentAtomicNN(
        ['Be'], SymmetryFunction(['Be']), hidden_sizes=[8, 8],
        export_properties=('energy', 'forces', 'stress'))
    nn.attach_transformer(UniversalTransformer(['Be'], rcut=5.0, angular=False))
    nn.initialize_variables(seed=5)
    path = str(tmp_path / 'td.pb')
    nn.export(path, precision='high')
    loaded = load_graph_model(path)
    assert type(loaded.nn) is TemperatureDependentAtomicNN
    assert loaded.nn.is_finite_temperature and loaded.nn.variational_energy == 'free_energy'
    assert 'free_energy' in loaded.ops and 'eentropy' in loaded.ops
    assert loaded.nn.as_dict() == nn.as_dict()
    _same_variables(nn.variables, loaded.nn.variables)


@pytest.mark.gpu
def test_pb_export_evaluates_like_the_source_model(tmp_path):
    from tensoralloy_b200.atoms import bulk_fcc
    from tensoralloy_b200.calculator import TensorAlloyCalculator
    from tensoralloy_b200.nn.eam import EamAlloyNN
    from tensoralloy_b200.precision import precision_scope
    atoms = bulk_fcc('Ni', 3.52, (3, 3, 3))
    rng = np.random.default_rng(2)
    atoms.positions += rng.normal(scale=0.05, size=atoms.positions.shape)
    nn = EamAlloyNN(['Ni'], custom_potentials={'Ni': {'rho': 'zjw04', 'embed': 'nn'},
                                               'NiNi': {'phi': 'zjw04'}}, hidden_sizes=[8, 8])
    nn.attach_transformer(UniversalTransformer(['Ni'], rcut=6.0))
    nn.initialize_variables(seed=4)
    nn.set_variable('EAM/Shared/Ni/r_eq', 2.5)
    path = str(tmp_path / 'ni.pb')
    with precision_scope('high'):
        nn.export(path)
        a = TensorAlloyCalculator(nn)
        a.calculate(atoms, properties=['energy', 'forces', 'stress'])
    b = TensorAlloyCalculator(path)
    b.calculate(atoms, properties=['energy', 'forces', 'stress'])
    # float64 constants: the re-loaded model holds the same numbers
    assert abs(a.results['energy'] - b.results['energy']) < 1e-12 * len(atoms)
    assert np.abs(a.results['forces'] - b.results['forces']).max() < 1e-12
    assert np.abs(a.results['stress'] - b.results['stress']).max() < 1e-12
    assert np.abs(a.results['forces']).max() > 1e-3


def test_npz_round_trip_keeps_the_new_mode_details(tmp_path):
    """`is_T_symmetric` and `max_moment` = 3 exist in the reference's new mode only
    (grap.py:434-457, 485-494): a file carrying either reads back as a new-mode model;
    plain files with moments <= 2 stay legacy (the two agree there)."""
    from tensoralloy_b200.io import native
    from tensoralloy_b200.nn.atomic import GenericRadialAtomicPotential as Grap
    par = dict(rl=[1.5, 2.5], pl=[2.0, 3.0])

    def model(**kw):
        desc = Grap(['Be'], 'pexp', par, **kw)
        nn = AtomicNN(['Be'], desc, hidden_sizes=[8, 8], minmax_scale=False,
                      export_properties=('energy', 'forces', 'stress'))
        nn.attach_transformer(UniversalTransformer(['Be'], rcut=5.0, angular=False))
        nn.initialize_variables(seed=1)
        return nn

    for kw, flags, moments in ((dict(moment_tensors=3, symmetric=True, legacy_mode=False), 3,
                                (0, 1, 2, 3)),
                               (dict(moment_tensors=2, symmetric=True, legacy_mode=False), 3,
                                (0, 1, 2)),
                               (dict(moment_tensors=3, legacy_mode=False), 1, (0, 1, 2, 3)),
                               (dict(moment_tensors=[0, 1, 2]), 0, (0, 1, 2))):
        src = model(**kw)
        path = str(tmp_path / f"m{flags}_{len(moments)}.npz")
        src.export_to_lammps_native(path)
        z = np.load(path)
        assert int(z["max_moment"]) == moments[-1]
        assert int(z["is_T_symmetric"]) == int(bool(kw.get('symmetric', False)))
        back, _ = native.read_lammps_native(path)
        d = back.descriptor
        assert d.grap_flags() == flags and d.moments() == moments
        assert d.dimension() == src.descriptor.dimension()
        for key in ("Atomic/Be/Conv1d1/kernel", "Atomic/Be/Output/kernel"):
            assert np.array_equal(back.get_variable(key), src.get_variable(key))


def test_npz_round_trip_of_a_filter_network_model(tmp_path):
    """atomic.py:408-438: `use_fnn = 1` and the `fnn::*` keys of the GRAP `nn` algorithm."""
    from tensoralloy_b200.io import native
    from tensoralloy_b200.nn.atomic import GenericRadialAtomicPotential as Grap
    from tensoralloy_b200.nn.atomic.grap_nn import filter_params
    desc = Grap(['Be', 'W'], 'nn', dict(num_filters=5, hidden_sizes=[8, 8],
                                        activation='tanh', use_resnet_dt=True),
                moment_tensors=2, symmetric=True, legacy_mode=False,
                cutoff_function='polynomial')
    nn = AtomicNN(['Be', 'W'], desc, hidden_sizes=[8, 8], activation='softplus',
                  minmax_scale=False, export_properties=('energy', 'forces', 'stress'))
    nn.attach_transformer(UniversalTransformer(['Be', 'W'], rcut=5.0, angular=False))
    nn.initialize_variables(seed=2)
    nn.set_variable("Atomic/Filters/Conv3d2/bias", np.linspace(-0.2, 0.3, 8))
    path = str(tmp_path / 'fnn.npz')
    nn.export_to_lammps_native(path)
    z = np.load(path)
    assert int(z["use_fnn"]) == 1 and "descriptor::method" not in z.files
    assert int(z["fnn::nlayers"]) == 3 and z["fnn::layer_sizes"].tolist() == [8, 8, 5]
    assert int(z["fnn::num_filters"]) == 5 and int(z["fnn::actfn"]) == 2       # tanh
    assert int(z["fnn::use_resnet_dt"]) == 1 and int(z["fnn::apply_output_bias"]) == 0
    assert int(z["fnn::h_abck_modifier"]) == 0
    assert z["fnn::weights_0_0"].shape == (8,) and z["fnn::weights_0_1"].shape == (8, 8)
    assert z["fnn::weights_0_2"].shape == (8, 5) and "fnn::biases_0_2" not in z.files
    assert int(z["max_moment"]) == 2 and int(z["is_T_symmetric"]) == 1
    back, _ = native.read_lammps_native(path)
    d = back.descriptor
    assert d.algorithm == 'nn' and d.grap_flags() == 3 and d.moments() == (0, 1, 2)
    assert d.as_dict()["parameters"] == desc.as_dict()["parameters"]
    a, b = filter_params(nn), filter_params(back)
    for x, y in zip(a['weights'], b['weights']):
        assert np.array_equal(x, y)
    for x, y in zip(a['biases'][:-1], b['biases'][:-1]):
        assert np.array_equal(x, y)
    for key in ("Atomic/W/Conv1d1/kernel", "Atomic/Be/Output/kernel"):
        assert np.array_equal(back.get_variable(key), nn.get_variable(key))


def test_filter_network_initialised_from_an_npz_checkpoint(tmp_path):
    """grap.py:248-262, convolutional.py:219-254: `ckpt` = a `use_fnn` npz file gives the
    architecture and the initial filter weights."""
    from tensoralloy_b200.nn.atomic import GenericRadialAtomicPotential as Grap
    from tensoralloy_b200.nn.atomic.grap_nn import filter_params

    def model(parameters, seed):
        desc = Grap(['Be'], 'nn', parameters, moment_tensors=1, legacy_mode=False)
        nn = AtomicNN(['Be'], desc, hidden_sizes=[8], minmax_scale=False)
        nn.attach_transformer(UniversalTransformer(['Be'], rcut=5.0, angular=False))
        nn.initialize_variables(seed=seed)
        return nn

    src = model(dict(num_filters=3, hidden_sizes=[6, 6], activation='tanh',
                     use_resnet_dt=False), seed=4)
    src.set_variable("Atomic/Filters/Conv3d1/bias", np.linspace(0.1, 0.6, 6))
    path = str(tmp_path / 'filters.npz')
    src.export_to_lammps_native(path)
    new = model(dict(ckpt=path, num_filters=99, hidden_sizes=[1]), seed=77)
    a = new.descriptor.algorithm_object
    assert (a.hidden_sizes, a.num_filters, a.activation, a.use_resnet_dt) == \
        ([6, 6], 3, 'tanh', False)
    assert new.descriptor.dimension() == 1 * 3 * 2
    p, q = filter_params(src), filter_params(new)
    for x, y in zip(p['weights'], q['weights']):
        assert np.array_equal(x, y)
    assert np.array_equal(p['biases'][0], q['biases'][0])
    plain = str(tmp_path / 'plain.npz')
    _grap_model(elements=('Be',)).export_to_lammps_native(plain)
    with pytest.raises(KeyError):                       # no fnn:: keys in a closed-form file
        model(dict(ckpt=plain), seed=1)
