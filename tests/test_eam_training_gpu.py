"""Training step of the EAM / ADP family (config 4 ii): loss (energy/atom + forces +
stress RMSE) and its PARAMETER GRADIENTS from the GPU path (libtab200 batch lists, pair
export, force op + its transpose kernel; torch for the scalar functions) vs the oracle's
torch double-backward on the CPU.  Models: EamAlloyNN with 'nn' and zjw04 functions
(shared empirical variables), AdpNN with 'nn' phi / rho / u / w."""
import numpy as np
import pytest
import torch

from oracle import atomic as oat
from oracle import potentials as opot
from oracle import training as otr
from tensoralloy_b200.atoms import Atoms, bulk_fcc
from tensoralloy_b200.nn.eam import AdpNN, EamAlloyNN
from tensoralloy_b200.nn.eam.training import EamTrainer
from tensoralloy_b200.precision import precision_scope
from tensoralloy_b200.transformer import UniversalTransformer

pytestmark = pytest.mark.gpu
ELEMENTS = ['Mo', 'Ni']
RC = 5.0


def make_structures(n_struct, seed=0):
    rng = np.random.default_rng(seed)
    out = []
    for k in range(n_struct):
        a = 3.45 + 0.3 * rng.random()
        base = bulk_fcc('Ni', a, (2, 2, 2) if k % 2 else (3, 3, 3))
        sym = ['Mo' if x < 0.4 else 'Ni' for x in rng.random(len(base))]
        pos = base.positions + rng.normal(scale=0.08, size=base.positions.shape)
        atoms = Atoms(sym, pos, base.cell, True)
        out.append(dict(atoms=atoms, symbols=sym, positions=pos, cell=base.cell,
                        pbc=[1, 1, 1], energy=-4.0 * len(base) + rng.normal(),
                        forces=rng.normal(scale=0.3, size=pos.shape),
                        stress=rng.normal(scale=0.01, size=6)))
    return out


def _oracle_fns(nn):
    """Oracle callables closing over LEAF tensors that carry the model's current values;
    returns (fns, leaves_fn) with leaves keyed by the reference variable names."""
    provider = nn._nn
    pots, shared = {}, {}
    nn_leaves = {}

    def empirical(name):
        if name not in pots:
            pots[name] = opot.get_potential(name)
            pots[name].track_parameters()
        return pots[name]

    def mlp(fn, section):
        arrays = provider.weights(fn, section)
        pre = provider.prefix(fn, section)
        rank = 1 if fn == 'embed' else 2
        nh = (len(arrays) - 1) // 2
        W, b = [], []
        for k in range(nh):
            for what, arr, dst in (('kernel', arrays[2 * k], W), ('bias', arrays[2 * k + 1], b)):
                name = f"{pre}/Conv{rank}d{k + 1}/{what}"
                if name not in nn_leaves:
                    nn_leaves[name] = torch.tensor(arr, dtype=torch.float64, requires_grad=True)
                dst.append(nn_leaves[name])
        name = f"{pre}/Output/kernel"
        if name not in nn_leaves:
            nn_leaves[name] = torch.tensor(arrays[-1], dtype=torch.float64, requires_grad=True)
        W.append(nn_leaves[name])
        return lambda x: oat.mlp(x[:, None], W, b, nn._activation)

    def dispatch(fn):
        def call(x, key):
            if nn.potentials[key][fn] == 'nn':
                return mlp(fn, key)(x)
            return getattr(empirical(nn.potentials[key][fn]), fn)(x, key)
        return call

    fns = {fn: dispatch(fn) for fn in ('rho', 'phi', 'embed', 'dipole', 'quadrupole')}

    def leaves_fn():
        out = dict(nn_leaves)
        for pot in pots.values():
            for (section, key), t in pot.leaves.items():
                out.setdefault(f"{nn.scope}/Shared/{section}/{key}", t)
        return out
    return fns, leaves_fn


def _check(nn, kind, structs, n_steps=10):
    # every shared variable as a leaf: the gradient check covers the parameters the
    # reference keeps frozen (zjw04 embedding parameters, r_eq) too
    tr = EamTrainer(nn, freeze_reference_fixed=False)
    for s in structs:
        tr.add_structure(s['atoms'], s['energy'], s['forces'], s['stress'])
    loss, parts = tr.gradients()
    fns, leaves_fn = _oracle_fns(nn)
    ref_loss, ref_parts, ref_g = otr.eam_loss_and_grads(kind, ELEMENTS, structs, fns,
                                                        leaves_fn, RC)
    print(kind, 'loss', ref_loss, ref_parts, 'd', abs(loss.item() - ref_loss))
    assert abs(loss.item() - ref_loss) < 1e-9 * max(1.0, abs(ref_loss))
    for key in ('energy', 'forces', 'stress'):
        assert abs(parts[key].item() - ref_parts[key]) < 1e-9 * max(1.0, abs(ref_parts[key]))
    named = tr.named_parameters()
    # every trainable variable is known to the optimiser before the first step
    assert {id(t) for t in named.values() if t.requires_grad} == {id(t) for t in tr.params}
    checked = 0
    for name, r in ref_g.items():
        assert name in named, name
        if r is None:
            continue
        g = named[name].grad
        assert g is not None, name
        g = g.cpu().numpy()
        scale = max(1.0, np.abs(r).max())
        assert np.abs(g - r).max() < 1e-7 * scale, (name, np.abs(g - r).max(), scale)
        checked += 1
    assert checked >= 6
    # the same step captured in one CUDA graph: identical gradients, then training
    eager = {name: t.grad.clone() for name, t in named.items() if t.grad is not None}
    assert tr.enable_graph(), getattr(tr, 'graph_error', '')
    opt = torch.optim.Adam(tr.params, lr=1e-3)
    tr._graph[0].replay()
    for name, g in eager.items():
        assert torch.allclose(named[name].grad, g, rtol=1e-12, atol=1e-14), name
    l0 = loss.item()
    for _ in range(n_steps):
        l, _ = tr.train_step(opt)
    assert l.item() < l0
    tr.sync_to_model()
    return tr


def test_eam_alloy_nn_and_zjw04_parameter_gradients():
    structs = make_structures(3)
    with precision_scope('high'):
        nn = EamAlloyNN(ELEMENTS, custom_potentials={
            'Ni': {'rho': 'zjw04', 'embed': 'nn'}, 'Mo': {'rho': 'nn', 'embed': 'zjw04'},
            'MoNi': {'phi': 'nn'}, 'NiNi': {'phi': 'zjw04'}, 'MoMo': {'phi': 'zjw04'}},
            hidden_sizes={'Ni': {'embed': [12]}, 'Mo': {'rho': [8, 6]},
                          'MoNi': {'phi': [16, 8]}},
            minimize_properties=('energy', 'forces', 'stress'))
        nn.attach_transformer(UniversalTransformer(ELEMENTS, rcut=RC))
        nn.initialize_variables(seed=3)
        for name, value in list(nn.variables.items()):
            if name.endswith('Output/kernel'):
                nn.set_variable(name, value * 0.2)
        tr = _check(nn, 'alloy', structs)
        # trained empirical values flow back into the model's shared variables
        name = 'EAM/Shared/Ni/A'
        assert abs(nn.get_variable(name) - tr.named_parameters()[name].item()) < 1e-15


def test_adp_all_nn_parameter_gradients():
    structs = make_structures(2, seed=4)
    with precision_scope('high'):
        nn = AdpNN(ELEMENTS, hidden_sizes=[8, 8],
                   minimize_properties=('energy', 'forces', 'stress'))
        nn.attach_transformer(UniversalTransformer(ELEMENTS, rcut=RC))
        nn.initialize_variables(seed=5)
        for name, value in list(nn.variables.items()):
            if name.endswith('Output/kernel'):
                scale = 0.02 if ('Dipole' in name or 'Quadrupole' in name) else 0.2
                nn.set_variable(name, value * scale)
        _check(nn, 'adp', structs)


def test_adp_mishinh_and_nn_parameter_gradients():
    """Trainable MishinH embedding / dipole / quadrupole functions (mishin.py:20-315) next to
    zjw04 and 'nn' functions in one ADP model."""
    structs = make_structures(2, seed=12)
    with precision_scope('high'):
        nn = AdpNN(ELEMENTS, custom_potentials={
            'Mo': {'rho': 'zjw04', 'embed': 'nn'}, 'Ni': {'rho': 'nn', 'embed': 'zjw04'},
            'MoMo': {'phi': 'zjw04', 'dipole': 'mishinh', 'quadrupole': 'nn'},
            'MoNi': {'phi': 'nn', 'dipole': 'mishinh', 'quadrupole': 'mishinh'},
            'NiNi': {'phi': 'zjw04', 'dipole': 'nn', 'quadrupole': 'mishinh'}},
            hidden_sizes=[8, 6], minimize_properties=('energy', 'forces', 'stress'))
        nn.attach_transformer(UniversalTransformer(ELEMENTS, rcut=RC))
        nn.initialize_variables(seed=6)
        for name, value in list(nn.variables.items()):
            if name.endswith('Output/kernel'):
                scale = 0.02 if ('Dipole' in name or 'Quadrupole' in name) else 0.2
                nn.set_variable(name, value * scale)
        tr = _check(nn, 'adp', structs)
        assert tr.named_parameters()['ADP/Shared/MoNi/q1'].grad is not None
        # by default the reference's never-trained parameters stay out of the optimiser
        frozen = EamTrainer(nn).named_parameters()
        assert not frozen['ADP/Shared/Ni/F0'].requires_grad
        assert frozen['ADP/Shared/MoNi/d1'].requires_grad


def test_eam_fs_all_nn_parameter_gradients():
    """Finnis-Sinclair: rho is a function of the ORDERED pair (fs.py:180-203)."""
    from tensoralloy_b200.nn.eam import EamFsNN
    structs = make_structures(2, seed=8)
    with precision_scope('high'):
        nn = EamFsNN(ELEMENTS, hidden_sizes=[8, 6],
                     minimize_properties=('energy', 'forces', 'stress'))
        nn.attach_transformer(UniversalTransformer(ELEMENTS, rcut=RC))
        nn.initialize_variables(seed=9)
        for name, value in list(nn.variables.items()):
            if name.endswith('Output/kernel'):
                nn.set_variable(name, value * 0.2)
        _check(nn, 'fs', structs)


def test_eam_training_step_in_medium_precision():
    structs = make_structures(2, seed=31)
    out = {}
    for prec in ('high', 'medium'):
        with precision_scope(prec):
            nn = EamAlloyNN(ELEMENTS, hidden_sizes=[8, 8],
                            minimize_properties=('energy', 'forces', 'stress'))
            nn.attach_transformer(UniversalTransformer(ELEMENTS, rcut=RC))
            nn.initialize_variables(seed=3)
            for name, value in list(nn.variables.items()):
                if name.endswith('Output/kernel'):
                    nn.set_variable(name, value * 0.2)
            tr = EamTrainer(nn)
            for s in structs:
                tr.add_structure(s['atoms'], s['energy'], s['forces'], s['stress'])
            loss, _ = tr.gradients()
            assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in tr.params)
            out[prec] = loss.item()
    assert abs(out['medium'] - out['high']) < 1e-4 * abs(out['high'])
