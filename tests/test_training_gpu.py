"""Training step of AtomicNN (config 4): loss (energy/atom + forces + stress RMSE)
and its PARAMETER GRADIENTS from the GPU path (libtab200 descriptors + force op +
JVP kernel, torch MLP) vs the oracle's torch double-backward on the CPU."""
import numpy as np
import pytest
import torch

from oracle import training as otr
from tensoralloy_b200.atoms import Atoms, bulk_fcc
from tensoralloy_b200.nn.atomic import AtomicNN, SymmetryFunction
from tensoralloy_b200.nn.atomic.training import AtomicNNTrainer
from tensoralloy_b200.precision import precision_scope
from tensoralloy_b200.transformer import UniversalTransformer

pytestmark = pytest.mark.gpu


def make_structures(n_struct, seed=0):
    rng = np.random.default_rng(seed)
    out = []
    for k in range(n_struct):
        a = 3.3 + 0.4 * rng.random()
        base = bulk_fcc('Ni', a, (2, 2, 2))
        sym = ['Mo' if x < 0.4 else 'Ni' for x in rng.random(len(base))]
        pos = base.positions + rng.normal(scale=0.1, size=base.positions.shape)
        atoms = Atoms(sym, pos, base.cell, True)
        out.append(dict(atoms=atoms, symbols=sym, positions=pos, cell=base.cell,
                        pbc=[1, 1, 1], energy=-4.0 * len(base) + rng.normal(),
                        forces=rng.normal(scale=0.3, size=pos.shape),
                        stress=rng.normal(scale=0.01, size=6)))
    return out


def test_parameter_gradients_match_oracle():
    structs = make_structures(3)
    elements = ['Mo', 'Ni']
    rc = 4.5
    with precision_scope('high'):
        nn = AtomicNN(elements, SymmetryFunction(elements), hidden_sizes=[16, 16],
                      activation='softplus', use_resnet_dt=True, minmax_scale=False,
                      minimize_properties=('energy', 'forces', 'stress'),
                      export_properties=('energy', 'forces', 'stress'))
        nn.attach_transformer(UniversalTransformer(elements, rcut=rc, angular=True))
        nn.initialize_variables(seed=3)
        for el in elements:
            key = f"Atomic/{el}/Output/kernel"
            nn.set_variable(key, nn.get_variable(key) * 0.05)
        tr = AtomicNNTrainer(nn)
        for s in structs:
            tr.add_structure(s['atoms'], s['energy'], s['forces'], s['stress'])
        loss, parts = tr.gradients()
        params = {el: nn.mlp_params(el) for el in elements}
        ref_loss, ref_parts, ref_g = otr.loss_and_grads(elements, structs, params, rc)
        assert abs(loss.item() - ref_loss) < 1e-9 * max(1.0, abs(ref_loss))
        for key in ('energy', 'forces', 'stress'):
            assert abs(parts[key].item() - ref_parts[key]) < 1e-9
        for el in elements:
            L = tr.layers[el]
            for k, w in enumerate(L['W']):
                g = w.grad.cpu().numpy()
                r = ref_g[el][0][k]
                assert np.abs(g - r).max() < 1e-8 * max(1.0, np.abs(r).max()), (el, k)
            for k, b in enumerate(L['b']):
                if b is None or ref_g[el][1].get(k) is None:
                    continue
                g = b.grad.cpu().numpy()
                r = ref_g[el][1][k]
                assert np.abs(g - r).max() < 1e-8 * max(1.0, np.abs(r).max()), (el, k)
        # the same step captured in one CUDA graph gives the same gradients
        eager = [p.grad.clone() for p in tr.params]
        assert tr.enable_graph(), getattr(tr, 'graph_error', '')
        tr._graph[0].replay()
        for p, g in zip(tr.params, eager):
            assert torch.allclose(p.grad, g, rtol=1e-12, atol=1e-14)
        # a few Adam steps reduce the loss; parameters flow back into the model
        opt = torch.optim.Adam(tr.params, lr=1e-3)
        l0 = loss.item()
        for _ in range(20):
            l, _ = tr.train_step(opt)
        assert l.item() < l0
        tr.sync_to_model()
        assert np.allclose(nn.mlp_params('Ni')['weights'][0],
                           tr.layers['Ni']['W'][0].detach().cpu().numpy())


def test_grap_parameter_gradients_match_oracle():
    """The default descriptor of the reference (GenericRadialAtomicPotential, legacy mode)
    with multipole moments 0, 1, 2: force op + JVP kernel carry the moment terms."""
    from tensoralloy_b200.nn.atomic import GenericRadialAtomicPotential
    structs = make_structures(2, seed=5)
    elements = ['Mo', 'Ni']
    rc = 4.5
    with precision_scope('high'):
        desc = GenericRadialAtomicPotential(elements, algorithm='pexp',
                                            parameters=dict(rl=[1.5, 2.5], pl=[2.0, 3.0]),
                                            param_space_method='pair',
                                            moment_tensors=[0, 1, 2])
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
                    cutoff='cosine')
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


@pytest.mark.parametrize("special", [False, True])
def test_temperature_dependent_parameter_gradients_match_oracle(special):
    """TemperatureDependentAtomicNN / BeNN: U, F = U - T S and S losses + forces and stress
    of the free energy (finite_temperature.py:338-366, basic.py:190-202)."""
    from tensoralloy_b200.nn.atomic import BeNN, TemperatureDependentAtomicNN
    from tensoralloy_b200.nn.atomic.training import TemperatureDependentTrainer
    rng = np.random.default_rng(11)
    structs = []
    for k in range(2):
        base = bulk_fcc('Ni', 3.2 + 0.2 * k, (2, 2, 2))
        sym = ['Be'] * len(base)
        pos = base.positions + rng.normal(scale=0.1, size=base.positions.shape)
        atoms = Atoms(sym, pos, base.cell, True)
        atoms.info['etemperature'] = 0.1 + 0.2 * k
        structs.append(dict(atoms=atoms, symbols=sym, positions=pos, cell=base.cell,
                            pbc=[1, 1, 1], etemperature=atoms.info['etemperature'],
                            energy=-3.0 * len(base), free_energy=-3.1 * len(base),
                            eentropy=0.3 * len(base),
                            forces=rng.normal(scale=0.3, size=pos.shape),
                            stress=rng.normal(scale=0.01, size=6)))
    minimize = ('energy', 'free_energy', 'eentropy', 'forces', 'stress')
    with precision_scope('high'):
        cls = BeNN if special else TemperatureDependentAtomicNN
        nn = cls(['Be'], SymmetryFunction(['Be']), hidden_sizes=[12, 12], minmax_scale=False,
                 activation='softplus', minimize_properties=minimize,
                 export_properties=('energy', 'forces', 'stress'),
                 finite_temperature=dict(activation='tanh', layers=[16, 8],
                                         algo='default' if special else 'Sommerfeld'))
        nn.attach_transformer(UniversalTransformer(['Be'], rcut=4.5, angular=True))
        nn.initialize_variables(seed=2)
        for head in ('U', 'S'):
            key = f"TD/Be/{head}/Output/kernel"
            nn.set_variable(key, nn.get_variable(key) * 0.2)
        key = "TD/Be/H/Conv1d1/kernel"
        nn.set_variable(key, nn.get_variable(key) * 0.2)
        tr = TemperatureDependentTrainer(nn)
        for s in structs:
            tr.add_structure(s['atoms'], s['energy'], s['forces'], s['stress'],
                             free_energy=s['free_energy'], eentropy=s['eentropy'])
        loss, parts = tr.gradients()
        ref_loss, ref_parts, ref_g = otr.td_loss_and_grads(
            ['Be'], structs, {'Be': nn.td_params('Be')}, 4.5, minimize)
        assert abs(loss.item() - ref_loss) < 1e-9 * max(1.0, abs(ref_loss))
        for key, val in ref_parts.items():
            assert abs(parts[key].item() - val) < 1e-9 * max(1.0, abs(val)), key
        checked = 0
        for name, r in ref_g.items():
            assert name in tr.named, name
            if r is None:
                continue
            g = tr.named[name].grad.cpu().numpy().reshape(r.shape)
            assert np.abs(g - r).max() < 1e-8 * max(1.0, np.abs(r).max()), name
            checked += 1
        assert checked >= 10
        opt = torch.optim.Adam(tr.params, lr=1e-3)
        l0 = loss.item()
        for _ in range(10):
            l, _ = tr.train_step(opt)
        assert l.item() < l0
        tr.sync_to_model()
        assert np.allclose(nn.get_variable('TD/Be/U/Output/kernel').reshape(-1),
                           tr.named['TD/Be/U/Output/kernel'].detach().cpu().numpy().reshape(-1))


def test_training_step_in_medium_precision():
    """'medium' (float32, the reference's default precision): the same training step runs
    with float32 kernels and float32 torch leaves; loss within 1e-4 relative of float64."""
    structs = make_structures(2, seed=21)
    elements = ['Mo', 'Ni']
    out = {}
    for prec in ('high', 'medium'):
        with precision_scope(prec):
            nn = AtomicNN(elements, SymmetryFunction(elements), hidden_sizes=[16, 16],
                          minmax_scale=False,
                          minimize_properties=('energy', 'forces', 'stress'))
            nn.attach_transformer(UniversalTransformer(elements, rcut=4.5, angular=True))
            nn.initialize_variables(seed=3)
            for el in elements:
                key = f"Atomic/{el}/Output/kernel"
                nn.set_variable(key, nn.get_variable(key) * 0.05)
            tr = AtomicNNTrainer(nn)
            for s in structs:
                tr.add_structure(s['atoms'], s['energy'], s['forces'], s['stress'])
            loss, _ = tr.gradients()
            assert tr.params[0].dtype == (torch.float64 if prec == 'high' else torch.float32)
            assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in tr.params)
            out[prec] = loss.item()
    assert abs(out['medium'] - out['high']) < 1e-4 * abs(out['high'])
