"""GRAP with the trainable `nn` algorithm (filter network; nn/atomic/grap.py:211-269,
619-646): the torch side of `GrapFilterTrainer` against the oracle on the CPU.  The
neighbour lists come from the oracle's ASE restatement and the library's pair-force op
(`tab_pair_forces` / `tab_pair_jvp`, include/tab200.h) is replaced by its definition in
torch -- F_i = sum_{p in row i} g_p - sum_{p -> i} g_p, W = sum_p sym(g_p (x) D_p) -- so
that the descriptor algebra, both networks and the double backward are checked without a
GPU; tests/test_zzz_grap_filter_gpu.py runs the same comparison through the library."""
import numpy as np
import pytest
import torch

from oracle import neighbor as onl
from oracle import training as otr
from tensoralloy_b200.atoms import Atoms, bulk_fcc
from tensoralloy_b200.nn.atomic import AtomicNN, GenericRadialAtomicPotential
from tensoralloy_b200.nn.atomic.grap_nn import GrapFilterTrainer, filter_params
from tensoralloy_b200.precision import precision_scope
from tensoralloy_b200.transformer import UniversalTransformer


def make_structures(n_struct=2, seed=5):
    rng = np.random.default_rng(seed)
    out = []
    for k in range(n_struct):
        base = bulk_fcc('Ni', 3.3 + 0.4 * rng.random(), (2, 2, 2))
        sym = ['Mo' if x < 0.4 else 'Ni' for x in rng.random(len(base))]
        pos = base.positions + rng.normal(scale=0.1, size=base.positions.shape)
        out.append(dict(atoms=Atoms(sym, pos, base.cell, True), symbols=sym, positions=pos,
                        cell=np.asarray(base.cell), pbc=[1, 1, 1],
                        energy=-4.0 * len(base) + rng.normal(),
                        forces=rng.normal(scale=0.3, size=pos.shape),
                        stress=rng.normal(scale=0.01, size=6)))
    return out


def make_model(elements, rc, max_moment, symmetric, cutoff='cosine', algorithm='nn',
               h_abck_modifier=0):
    parameters = dict(num_filters=5, hidden_sizes=[8, 8], activation='softplus',
                      use_resnet_dt=True, h_abck_modifier=h_abck_modifier) \
        if algorithm == 'nn' else \
        dict(rl=[1.5, 2.5, 2.0], pl=[2.0, 3.0, 1.0])
    desc = GenericRadialAtomicPotential(
        elements, algorithm, parameters,
        moment_tensors=max_moment, cutoff_function=cutoff, symmetric=symmetric,
        legacy_mode=False)
    nn = AtomicNN(elements, desc, hidden_sizes=[16, 16], activation='softplus',
                  minmax_scale=False, minimize_properties=('energy', 'forces', 'stress'),
                  export_properties=('energy', 'forces', 'stress'))
    nn.attach_transformer(UniversalTransformer(elements, rcut=rc, angular=False))
    nn.initialize_variables(seed=3)
    rng = np.random.default_rng(9)
    for k in ((1, 2) if algorithm == 'nn' else ()):      # non-zero filter biases
        key = f"Atomic/Filters/Conv3d{k}/bias"
        nn.set_variable(key, rng.normal(size=nn.get_variable(key).shape) * 0.3)
    for el in elements:
        key = f"Atomic/{el}/Output/kernel"
        nn.set_variable(key, nn.get_variable(key) * 0.05)
    return nn


def torch_pair_force(i, j, sid_of_pair, n, nb, D):
    def apply(g, _nbr):
        F = torch.zeros(n, 3, dtype=g.dtype).index_add(0, i, g).index_add(0, j, -g)
        outer = g[:, :, None] * D[:, None, :]
        sym = 0.5 * (outer + outer.transpose(1, 2))
        W = torch.zeros(nb, 3, 3, dtype=g.dtype).index_add(0, sid_of_pair, sym)
        return F, W
    return apply


def cpu_trainer(nn, structs, rc):
    """GrapFilterTrainer on the CPU over the oracle's neighbour lists."""
    elements = nn.elements
    I, J, Ds, types, sid = [], [], [], [], []
    off = 0
    for s, st in enumerate(structs):
        i, j, S = onl.neighbor_list(st['positions'], st['cell'], st['pbc'], rc)[:3]
        order = np.argsort(i, kind='stable')
        i, j, S = i[order], j[order], S[order]
        Ds.append(st['positions'][j] - st['positions'][i] + S @ st['cell'])
        I.append(i + off)
        J.append(j + off)
        sid.append(np.full(len(i), s))
        types.append([elements.index(x) for x in st['symbols']])
        off += len(st['positions'])
    i = torch.as_tensor(np.concatenate(I)).long()
    j = torch.as_tensor(np.concatenate(J)).long()
    D = torch.as_tensor(np.concatenate(Ds), dtype=torch.float64)
    tr = GrapFilterTrainer(
        nn, device='cpu',
        pair_force=torch_pair_force(i, j, torch.as_tensor(np.concatenate(sid)).long(), off,
                                    len(structs), D))
    for st in structs:
        tr.add_structure(st['atoms'], st['energy'], st['forces'], st['stress'])
    tr._batch = tr._make_batch(None, np.concatenate(types), i, j, D)
    return tr


@pytest.mark.parametrize("max_moment,symmetric,cutoff", [(0, False, 'cosine'),
                                                         (2, False, 'polynomial'),
                                                         (3, True, 'cosine')])
def test_filter_network_training_step_matches_oracle(max_moment, symmetric, cutoff):
    elements, rc = ['Mo', 'Ni'], 4.5
    structs = make_structures()
    with precision_scope('high'):
        nn = make_model(elements, rc, max_moment, symmetric, cutoff)
        tr = cpu_trainer(nn, structs, rc)
        loss, parts = tr.gradients()
        fp = filter_params(nn)
        leaves = [torch.tensor(w, dtype=torch.float64, requires_grad=True)
                  for w in fp['weights']] + \
                 [torch.tensor(v, dtype=torch.float64, requires_grad=True)
                  for v in fp['biases'] if v is not None]
        nW = len(fp['weights'])
        grid = dict(weights=leaves[:nW], biases=leaves[nW:] + [None],
                    activation=fp['activation'], use_resnet_dt=fp['use_resnet_dt'])
        grap = dict(algorithm='nn', grid=grid, moments=list(range(max_moment + 1)),
                    cutoff=cutoff, new_mode=True, symmetric=symmetric)
        params = {el: nn.mlp_params(el) for el in elements}
        ref_loss, ref_parts, ref_g = otr.loss_and_grads(elements, structs, params, rc,
                                                        angular=False, grap=grap,
                                                        extra_leaves=leaves)
    assert abs(loss.item() - ref_loss) < 1e-10 * max(1.0, abs(ref_loss))
    for key in ('energy', 'forces', 'stress'):
        assert abs(parts[key].item() - ref_parts[key]) < 1e-10
    for el in elements:
        for k, w in enumerate(tr.layers[el]['W']):
            r = ref_g[el][0][k]
            assert np.abs(w.grad.numpy() - r).max() < 1e-9 * max(1.0, np.abs(r).max()), (el, k)
    mine = [w.grad for w in tr.filters['W']] + [v.grad for v in tr.filters['b'] if v is not None]
    assert len(mine) == len(ref_g['__extra__'])
    for k, (g, r) in enumerate(zip(mine, ref_g['__extra__'])):
        assert g is not None and np.abs(r).max() > 0, k
        assert np.abs(g.numpy() - r).max() < 1e-9 * max(1.0, np.abs(r).max()), k


def test_filter_variables_round_trip_and_frozen_filters():
    elements, rc = ['Mo', 'Ni'], 4.5
    structs = make_structures(1)
    with precision_scope('high'):
        nn = make_model(elements, rc, 1, False)
        assert nn.get_variable("Atomic/Filters/Conv3d1/kernel").shape == (1, 1, 1, 1, 8)
        assert nn.get_variable("Atomic/Filters/Output/kernel").shape == (1, 1, 1, 8, 5)
        assert "Atomic/Filters/Output/bias" not in nn.variables        # output_bias=False
        with pytest.raises(ValueError, match="GrapFilterTrainer"):
            nn._device_model()
        with pytest.raises(NotImplementedError, match="GrapFilterTrainer"):
            nn.evaluate_batch(None)
        tr = cpu_trainer(nn, structs, rc)
        E0, F0, S0 = tr.evaluate()
        assert F0.shape == (len(structs[0]['positions']), 3) and S0.shape == (1, 6)
        with torch.no_grad():
            tr.filters['W'][0].mul_(1.5)
            tr.layers['Ni']['W'][0].mul_(0.5)
        tr.sync_to_model()
        tr2 = cpu_trainer(nn, structs, rc)
        E1, F1, _ = tr.evaluate()
        E2, F2, _ = tr2.evaluate()
        assert torch.equal(E1, E2) and torch.equal(F1, F2) and not torch.equal(E0, E1)
        # trainable = False: the filter weights are not leaves of the optimiser
        desc = GenericRadialAtomicPotential(elements, 'nn', dict(num_filters=3,
                                                                 hidden_sizes=[4],
                                                                 trainable=False),
                                            moment_tensors=0, legacy_mode=False)
        frozen = AtomicNN(elements, desc, hidden_sizes=[8], minmax_scale=False)
        frozen.attach_transformer(UniversalTransformer(elements, rcut=rc, angular=False))
        frozen.initialize_variables(seed=1)
        trf = cpu_trainer(frozen, structs, rc)
        assert all(not w.requires_grad for w in trf.filters['W'])
        assert len(trf.params) == sum(len(trf.layers[e]['W']) +
                                      sum(v is not None for v in trf.layers[e]['b'])
                                      for e in elements)


@pytest.mark.parametrize("modifier", [0, 1, 2])
def test_filter_evaluator_matches_oracle_energy_forces_stress(modifier):
    """The inference side (`AtomicNN._evaluate` of a model with the `nn` algorithm):
    per-atom energies, forces and virial against the oracle's autograd, CPU stand-in for the
    pair-force op.  `h_abck_modifier` 1 / 2 (grap.py:621-632): the filter input is r / r_cov or
    exp(-r / r_cov) of the centre element (Mo 1.54 A, Ni 1.24 A, Cordero 2008)."""
    from oracle import atomic as oat
    from tensoralloy_b200.nn.atomic.grap_nn import FilterEvaluator
    elements, rc = ['Mo', 'Ni'], 4.5
    st = make_structures(1, seed=8)[0]
    with precision_scope('high'):
        nn = make_model(elements, rc, 3, True, 'polynomial', h_abck_modifier=modifier)
        i, j, S = onl.neighbor_list(st['positions'], st['cell'], st['pbc'], rc)[:3]
        D = torch.as_tensor(st['positions'][j] - st['positions'][i] + S @ st['cell'])
        ti, tj = torch.as_tensor(i).long(), torch.as_tensor(j).long()
        n = len(st['positions'])
        ev = FilterEvaluator(nn, device='cpu',
                             pair_force=torch_pair_force(ti, tj, torch.zeros(len(i)).long(),
                                                         n, 1, D))
        types = np.array([elements.index(x) for x in st['symbols']])
        e_atom, F, W = ev(None, types, pairs=(ti, tj, D))
        fp = dict(filter_params(nn), h_abck_modifier=modifier, rcov={'Mo': 1.54, 'Ni': 1.24})
        grap = dict(algorithm='nn', grid=fp, moments=[0, 1, 2, 3], cutoff='polynomial',
                    new_mode=True, symmetric=True)
        params = {el: nn.mlp_params(el) for el in elements}
        ref = oat.atomic_evaluate(elements, st['symbols'], st['positions'], st['cell'],
                                  st['pbc'], rc, params, angular=False, grap=grap)
        if modifier:
            # the modifier changes the result (it is not silently ignored)
            plain = dict(grap, grid=dict(fp, h_abck_modifier=0))
            other = oat.atomic_evaluate(elements, st['symbols'], st['positions'], st['cell'],
                                        st['pbc'], rc, params, angular=False, grap=plain)
            assert abs(other['energy'] - ref['energy']) > 1e-6
    vol = abs(np.linalg.det(st['cell']))
    assert abs(e_atom.sum().item() - ref['energy']) / n < 1e-12
    assert np.abs(F.numpy() - ref['forces']).max() < 1e-10
    st6 = np.array([W[0, a, b].item() for a, b in ((0, 0), (1, 1), (2, 2), (1, 2), (0, 2),
                                                    (0, 1))]) / vol
    assert np.abs(st6 - ref['stress']).max() < 1e-10
    assert np.abs(ref['forces']).max() > 1e-3


@pytest.mark.parametrize("algorithm,max_moment", [('pexp', 4), ('nn', 5)])
def test_moments_4_and_5_use_the_full_moment_tensors(algorithm, max_moment):
    """grap.py:537-594, 655-660: with max_moment > 3 every moment is the full 3^m tensor with
    unit weights and `symmetric` has no effect; closed-form algorithms take the same torch
    path as the filter network.  Loss and every parameter gradient against the oracle."""
    elements, rc = ['Mo', 'Ni'], 4.5
    structs = make_structures()
    with precision_scope('high'):
        nn = make_model(elements, rc, max_moment, True, 'cosine', algorithm)
        assert nn.descriptor.uses_torch_path()
        assert nn.descriptor.dimension() == 2 * (5 if algorithm == 'nn' else 3) * (max_moment + 1)
        with pytest.raises(ValueError, match="GrapFilterTrainer"):
            nn._device_model()
        tr = cpu_trainer(nn, structs, rc)
        loss, parts = tr.gradients()
        leaves = []
        if algorithm == 'nn':
            fp = filter_params(nn)
            leaves = [torch.tensor(w, dtype=torch.float64, requires_grad=True)
                      for w in fp['weights']] + \
                     [torch.tensor(v, dtype=torch.float64, requires_grad=True)
                      for v in fp['biases'] if v is not None]
            nW = len(fp['weights'])
            grid = dict(weights=leaves[:nW], biases=leaves[nW:] + [None],
                        activation=fp['activation'], use_resnet_dt=fp['use_resnet_dt'])
        else:
            grid = nn.descriptor.radial_sets()
        grap = dict(algorithm=algorithm, grid=grid, moments=list(range(max_moment + 1)),
                    cutoff='cosine', new_mode=True, symmetric=True)
        params = {el: nn.mlp_params(el) for el in elements}
        ref_loss, ref_parts, ref_g = otr.loss_and_grads(elements, structs, params, rc,
                                                        angular=False, grap=grap,
                                                        extra_leaves=leaves)
    assert abs(loss.item() - ref_loss) < 1e-10 * max(1.0, abs(ref_loss))
    for key in ('energy', 'forces', 'stress'):
        assert abs(parts[key].item() - ref_parts[key]) < 1e-10
    for el in elements:
        for k, w in enumerate(tr.layers[el]['W']):
            r = ref_g[el][0][k]
            assert np.abs(w.grad.numpy() - r).max() < 1e-9 * max(1.0, np.abs(r).max()), (el, k)
    if algorithm == 'nn':
        mine = [w.grad for w in tr.filters['W']] + \
               [v.grad for v in tr.filters['b'] if v is not None]
        for g, r in zip(mine, ref_g['__extra__']):
            assert np.abs(g.numpy() - r).max() < 1e-9 * max(1.0, np.abs(r).max())
    else:
        assert tr.filters is None
        tr.sync_to_model()


def test_closed_form_descriptors_on_the_torch_path_equal_the_kernel_formulation():
    """For moments <= 3 the torch path and the oracle's legacy-style restatement (which the
    CUDA kernels are tested against) must give the same descriptors: ties the two product
    paths together."""
    from oracle import atomic as oat
    from tensoralloy_b200.nn.atomic.grap_nn import closed_form_radial, filter_descriptors
    st = make_structures(1, seed=3)[0]
    elements, rc = ['Mo', 'Ni'], 4.5
    i, j, S = onl.neighbor_list(st['positions'], st['cell'], st['pbc'], rc)[:3]
    D = torch.as_tensor(st['positions'][j] - st['positions'][i] + S @ st['cell'])
    types = np.array([elements.index(x) for x in st['symbols']])
    ti, tj = torch.as_tensor(types[i]), torch.as_tensor(types[j])
    term = torch.where(ti == tj, torch.zeros_like(ti), tj - (tj > ti).long() + 1)
    n = len(types)
    for algorithm, grid, cutoff in (('sf', [(0.5, 0.0), (4.0, 1.0)], 'cosine'),
                                    ('morse', [(0.5, 1.2, 2.2), (1.0, 0.8, 2.6)], 'polynomial'),
                                    ('density', [(1.0, 3.0, 2.2), (2.0, 5.0, 2.4)], 'cosine'),
                                    ('pexp', [(1.5, 2.0), (2.5, 3.0)], 'polynomial')):
        G = filter_descriptors(D, torch.as_tensor(i).long() * 2 + term, n * 2,
                               closed_form_radial(algorithm, grid, rc), cutoff, rc, 2, False,
                               1e-14).reshape(n, -1)
        ref = oat.grap_descriptors(elements, types, torch.as_tensor(st['positions']),
                                   torch.as_tensor(st['cell']), i, j, S, rc, algorithm, grid,
                                   (0, 1, 2), cutoff)
        assert float((G - ref).abs().max()) < 1e-10 * max(1.0, float(ref.abs().max())), algorithm


def test_training_loop_with_the_train_op_reduces_the_loss():
    """trainer.train_step(TrainOp): gradients -> update -> moving averages, end to end on the
    CPU stand-in; the loss of the fixed batch goes down and the averaged parameters can be
    written back into the model."""
    from tensoralloy_b200.nn.opt import OptParameters, TrainOp
    elements, rc = ['Mo', 'Ni'], 4.5
    structs = make_structures(2, seed=11)
    with precision_scope('high'):
        nn = make_model(elements, rc, 2, False)
        tr = cpu_trainer(nn, structs, rc)
        op = TrainOp(tr.params, OptParameters(method='adam', learning_rate=2e-3,
                                              decay_function='exponential', decay_rate=0.9,
                                              decay_steps=10))
        first, _ = tr.train_step(op)
        for _ in range(40):
            last, parts = tr.train_step(op)
        assert op.global_step == 41 and abs(op.learning_rate() - 2e-3 * 0.9 ** 4.1) < 1e-15
        assert last.item() < 0.9 * first.item(), (first.item(), last.item())
        assert set(parts) == {'energy', 'forces', 'stress'}
        before = nn.get_variable("Atomic/Filters/Output/kernel").copy()
        op.swap_in_ema()
        tr.sync_to_model()
        after = nn.get_variable("Atomic/Filters/Output/kernel")
        assert after.shape == before.shape and not np.array_equal(after, before)
