"""GRAP `nn` algorithm (filter network) through the library on the GPU: batched lists,
`tab_pairs_export`, `tab_pair_forces` / `tab_pair_jvp` under `GrapFilterTrainer`, against the
oracle's double backward (same comparison as tests/test_grap_filter_network.py, which
replaces the library's pair-force op by its torch definition)."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(__file__))
import test_grap_filter_network as cpu          # noqa: E402  (helpers only)
from oracle import training as otr              # noqa: E402
from tensoralloy_b200.nn.atomic.grap_nn import GrapFilterTrainer, filter_params  # noqa: E402
from tensoralloy_b200.precision import precision_scope                           # noqa: E402

pytestmark = pytest.mark.gpu


def test_filter_network_training_step_on_the_gpu():
    elements, rc = ['Mo', 'Ni'], 4.5
    structs = cpu.make_structures()
    with precision_scope('high'):
        nn = cpu.make_model(elements, rc, 3, True)
        tr = GrapFilterTrainer(nn)
        for st in structs:
            tr.add_structure(st['atoms'], st['energy'], st['forces'], st['stress'])
        loss, parts = tr.gradients()
        fp = filter_params(nn)
        leaves = [torch.tensor(w, dtype=torch.float64, requires_grad=True)
                  for w in fp['weights']] + \
                 [torch.tensor(v, dtype=torch.float64, requires_grad=True)
                  for v in fp['biases'] if v is not None]
        nW = len(fp['weights'])
        grid = dict(weights=leaves[:nW], biases=leaves[nW:] + [None],
                    activation=fp['activation'], use_resnet_dt=fp['use_resnet_dt'])
        grap = dict(algorithm='nn', grid=grid, moments=[0, 1, 2, 3], cutoff='cosine',
                    new_mode=True, symmetric=True)
        params = {el: nn.mlp_params(el) for el in elements}
        ref_loss, ref_parts, ref_g = otr.loss_and_grads(elements, structs, params, rc,
                                                        angular=False, grap=grap,
                                                        extra_leaves=leaves)
        assert abs(loss.item() - ref_loss) < 1e-9 * max(1.0, abs(ref_loss))
        for key in ('energy', 'forces', 'stress'):
            assert abs(parts[key].item() - ref_parts[key]) < 1e-9
        for el in elements:
            for k, w in enumerate(tr.layers[el]['W']):
                r = ref_g[el][0][k]
                assert np.abs(w.grad.cpu().numpy() - r).max() < 1e-8 * max(1.0, np.abs(r).max())
        mine = [w.grad for w in tr.filters['W']] + \
               [v.grad for v in tr.filters['b'] if v is not None]
        for g, r in zip(mine, ref_g['__extra__']):
            assert np.abs(g.cpu().numpy() - r).max() < 1e-8 * max(1.0, np.abs(r).max())
        # E / F / stress of the same structures through the same ops
        E, F, S = tr.evaluate()
        assert E.shape == (2,) and F.shape[1] == 3 and S.shape == (2, 6)
        assert torch.isfinite(F).all()


def test_calculator_serves_a_filter_network_model():
    """TensorAlloyCalculator over an AtomicNN with the GRAP `nn` algorithm: energy, forces and
    stress against the oracle (1e-10 eV/atom, 1e-8 eV/A)."""
    from oracle import atomic as oat
    from tensoralloy_b200.calculator import TensorAlloyCalculator
    elements, rc = ['Mo', 'Ni'], 4.5
    st = cpu.make_structures(1, seed=8)[0]
    atoms = st['atoms']
    with precision_scope('high'):
        nn = cpu.make_model(elements, rc, 3, True, 'polynomial')
        calc = TensorAlloyCalculator(nn)
        calc.calculate(atoms, properties=['energy', 'forces', 'stress'])
        e, f, s = calc.results['energy'], calc.get_forces(atoms), calc.get_stress(atoms)
        fp = filter_params(nn)
        grap = dict(algorithm='nn', grid=fp, moments=[0, 1, 2, 3], cutoff='polynomial',
                    new_mode=True, symmetric=True)
        params = {el: nn.mlp_params(el) for el in elements}
        ref = oat.atomic_evaluate(elements, st['symbols'], st['positions'], st['cell'],
                                  st['pbc'], rc, params, angular=False, grap=grap)
    assert abs(e - ref['energy']) / len(atoms) < 1e-10
    assert np.abs(f - ref['forces']).max() < 1e-8
    assert np.abs(s - ref['stress']).max() < 1e-8
