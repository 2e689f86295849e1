"""Small end-to-end pass over the kernels added this round, meant to run under
`compute-sanitizer --tool memcheck` (dev tool): batch lists, EAM / ADP / AtomicNN on a batch,
pair operators, decomposed evaluation with recomputed inner halo."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tensoralloy_b200.atoms import Atoms, bulk_fcc
from tensoralloy_b200.calculator import TensorAlloyCalculator
from tensoralloy_b200.domain_atomic import run_loopback
from tensoralloy_b200.nn.atomic import AtomicNN, SymmetryFunction
from tensoralloy_b200.nn.atomic.training import AtomicNNTrainer
from tensoralloy_b200.nn.eam import AdpNN, EamAlloyNN
from tensoralloy_b200.nn.eam.training import EamTrainer
from tensoralloy_b200.precision import precision_scope
from tensoralloy_b200.transformer import UniversalTransformer

rng = np.random.default_rng(0)
EL = ['Mo', 'Ni']


def alloy(reps, a=3.6):
    base = bulk_fcc('Ni', a, reps)
    sym = ['Mo' if x < 0.4 else 'Ni' for x in rng.random(len(base))]
    return Atoms(sym, base.positions + rng.normal(scale=0.08, size=base.positions.shape),
                 base.cell, True)


with precision_scope('high'):
    images = [alloy((2, 2, 2)), alloy((3, 2, 2)), alloy((1, 1, 1)), alloy((2, 3, 2))]
    cp = {'Mo': {'rho': 'zjw04', 'embed': 'zjw04'}, 'Ni': {'rho': 'zjw04', 'embed': 'zjw04'},
          'MoMo': {'phi': 'zjw04', 'dipole': 'mishinh', 'quadrupole': 'mishinh'},
          'MoNi': {'phi': 'zjw04', 'dipole': 'mishinh', 'quadrupole': 'mishinh'},
          'NiNi': {'phi': 'zjw04', 'dipole': 'mishinh', 'quadrupole': 'mishinh'}}
    for nn in (EamAlloyNN(EL, custom_potentials='zjw04',
                          export_properties=('energy', 'forces', 'stress')),
               AdpNN(EL, custom_potentials=cp, export_properties=('energy', 'forces', 'stress'))):
        nn.attach_transformer(UniversalTransformer(EL, rcut=5.0))
        res = TensorAlloyCalculator(nn).calculate_batch(images)
        print(type(nn).__name__, 'batch', [float(r['energy']) for r in res][:2])
        big = alloy((8, 2, 2))
        e, f, w, _ = run_loopback(nn._device_model(), big.positions,
                                  nn.transformer.get_types(big), np.asarray(big.cell), 5.0, 2)
        print(type(nn).__name__, 'dd', e)
    an = AtomicNN(EL, SymmetryFunction(EL), hidden_sizes=[8, 8], minmax_scale=False,
                  minimize_properties=('energy', 'forces', 'stress'),
                  export_properties=('energy', 'forces', 'stress'))
    an.attach_transformer(UniversalTransformer(EL, rcut=4.5, acut=4.0, angular=True))
    an.initialize_variables(seed=1)
    print('atomic batch', float(TensorAlloyCalculator(an).calculate_batch(images)[0]['energy']))
    big = alloy((8, 2, 2))
    print('atomic dd', run_loopback(an._device_model(), big.positions,
                                    an.transformer.get_types(big), np.asarray(big.cell), 4.5, 2)[0])
    tr = AtomicNNTrainer(an)
    for a in images[:2]:
        tr.add_structure(a, -4.0 * len(a), rng.normal(size=(len(a), 3)), rng.normal(size=6))
    print('atomic train', float(tr.gradients()[0]))
    en = EamAlloyNN(EL, hidden_sizes=[4, 4], minimize_properties=('energy', 'forces', 'stress'))
    en.attach_transformer(UniversalTransformer(EL, rcut=5.0))
    et = EamTrainer(en)
    for a in images[:2]:
        et.add_structure(a, -4.0 * len(a), rng.normal(size=(len(a), 3)), rng.normal(size=6))
    print('eam train', float(et.gradients()[0]))
torch.cuda.synchronize()
print('done')
