"""cProfile of repeated TensorAlloyCalculator.calculate calls on a small structure (dev tool)."""
import cProfile, os, pstats, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tensoralloy_b200.atoms import bulk_fcc
from tensoralloy_b200.calculator import TensorAlloyCalculator
from tensoralloy_b200.nn.eam import EamAlloyNN
from tensoralloy_b200.precision import precision_scope
from tensoralloy_b200.transformer import UniversalTransformer

with precision_scope('high'):
    atoms = bulk_fcc('Ni', 3.52, (4, 4, 4))
    atoms.positions += np.random.default_rng(611).normal(scale=0.05, size=atoms.positions.shape)
    nn = EamAlloyNN(['Ni'], custom_potentials='zjw04', export_properties=['energy', 'forces', 'stress'])
    nn.attach_transformer(UniversalTransformer(['Ni'], rcut=6.5))
    calc = TensorAlloyCalculator(nn)
    for _ in range(20):
        calc.calculate(atoms, ['energy', 'forces', 'stress'])
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(300):
        calc.calculate(atoms, ['energy', 'forces', 'stress'])
    pr.disable()
    pstats.Stats(pr).sort_stats('cumulative').print_stats(28)
