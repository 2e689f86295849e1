"""Wall-time breakdown of one TensorAlloyCalculator.calculate call (dev tool)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tensoralloy_b200.atoms import Atoms
from tensoralloy_b200.nn.atomic import AtomicNN, SymmetryFunction
from tensoralloy_b200.precision import precision_scope, get_float_dtype
from tensoralloy_b200.transformer import UniversalTransformer

d = np.load('tests/golden/Be_liquid_4000K.npz')
atoms = Atoms(list(d['symbols']), d['positions'][1], d['cells'][1], True)
with precision_scope('high'):
    nn = AtomicNN(['Be'], SymmetryFunction(['Be']), minmax_scale=False,
                  export_properties=('energy', 'forces', 'stress'))
    clf = UniversalTransformer(['Be'], rcut=5.0, acut=5.0, angular=True)
    nn.attach_transformer(clf)
    nn.initialize_variables()
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        feats = clf.get_constant_features(atoms)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        raw = nn._evaluate(feats, True, True, True)
        torch.cuda.synchronize(); t2 = time.perf_counter()
        model = nn._device_model()
        n = len(atoms)
        e = torch.zeros(16, dtype=torch.float64, device='cuda'); f = torch.zeros((n, 3), dtype=torch.float64, device='cuda')
        torch.cuda.synchronize(); t3 = time.perf_counter()
        for _ in range(10):
            model.eval(feats.nbr, 0, energy=e[0:1], forces=f, virial=e[1:10])
        torch.cuda.synchronize(); t4 = time.perf_counter()
        print(f"features(build) {1e3*(t1-t0):.3f} ms  evaluate(first, incl reverse idx + D2H) {1e3*(t2-t1):.3f} ms  kernels only {1e2*(t4-t3):.3f} ms/eval")
