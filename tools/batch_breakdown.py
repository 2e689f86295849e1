"""Wall/GPU-time breakdown of one calculate_batch call (dev tool): C2 Be x B and C1 Ni x B."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tensoralloy_b200 import _lib
from tensoralloy_b200.atoms import Atoms, bulk_fcc
from tensoralloy_b200.nn.atomic import AtomicNN, SymmetryFunction
from tensoralloy_b200.nn.eam import EamAlloyNN
from tensoralloy_b200.precision import precision_scope, get_float_dtype
from tensoralloy_b200.transformer import UniversalTransformer

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256


def sync():
    torch.cuda.synchronize()
    return time.perf_counter()


def run(name, nn, clf, frames):
    nn.attach_transformer(clf)
    model = nn._device_model()
    for rep in range(3):
        t0 = sync()
        bf = clf.get_batch_features(frames)
        t1 = sync()
        raws = nn.evaluate_batch(bf, True, True, True)
        t2 = sync()
        n, nb = bf.n_atoms, bf.n_struct
        e = torch.zeros(nb, dtype=torch.float64, device='cuda')
        v = torch.zeros((nb, 9), dtype=torch.float64, device='cuda')
        f = torch.zeros((n, 3), dtype=torch.float64, device='cuda')
        ea = torch.zeros(n, dtype=torch.float64, device='cuda')
        t3 = sync()
        for _ in range(5):
            model.eval(bf.nbr, get_float_dtype().tab_precision, energy=e, eatom=ea, forces=f, virial=v)
        t4 = sync()
        # the list build alone (device arrays already there)
        d_types = torch.as_tensor(bf.types).cuda()
        cells = np.stack([np.asarray(a.cell, dtype=float).reshape(3, 3) for a in frames])
        pbcs = np.ones((nb, 3), dtype=bool)
        t5 = sync()
        for _ in range(5):
            bf.nbr.build_batch(bf.d_pos, d_types, bf.offsets, cells, pbcs, clf.rcut)
        t6 = sync()
    print(f"{name} B={nb} atoms={n} nij={bf.nbr.sizes()[0]}: get_batch_features {1e3*(t1-t0):.2f} ms "
          f"(of which build_batch {1e3*(t6-t5)/5:.2f}) evaluate_batch {1e3*(t2-t1):.2f} ms "
          f"(of which kernels {1e3*(t4-t3)/5:.2f})", flush=True)


with precision_scope('high'):
    d = np.load(os.path.join(ROOT, 'tests', 'golden', 'Be_liquid_4000K.npz'))
    rng = np.random.default_rng(1)
    frames = [Atoms(list(d['symbols']), d['positions'][1 + k % 2] + rng.normal(scale=0.05, size=(128, 3)),
                    d['cells'][1], True) for k in range(B)]
    nn = AtomicNN(['Be'], SymmetryFunction(['Be']), minmax_scale=False,
                  export_properties=('energy', 'forces', 'stress'))
    run('C2', nn, UniversalTransformer(['Be'], rcut=5.0, acut=5.0, angular=True), frames)
    frames = []
    for k in range(B):
        a = bulk_fcc('Ni', 3.52, (4, 4, 4))
        a.positions += rng.normal(scale=0.05, size=a.positions.shape)
        frames.append(a)
    nn = EamAlloyNN(['Ni'], custom_potentials='zjw04', export_properties=['energy', 'forces', 'stress'])
    run('C1', nn, UniversalTransformer(['Ni'], rcut=6.5), frames)
