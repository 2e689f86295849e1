#!/usr/bin/env python
"""
Secondary measurements for the other BASELINE.json configs (bench.py keeps the
headline 1M-atom EAM line).  One JSON line per config; CUDA-event / wall timing
with a synchronize on both sides, warm-up excluded.

  C1  EAM Ni fcc 4x4x4 (256 atoms) through TensorAlloyCalculator (E+F+stress)
  C2  AtomicNN G2+G4, Be 128-atom liquids, batch of B structures (E+F+stress)
  C4  AtomicNN training step (Mo-Ni, energy+forces+stress loss, parameter grads)
  C5  analytic Hessian, Be hcp 5x5x3 (150 atoms), float64
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from tensoralloy_b200.atoms import Atoms, bulk_fcc, bulk_hcp          # noqa: E402
from tensoralloy_b200.calculator import TensorAlloyCalculator        # noqa: E402
from tensoralloy_b200.nn.atomic import AtomicNN, SymmetryFunction    # noqa: E402
from tensoralloy_b200.nn.atomic.training import AtomicNNTrainer      # noqa: E402
from tensoralloy_b200.nn.eam import EamAlloyNN                        # noqa: E402
from tensoralloy_b200.precision import precision_scope               # noqa: E402
from tensoralloy_b200.transformer import UniversalTransformer        # noqa: E402


def timed(fn, reps, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


def c1():
    atoms = bulk_fcc('Ni', 3.52, (4, 4, 4))
    atoms.positions += np.random.default_rng(611).normal(scale=0.05,
                                                         size=atoms.positions.shape)
    nn = EamAlloyNN(['Ni'], custom_potentials='zjw04',
                    export_properties=['energy', 'forces', 'stress'])
    nn.attach_transformer(UniversalTransformer(['Ni'], rcut=6.5))
    calc = TensorAlloyCalculator(nn)
    t = timed(lambda: calc.calculate(atoms, ['energy', 'forces', 'stress']), 50)
    return {"config": "C1 EAM Ni fcc 4x4x4 (256 atoms), zjw04 rc 6.5, "
                      "TensorAlloyCalculator.calculate (H2D, list build, E+F+stress, D2H)",
            "ms_per_call": t * 1e3, "atom_evals_per_s": 256 / t, "dtype": "f64"}


def c2(batch, batched=False):
    d = np.load(os.path.join(ROOT, 'tests', 'golden', 'Be_liquid_4000K.npz'))
    rng = np.random.default_rng(1)
    frames = []
    for k in range(batch):
        base = d['positions'][1 + k % 2]
        frames.append(Atoms(list(d['symbols']),
                            base + rng.normal(scale=0.05, size=base.shape),
                            d['cells'][1], True))
    nn = AtomicNN(['Be'], SymmetryFunction(['Be']), minmax_scale=False,
                  export_properties=('energy', 'forces', 'stress'))
    nn.attach_transformer(UniversalTransformer(['Be'], rcut=5.0, acut=5.0, angular=True))
    nn.initialize_variables()
    calc = TensorAlloyCalculator(nn)

    def run():
        for a in frames:
            calc.calculate(a, ['energy', 'forces', 'stress'])
    if not batched:
        t = timed(run, 5, warm=2)
        how = "sequential calculate() calls"
    else:
        t = timed(lambda: calc.calculate_batch(frames, ('energy', 'forces', 'stress')), 5,
                  warm=2)
        how = "ONE calculate_batch() call: H2D, batch list build, E+F+stress, D2H"
    return {"config": f"C2 AtomicNN G2+G4 (eta 4, beta 1, gamma 2, zeta 2) + MLP [64,32], "
                      f"Be 128 atoms rc=acut=5.0, batch {batch} ({how})",
            "ms_per_batch": t * 1e3, "structures_per_s": batch / t,
            "atom_evals_per_s": batch * 128 / t, "dtype": "f64"}


def c1_batch(batch=256):
    rng = np.random.default_rng(611)
    frames = []
    for k in range(batch):
        atoms = bulk_fcc('Ni', 3.52, (4, 4, 4))
        atoms.positions += rng.normal(scale=0.05, size=atoms.positions.shape)
        frames.append(atoms)
    nn = EamAlloyNN(['Ni'], custom_potentials='zjw04',
                    export_properties=['energy', 'forces', 'stress'])
    nn.attach_transformer(UniversalTransformer(['Ni'], rcut=6.5))
    calc = TensorAlloyCalculator(nn)
    t = timed(lambda: calc.calculate_batch(frames, ('energy', 'forces', 'stress')), 5, warm=2)
    return {"config": f"C1 x {batch}: EAM Ni fcc 4x4x4 (256 atoms) zjw04 rc 6.5, ONE "
                      "calculate_batch() call (H2D, batch list build, E+F+stress, D2H)",
            "ms_per_batch": t * 1e3, "structures_per_s": batch / t,
            "atom_evals_per_s": batch * 256 / t, "dtype": "f64"}


def c4(n_struct=32):
    rng = np.random.default_rng(0)
    elements = ['Mo', 'Ni']
    nn = AtomicNN(elements, SymmetryFunction(elements), hidden_sizes=[64, 32],
                  minmax_scale=False,
                  minimize_properties=('energy', 'forces', 'stress'),
                  export_properties=('energy', 'forces', 'stress'))
    nn.attach_transformer(UniversalTransformer(elements, rcut=5.0, angular=True))
    nn.initialize_variables()
    tr = AtomicNNTrainer(nn)
    for _ in range(n_struct):
        a = 3.3 + 0.4 * rng.random()
        base = bulk_fcc('Ni', a, (3, 3, 3))
        sym = ['Mo' if x < 0.4 else 'Ni' for x in rng.random(len(base))]
        atoms = Atoms(sym, base.positions + rng.normal(scale=0.1, size=base.positions.shape),
                      base.cell, True)
        tr.add_structure(atoms, -4.0 * len(base), rng.normal(scale=0.3, size=(len(base), 3)),
                         rng.normal(scale=0.01, size=6))
    opt = torch.optim.Adam(tr.params, lr=1e-3)
    graphed = tr.enable_graph()
    t = timed(lambda: tr.train_step(opt), 5, warm=2)
    return {"config": f"C4 AtomicNN training step, Mo-Ni, {n_struct} structures x 108 atoms, "
                      f"G2+G4 D=20, MLP [64,32], loss = E/atom + F + stress RMSE, Adam"
                      f"{', loss + backward as one CUDA graph' if graphed else ''}",
            "ms_per_step": t * 1e3, "structures_per_s": n_struct / t, "dtype": "f64"}


def c4_adp(n_struct=256):
    from tensoralloy_b200.nn.eam import AdpNN
    from tensoralloy_b200.nn.eam.training import EamTrainer
    rng = np.random.default_rng(0)
    elements = ['Mo', 'Ni']
    nn = AdpNN(elements, hidden_sizes=[32, 32],
               minimize_properties=('energy', 'forces', 'stress'))
    nn.attach_transformer(UniversalTransformer(elements, rcut=5.0))
    nn.initialize_variables()
    tr = EamTrainer(nn)
    for _ in range(n_struct):
        a = 3.3 + 0.4 * rng.random()
        base = bulk_fcc('Ni', a, (3, 3, 3))
        sym = ['Mo' if x < 0.4 else 'Ni' for x in rng.random(len(base))]
        atoms = Atoms(sym, base.positions + rng.normal(scale=0.1, size=base.positions.shape),
                      base.cell, True)
        tr.add_structure(atoms, -4.0 * len(base), rng.normal(scale=0.3, size=(len(base), 3)),
                         rng.normal(scale=0.01, size=6))
    opt = torch.optim.Adam(tr.params, lr=1e-3)
    graphed = tr.enable_graph()
    t = timed(lambda: tr.train_step(opt), 5, warm=2)
    return {"config": f"C4(ii) AdpNN training step, Mo-Ni, {n_struct} structures x 108 atoms, "
                      f"phi / rho / embed / dipole / quadrupole as 'nn' MLPs [32,32], rc 5.0, "
                      f"loss = E/atom + F + stress RMSE, Adam"
                      f"{', loss + backward as one CUDA graph' if graphed else ''}",
            "ms_per_step": t * 1e3, "structures_per_s": n_struct / t, "dtype": "f64"}


def c5():
    atoms = bulk_hcp('Be', 2.2644, 3.5673, (5, 5, 3))
    atoms.positions += np.random.default_rng(8).normal(scale=0.01,
                                                       size=atoms.positions.shape)
    nn = EamAlloyNN(['Be'], custom_potentials='Be/1',
                    export_properties=['energy', 'forces', 'hessian'])
    nn.attach_transformer(UniversalTransformer(['Be'], rcut=5.0))
    calc = TensorAlloyCalculator(nn)
    t = timed(lambda: calc.calculate(atoms, ['energy', 'forces', 'hessian']), 10)
    return {"config": "C5 analytic Hessian, Be hcp 5x5x3 (150 atoms), AgrawalBe rc 5.0 "
                      "(list build + E+F + dense [450,450] Hessian + D2H)",
            "ms_per_call": t * 1e3, "dtype": "f64"}


def medium():
    """'medium' (float32 kernels, the reference's default precision) on the batch path."""
    out = []
    with precision_scope('medium'):
        for fn in (c1_batch, lambda: c2(256, True)):
            r = fn()
            r['dtype'] = 'f32'
            out.append(r)
    return out


def main():
    with precision_scope('high'):
        for fn in (c1, c1_batch, lambda: c2(1), lambda: c2(32), lambda: c2(32, True),
                   lambda: c2(256, True), c4, lambda: c4(256), lambda: c4_adp(32),
                   c4_adp, c5):
            print(json.dumps(fn()), flush=True)
    for r in medium():
        print(json.dumps(r), flush=True)


if __name__ == '__main__':
    main()
