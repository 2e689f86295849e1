"""Multi-GPU check of the AtomicNN slab decomposition (run under torchrun, one rank per
GPU): 2 rc halo exchange over NCCL send/recv against a single-GPU evaluation of the same
structure.  Exit code 0 = every rank agrees to 1e-10 eV/atom, 1e-8 eV/A.  Prints the step
time (max over ranks is taken by the caller from the per-rank lines)."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tensoralloy_b200.atoms import Atoms, bulk_fcc                       # noqa: E402
from tensoralloy_b200.domain_atomic import AtomicSlabDomain             # noqa: E402
from tensoralloy_b200.nn.atomic import AtomicNN, SymmetryFunction       # noqa: E402
from tensoralloy_b200.precision import precision_scope                  # noqa: E402
from tensoralloy_b200.transformer import UniversalTransformer           # noqa: E402


def main():
    world = int(os.environ['WORLD_SIZE'])
    rank = int(os.environ['RANK'])
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    nx = int(sys.argv[1]) if len(sys.argv) > 1 else 6 * world
    ny = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    elements, rc, acut = ['Mo', 'Ni'], 4.6, 4.0
    rng = np.random.default_rng(9)
    base = bulk_fcc('Ni', 3.6, (nx, ny, ny))
    sym = ['Mo' if x < 0.4 else 'Ni' for x in rng.random(len(base))]
    pos = base.positions + rng.normal(scale=0.1, size=base.positions.shape)
    atoms = Atoms(sym, pos, base.cell, True)
    with precision_scope('high'):
        nn = AtomicNN(elements, SymmetryFunction(elements), minmax_scale=False,
                      hidden_sizes=[32, 16], export_properties=('energy', 'forces', 'stress'))
        clf = UniversalTransformer(elements, rcut=rc, acut=acut, angular=True)
        nn.attach_transformer(clf)
        nn.initialize_variables(seed=7)
        for el in nn.elements:
            key = f"Atomic/{el}/Output/kernel"
            nn.set_variable(key, nn.get_variable(key) * 0.02)
        model = nn._device_model()
        types = clf.get_types(atoms)
        dom = AtomicSlabDomain(model, pos, types, np.asarray(atoms.cell), max(rc, acut),
                               world, rank)
        for _ in range(2):
            E, F, V = dom.step()
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        for _ in range(5):
            E, F, V = dom.step()
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) / 5 * 1e3
        # single-GPU answer on this rank
        raw = nn._evaluate(clf.get_constant_features(atoms), True, True, True)
    n = len(atoms)
    de = abs(E - raw['energy']) / n
    df = np.abs(F.cpu().numpy() - raw['forces'][dom.owned]).max()
    dv = np.abs(V - raw['virial']).max() / n
    good = de < 1e-10 and df < 1e-8 and dv < 1e-8
    print(f"rank {rank}/{world}: atoms {n} own {len(dom.owned)} rows "
          f"{dom.rank_state.n_rows} step {ms:.3f} ms dE/N={de:.2e} dF={df:.2e} dV/N={dv:.2e} "
          f"{'OK' if good else 'FAIL'}", flush=True)
    flag = torch.tensor([1 if good else 0], device='cuda')
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == '__main__':
    main()
