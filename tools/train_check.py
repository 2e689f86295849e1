"""Structure-parallel training step across GPUs (config 4; run under torchrun, one rank per
GPU): every rank evaluates loss and parameter gradients of ITS share of the batch (one batch
neighbour handle per rank), the gradients are averaged with one flat NCCL all-reduce
(reference: MirroredStrategy + MEAN aggregation).  Check: the all-reduced gradient equals
the mean of the per-share gradients recomputed locally.  Prints structures/s."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tensoralloy_b200.atoms import Atoms, bulk_fcc                       # noqa: E402
from tensoralloy_b200.nn.atomic import AtomicNN, SymmetryFunction       # noqa: E402
from tensoralloy_b200.nn.atomic.training import AtomicNNTrainer         # noqa: E402
from tensoralloy_b200.precision import precision_scope                  # noqa: E402
from tensoralloy_b200.transformer import UniversalTransformer           # noqa: E402

ELEMENTS = ['Mo', 'Ni']


def make_structures(n_struct, seed=0):
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n_struct):
        a = 3.3 + 0.4 * rng.random()
        base = bulk_fcc('Ni', a, (3, 3, 3))
        sym = ['Mo' if x < 0.4 else 'Ni' for x in rng.random(len(base))]
        atoms = Atoms(sym, base.positions + rng.normal(scale=0.1, size=base.positions.shape),
                      base.cell, True)
        out.append((atoms, -4.0 * len(base) + rng.normal(),
                    rng.normal(scale=0.3, size=(len(base), 3)), rng.normal(scale=0.01, size=6)))
    return out


def trainer(structs):
    nn = AtomicNN(ELEMENTS, SymmetryFunction(ELEMENTS), hidden_sizes=[64, 32],
                  minmax_scale=False, minimize_properties=('energy', 'forces', 'stress'))
    nn.attach_transformer(UniversalTransformer(ELEMENTS, rcut=5.0, angular=True))
    nn.initialize_variables(seed=3)
    tr = AtomicNNTrainer(nn)
    for s in structs:
        tr.add_structure(*s)
    return tr


def main():
    world = int(os.environ['WORLD_SIZE'])
    rank = int(os.environ['RANK'])
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    per_rank = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    check = per_rank <= 8
    structs = make_structures(per_rank * world)
    with precision_scope('high'):
        tr = trainer(structs[rank * per_rank:(rank + 1) * per_rank])
        tr.gradients()
        tr.allreduce_gradients(dist, world)
        mine = [p.grad.clone() for p in tr.params]
        ok = True
        if check:       # mean of the per-share gradients, all recomputed on this rank
            acc = None
            for r in range(world):
                t2 = trainer(structs[r * per_rank:(r + 1) * per_rank])
                t2.gradients()
                g = [p.grad for p in t2.params]
                acc = g if acc is None else [a + b for a, b in zip(acc, g)]
            err = max(float((a / world - b).abs().max()) for a, b in zip(acc, mine))
            scale = max(float(b.abs().max()) for b in mine)
            ok = err < 1e-10 * max(1.0, scale)
            print(f"rank {rank}: all-reduced gradient vs local mean: {err:.2e} "
                  f"(scale {scale:.2e}) {'OK' if ok else 'FAIL'}", flush=True)
        opt = torch.optim.Adam(tr.params, lr=1e-3)
        for _ in range(3):
            tr.train_step(opt, dist, world)
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        steps = 10
        for _ in range(steps):
            tr.train_step(opt, dist, world)
        torch.cuda.synchronize()
        dist.barrier()
        ms = (time.perf_counter() - t0) / steps * 1e3
    if rank == 0:
        print(f"world {world}: {per_rank} structures/rank x 108 atoms, {ms:.3f} ms per step, "
              f"{per_rank * world / ms * 1e3:.0f} structures/s", flush=True)
    flag = torch.tensor([1 if ok else 0], device='cuda')
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == '__main__':
    main()
