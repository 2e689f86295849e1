"""Multi-GPU check of the slab decomposition (run under torchrun, one rank per GPU):
peer-memory halo exchange + CUDA-graph step against a single-GPU evaluation of the
same structure.  Exit code 0 = every rank agrees to 1e-10 eV/atom, 1e-8 eV/A."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tensoralloy_b200 import _lib                                  # noqa: E402
from tensoralloy_b200.atoms import fcc_positions                   # noqa: E402
from tensoralloy_b200.domain import SlabDomain                     # noqa: E402
from tensoralloy_b200.nn.eam.potentials import get_potential       # noqa: E402


def main():
    world = int(os.environ['WORLD_SIZE'])
    rank = int(os.environ['RANK'])
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    cells = int(sys.argv[1]) if len(sys.argv) > 1 else 4 * world
    a, rc, sigma, seed = 3.52, 6.5, 0.05, 611
    pot = get_potential('zjw04')
    model = _lib.EamModel(_lib.EAM_ALLOY, 1, [pot.rho('Ni')], [pot.phi('NiNi')],
                          [pot.embed('Ni')])
    dom = SlabDomain(model, cells, a, rc, sigma, seed, world, rank, scaling='strong')
    # single-GPU answer on this rank
    pos, cell = fcc_positions(a, cells, cells, cells)
    pos = pos + np.random.default_rng(seed).normal(scale=sigma, size=pos.shape)
    pos[:, 0] = np.mod(pos[:, 0], cell[0, 0])
    nl = _lib.NeighborList()
    d_pos = torch.tensor(pos, device='cuda')
    nl.build(d_pos, None, cell, [1, 1, 1], rc)
    e = torch.zeros(1, dtype=torch.float64, device='cuda')
    f = torch.zeros((len(pos), 3), dtype=torch.float64, device='cuda')
    v = torch.zeros(9, dtype=torch.float64, device='cuda')
    model.eval(nl, 0, energy=e, forces=f, virial=v)
    own = np.flatnonzero(dom.layout.owned_mask(pos[:, 0]))
    f_ref = f.cpu().numpy()[own]
    ok = True
    modes = ['eager']
    for mode in ('eager', 'graph', 'e2e'):
        if mode == 'graph' and dom.peer is None:
            print(f"rank {rank}: NCCL fallback path runs eagerly (no graph)", flush=True)
            continue
        if mode == 'graph' and not dom.enable_graph():
            print(f"rank {rank}: graph capture failed: {getattr(dom, 'graph_error', '?')}")
            ok = False
            continue
        for _ in range(3):
            dom.step_e2e() if mode == 'e2e' else dom.step()
        torch.cuda.synchronize()
        E, F, V = dom.results()
        de = abs(E - e.item()) / len(pos)
        df = np.abs(F - f_ref).max()
        dv = np.abs(V.reshape(-1) - v.cpu().numpy()).max() / len(pos)
        good = de < 1e-10 and df < 1e-8 and dv < 1e-8
        ok = ok and good
        print(f"rank {rank} {mode}: peer={dom.peer is not None} dE/N={de:.2e} dF={df:.2e} "
              f"dV/N={dv:.2e} {'OK' if good else 'FAIL'}", flush=True)
    # MD cycle: moving atoms, lists with a skin, collective rebuild decision, migration of the
    # atoms that leave their slab, CUDA-graph step re-captured after every rebuild; after 25
    # steps the decomposed forces equal a single-GPU evaluation of the SAME positions
    skin = 0.3
    md = SlabDomain(model, cells, a, rc, sigma, seed, world, rank, scaling='strong', skin=skin)
    md.vstep_max *= 4.0
    md.state[:, 3:6] *= 4.0          # faster atoms: more than one rebuild, atoms cross faces
    md.d_vel = md.state[:, 3:6].contiguous()
    if md.peer is not None:
        md.enable_graph()
    n0 = md.rank_state.n_owned
    for _ in range(25):
        md.md_step()
    torch.cuda.synchronize()
    pos_all, f_all = md.gather_global()
    tot = md._totals().cpu().numpy()
    nl2 = _lib.NeighborList()
    d_all = torch.tensor(pos_all, device='cuda')
    nl2.build(d_all, None, cell, [1, 1, 1], rc)
    model.eval(nl2, 0, energy=e, forces=f, virial=v)
    de = abs(tot[0] - e.item()) / len(pos)
    df = np.abs(f_all - f.cpu().numpy()).max()
    dv = np.abs(tot[1:10] - v.cpu().numpy()).max() / len(pos)
    good = de < 1e-10 and df < 1e-8 and dv < 1e-8 and md.rebuilds >= 2
    ok = ok and good
    print(f"rank {rank} md-cycle: rebuilds={md.rebuilds} owned {n0}->{md.rank_state.n_owned} "
          f"graph={md.graph is not None} dE/N={de:.2e} dF={df:.2e} dV/N={dv:.2e} "
          f"{'OK' if good else 'FAIL'}", flush=True)
    if dom.peer is None:
        print(f"rank {rank}: peer path unavailable: {getattr(dom, 'peer_error', '?')}")
    flag = torch.tensor([1 if ok else 0], device='cuda')
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == '__main__':
    main()
