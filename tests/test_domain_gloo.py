"""World-size-2 (and 3) gloo test of the slab-decomposition plumbing on CPU:
layout, sender-side periodic shifts and the ring-exchange ordering of
tensoralloy_b200.domain.DistComm.  Each rank assembles owned + halo positions and
checks with the oracle that every owned atom sees exactly the neighbours it has
in the global periodic structure (count and sorted distances)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from oracle import neighbor as onl
        from tensoralloy_b200.atoms import fcc_positions
        from tensoralloy_b200.domain import DistComm, SlabLayout
        rc = 6.5
        pos, cell = fcc_positions(3.52, 4 * world, 3, 3)
        rng = np.random.default_rng(611)
        pos = pos + rng.normal(scale=0.05, size=pos.shape)
        lx = cell[0, 0]
        pos[:, 0] = np.mod(pos[:, 0], lx)
        lay = SlabLayout(lx, world, rank, rc)
        own_idx = np.flatnonzero(lay.owned_mask(pos[:, 0]))
        owned = pos[own_idx]
        m_l, m_r = lay.send_masks(owned[:, 0])
        comm = DistComm(lay)
        n_l, n_r = comm.exchange_counts(int(m_l.sum()), int(m_r.sum()), 'cpu')
        s_l = torch.from_numpy(owned[m_l] + np.array([lay.shift_to_left, 0, 0]))
        s_r = torch.from_numpy(owned[m_r] + np.array([lay.shift_to_right, 0, 0]))
        r_l = torch.zeros((n_l, 3), dtype=torch.float64)
        r_r = torch.zeros((n_r, 3), dtype=torch.float64)
        comm.exchange(s_l.contiguous(), s_r.contiguous(), r_l, r_r)
        local = np.concatenate([owned, r_l.numpy(), r_r.numpy()])
        # halo atoms must lie just outside the slab on the proper side
        assert np.all(r_l.numpy()[:, 0] < lay.lo) and np.all(r_l.numpy()[:, 0] >= lay.lo - rc)
        assert np.all(r_r.numpy()[:, 0] >= lay.hi) and np.all(r_r.numpy()[:, 0] < lay.hi + rc)
        fcell, origin, pbc = lay.frame(cell[1, 1], cell[2, 2])
        li, lj, lS, ld, _ = onl.neighbor_list(local - origin, fcell, pbc, rc)
        gi, gj, gS, gd, _ = onl.neighbor_list(pos, cell, [1, 1, 1], rc)
        n_own = len(owned)
        ok = True
        for k, g in enumerate(own_idx):
            a = np.sort(ld[li == k])
            b = np.sort(gd[gi == g])
            if len(a) != len(b) or np.abs(a - b).max() > 1e-12:
                ok = False
                break
        t = torch.tensor([float(n_own)], dtype=torch.float64)
        comm.allreduce_sum(t)
        q.put((rank, ok, int(t.item()) == len(pos)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_ring_exchange_gloo(world):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, ok, total_ok in results:
        assert ok, f"rank {rank}: local neighbourhood differs from the global one"
        assert total_ok, "owned atoms do not partition the structure"


def _grad_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from tensoralloy_b200.nn.atomic.training import allreduce_mean_
        a = torch.zeros(3, 2, dtype=torch.float64, requires_grad=True)
        b = torch.zeros(4, dtype=torch.float64, requires_grad=True)
        a.grad = torch.full((3, 2), float(rank + 1), dtype=torch.float64)
        b.grad = None if rank == 0 else torch.arange(4, dtype=torch.float64)
        allreduce_mean_([a, b], dist, world)
        q.put((rank, a.grad.tolist(), b.grad.tolist()))
    finally:
        dist.destroy_process_group()


def test_flat_gradient_allreduce_mean_gloo():
    """Structure-parallel training: one flat all-reduce, MEAN aggregation
    (train/distribute_utils.py:84-159; potentials.py:41)."""
    world = 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_grad_worker, args=(r, world, port, q))
             for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, ga, gb in results:
        assert ga == [[1.5, 1.5]] * 3
        assert gb == [0.0, 0.5, 1.0, 1.5]


def _migrate_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from tensoralloy_b200.atoms import fcc_positions
        from tensoralloy_b200.domain import DistComm, SlabLayout
        rc, skin = 6.5, 0.4
        pos, cell = fcc_positions(3.52, 4 * world, 3, 3)
        rng = np.random.default_rng(611)
        pos = pos + rng.normal(scale=0.05, size=pos.shape)
        lx = cell[0, 0]
        pos[:, 0] = np.mod(pos[:, 0], lx)
        lay = SlabLayout(lx, world, rank, rc, skin)
        ids = np.flatnonzero(lay.owned_mask(pos[:, 0]))
        comm = DistComm(lay)
        state = torch.from_numpy(np.concatenate(
            [pos[ids], np.zeros((len(ids), 3)), ids[:, None].astype(np.float64)], axis=1))
        ok = True
        for cycle in range(4):
            # every atom drifts to the right by up to 1.5 A (same field on all ranks, by id):
            # some leave through the high face, rank world - 1 wraps them to rank 0
            drift = np.random.default_rng(cycle).uniform(0.0, 1.5, size=len(pos))
            state[:, 0] += torch.from_numpy(drift[state[:, 6].numpy().astype(np.int64)])
            state = comm.migrate(state)
            x = state[:, 0].numpy()
            ok = ok and bool(np.all(lay.owned_mask(x)) and np.all(x >= lay.lo - 1e-12))
            n = torch.tensor([float(state.shape[0])], dtype=torch.float64)
            comm.allreduce_sum(n)
            ok = ok and int(n.item()) == len(pos)
        # ids: every atom exactly once over the ranks
        mine = torch.zeros(len(pos), dtype=torch.float64)
        mine[state[:, 6].numpy().astype(np.int64)] += 1.0
        comm.allreduce_sum(mine)
        q.put((rank, ok, bool((mine == 1.0).all())))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_slab_migration_gloo(world):
    """Rebuild-time migration of owned atoms between slabs (domain.DistComm.migrate): after
    every cycle each rank holds exactly the atoms of its slab, nothing is lost or duplicated,
    atoms leaving through the periodic boundary arrive at the other end of the ring."""
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_migrate_worker, args=(r, world, port, q))
             for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, ok, once in results:
        assert ok, f"rank {rank}: atoms outside the slab / atom count changed"
        assert once, "an atom is owned by no rank or by two"
