"""Slab decomposition on ONE GPU: every rank of a 2- and 3-way
decomposition runs in-process (domain.run_loopback); the summed energy / virial
and the owned-atom forces must equal the single-domain result and the oracle."""
import numpy as np
import pytest
import torch

from oracle import eam as oeam
from oracle import potentials as opot
from tensoralloy_b200 import _lib
from tensoralloy_b200.atoms import fcc_positions
from tensoralloy_b200.domain import run_loopback
from tensoralloy_b200.nn.eam.potentials import get_potential

pytestmark = pytest.mark.gpu


def _model():
    pot = get_potential('zjw04')
    return _lib.EamModel(_lib.EAM_ALLOY, 1, [pot.rho('Ni')], [pot.phi('NiNi')],
                         [pot.embed('Ni')])


@pytest.mark.parametrize("world", [2, 3])
def test_slabs_match_single_domain(world):
    pos, cell = fcc_positions(3.52, 12, 5, 5)
    rng = np.random.default_rng(611)
    pos = pos + rng.normal(scale=0.05, size=pos.shape)
    model = _model()
    e, f, v = run_loopback(model, pos, cell, 6.5, world)
    ref = oeam.eam_evaluate(opot.get_potential('zjw04'), 'alloy', ['Ni'],
                            ['Ni'] * len(pos), pos, cell, [1, 1, 1], 6.5)
    n = len(pos)
    assert abs(e - ref['energy']) / n < 1e-10
    assert np.abs(f - ref['forces']).max() < 1e-8
    assert np.abs(v - ref['virial']).max() / n < 1e-8


def test_update_without_rebuild():
    """Lists with a skin reused after a small move: the result equals the oracle on a FRESH
    list of the moved positions (the reference rebuilds per call, universal.py:58)."""
    pos, cell = fcc_positions(3.52, 8, 4, 4)
    rng = np.random.default_rng(3)
    pos = pos + rng.normal(scale=0.05, size=pos.shape)
    model = _model()
    nl = _lib.NeighborList()
    nl.set_skin(0.3)
    d_pos = torch.tensor(pos, device='cuda')
    nl.build(d_pos, None, cell, [1, 1, 1], 6.5)
    pos2 = pos + rng.normal(scale=0.01, size=pos.shape)
    nl.update(torch.tensor(pos2, device='cuda'))
    disp, skin = nl.max_displacement()
    assert skin == 0.3 and 0.0 < disp < 0.15
    e = torch.zeros(1, dtype=torch.float64, device='cuda')
    f = torch.zeros((len(pos), 3), dtype=torch.float64, device='cuda')
    v = torch.zeros(9, dtype=torch.float64, device='cuda')
    model.eval(nl, 0, energy=e, forces=f, virial=v)
    ref = oeam.eam_evaluate(opot.get_potential('zjw04'), 'alloy', ['Ni'],
                            ['Ni'] * len(pos), pos2, cell, [1, 1, 1], 6.5)
    assert abs(e.item() - ref['energy']) / len(pos) < 1e-10
    assert np.abs(f.cpu().numpy() - ref['forces']).max() < 1e-8
    assert np.abs(v.cpu().numpy().reshape(3, 3) - ref['virial']).max() / len(pos) < 1e-8


@pytest.mark.parametrize("world", [2, 3])
def test_slabs_with_skin_reused_after_moves(world):
    """Decomposed run on lists with a skin: after two moves below skin / 2 the reused lists
    (halo = rc + skin) give the oracle's result on the moved structure."""
    pos, cell = fcc_positions(3.52, 12, 5, 5)
    rng = np.random.default_rng(611)
    pos = pos + rng.normal(scale=0.05, size=pos.shape)
    model = _model()
    moves = [rng.uniform(-0.04, 0.04, size=pos.shape) for _ in range(2)]
    e, f, v = run_loopback(model, pos, cell, 6.5, world, skin=0.4, moves=moves)
    pos2 = pos + moves[0] + moves[1]
    ref = oeam.eam_evaluate(opot.get_potential('zjw04'), 'alloy', ['Ni'],
                            ['Ni'] * len(pos), pos2, cell, [1, 1, 1], 6.5)
    n = len(pos)
    assert abs(e - ref['energy']) / n < 1e-10
    assert np.abs(f - ref['forces']).max() < 1e-8
    assert np.abs(v - ref['virial']).max() / n < 1e-8


def test_slabs_with_tile_build(monkeypatch):
    # the same decomposition with the block-per-tile list build (halo atoms are the
    # second group of every cell)
    monkeypatch.setenv('TAB_NBR_MODE', 'tile')
    test_slabs_match_single_domain(2)


def test_partition_and_send_set_kernels_match_the_host_formulation():
    """tab_dd_partition / tab_dd_send_sets (csrc/dd.cu) against the torch formulation of
    `DistComm.migrate` and `SlabLayout.send_masks`: same rows, same (stable) order."""
    import torch
    from tensoralloy_b200 import _lib
    from tensoralloy_b200.domain import SlabLayout
    rng = np.random.default_rng(17)
    for world, rank, n in ((4, 1, 10007), (4, 0, 5000), (4, 3, 777), (2, 1, 4096), (8, 5, 1)):
        lx, rc, skin = 80.0, 4.0, 0.5
        lay = SlabLayout(lx, world, rank, rc, skin)
        # atoms of this slab, some of which drifted into the adjacent slabs (and across the
        # periodic boundary)
        x = rng.uniform(lay.lo - 0.6 * lay.width, lay.hi + 0.6 * lay.width, n)
        special = np.array([lx, 0.0, -0.0, lay.lo, lay.hi, lay.lo + 2.5 * lay.width])
        x[: n // 50] = special[rng.integers(0, 6, n // 50)]
        state = np.concatenate((x[:, None], rng.normal(size=(n, 6))), axis=1)
        d_state = torch.tensor(state, device='cuda')
        cap = 4096
        keep = torch.full((n + 8, 7), np.nan, dtype=torch.float64, device='cuda')
        mail_l = torch.full((1 + cap * 7,), np.nan, dtype=torch.float64, device='cuda')
        mail_r = torch.full((1 + cap * 7,), np.nan, dtype=torch.float64, device='cuda')
        counts = torch.zeros(8, dtype=torch.int32, device='cuda')
        work = torch.zeros(3 * ((n + 255) // 256) + 8, dtype=torch.int32, device='cuda')
        _lib.dd_partition(d_state, lx, lay.width, world, rank, keep, mail_l, mail_r, cap, counts,
                          work)
        torch.cuda.synchronize()
        # host formulation (domain.py: DistComm.migrate)
        xs = torch.remainder(d_state[:, 0], lx)
        xs = torch.where(xs >= lx, torch.zeros_like(xs), xs)
        ref = d_state.clone()
        ref[:, 0] = xs
        owner = lay.owner_of(xs).to(torch.int64)
        is_keep = owner == rank
        to_l, to_r = owner == lay.left, owner == lay.right
        if world == 2:
            to_r, to_l = ~is_keep, torch.zeros_like(is_keep)
        lost = ~(is_keep | to_l | to_r)
        c = counts.tolist()
        assert c[:4] == [int(is_keep.sum()), int(to_l.sum()), int(to_r.sum()), int(lost.sum())]
        assert c[4] == 0
        assert torch.equal(keep[:c[0]], ref[is_keep])
        assert mail_l[0].item() == c[1] and mail_r[0].item() == c[2]
        assert torch.equal(mail_l[1:1 + 7 * c[1]].view(-1, 7), ref[to_l])
        assert torch.equal(mail_r[1:1 + 7 * c[2]].view(-1, 7), ref[to_r])
        # send sets of the kept atoms
        pos = ref[is_keep][:, 0:3].contiguous()
        m = pos.shape[0]
        idx = torch.full((2, m + 8), -1, dtype=torch.int64, device='cuda')
        sc = torch.zeros(2, dtype=torch.int32, device='cuda')
        _lib.dd_send_sets(pos, lay.lo + lay.reach, lay.hi - lay.reach, idx[0], idx[1], sc, work)
        m_l, m_r = lay.send_masks(pos[:, 0])
        n_l, n_r = sc.tolist()
        assert torch.equal(idx[0, :n_l], torch.nonzero(m_l).flatten())
        assert torch.equal(idx[1, :n_r], torch.nonzero(m_r).flatten())
    # a mailbox that is too small is reported, not overrun
    counts.zero_()
    small = torch.full((1 + 2 * 7,), np.nan, dtype=torch.float64, device='cuda')
    big = torch.tensor(np.concatenate((np.full((50, 1), 1.0), np.zeros((50, 6))), axis=1),
                       device='cuda')
    keep = torch.zeros((64, 7), dtype=torch.float64, device='cuda')
    _lib.dd_partition(big, 80.0, 20.0, 4, 1, keep, small, small.clone(), 2, counts, work)
    assert counts.tolist()[:5] == [0, 50, 0, 0, 1]
