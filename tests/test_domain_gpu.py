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
