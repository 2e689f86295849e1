"""GPU neighbour lists vs the ASE restatement (oracle/neighbor.py): bit-exact
after canonical sort (i, j, Sx, Sy, Sz).  Reference call sites:
transformer/universal.py:58, neighbor.py:84."""
import numpy as np
import pytest
import torch

from oracle import neighbor as onl
from tensoralloy_b200 import _lib
from tensoralloy_b200.atoms import bulk_fcc, fcc_positions

pytestmark = pytest.mark.gpu


def gpu_list(pos, cell, pbc, rc, types=None):
    nl = _lib.NeighborList()
    d_pos = torch.tensor(pos, dtype=torch.float64, device='cuda').contiguous()
    d_t = None if types is None else torch.tensor(types, dtype=torch.int32,
                                                  device='cuda')
    nl.build(d_pos, d_t, cell, pbc, rc)
    i, j, S = nl.export()
    torch.cuda.synchronize()
    return nl, i.cpu().numpy(), j.cpu().numpy(), S.cpu().numpy()


def assert_same_list(pos, cell, pbc, rc, brute=False):
    fn = onl.neighbor_list_brute if brute else onl.neighbor_list
    ri, rj, rS, _, _ = fn(pos, cell, pbc, rc)
    nl, gi, gj, gS = gpu_list(pos, cell, pbc, rc)
    nij, nnl, _ = nl.sizes()
    assert nij == len(ri)
    gi, gj, gS = onl.canonical_sort(gi, gj, gS)
    np.testing.assert_array_equal(gi, ri)
    np.testing.assert_array_equal(gj, rj)
    np.testing.assert_array_equal(gS, rS)
    assert nnl == np.bincount(ri, minlength=len(pos)).max()
    counts = nl.counts().cpu().numpy()
    np.testing.assert_array_equal(counts, np.bincount(ri, minlength=len(pos)))


def test_ni_fcc_256_perfect_and_rattled():
    atoms = bulk_fcc('Ni', 3.52, (4, 4, 4))
    assert_same_list(atoms.positions, atoms.cell, [1, 1, 1], 6.5)
    nl, gi, _, _ = gpu_list(atoms.positions, atoms.cell, [1, 1, 1], 6.5)
    assert nl.sizes()[0] == 22016      # 86 neighbours x 256
    rng = np.random.default_rng(611)
    pos = atoms.positions + rng.normal(scale=0.05, size=atoms.positions.shape)
    assert_same_list(pos, atoms.cell, [1, 1, 1], 6.5)
    assert_same_list(pos, atoms.cell, [1, 1, 1], 6.0)


def test_small_cell_multiple_images():
    # Ni 2x2x2 cubic (7.04 A) with rc 6.5 > L/2: same atom through several images
    atoms = bulk_fcc('Ni', 3.52, (2, 2, 2))
    assert_same_list(atoms.positions, atoms.cell, [1, 1, 1], 6.5, brute=True)
    # single conventional cell: rc spans two images
    atoms = bulk_fcc('Ni', 3.52, (1, 1, 1))
    assert_same_list(atoms.positions, atoms.cell, [1, 1, 1], 6.5, brute=True)


def test_triclinic_and_mixed_pbc():
    rng = np.random.default_rng(3)
    cell = np.array([[9.0, 0.0, 0.0], [2.5, 8.0, 0.0], [1.0, -1.5, 10.0]])
    pos = rng.random((120, 3)) @ cell
    for pbc in ([1, 1, 1], [1, 1, 0], [1, 0, 0], [0, 0, 0]):
        assert_same_list(pos, cell, pbc, 5.0, brute=True)


def test_unwrapped_positions():
    # atoms far outside the cell: S is relative to the positions as given
    rng = np.random.default_rng(5)
    atoms = bulk_fcc('Ni', 3.52, (3, 3, 3))
    pos = atoms.positions + rng.normal(scale=0.05, size=atoms.positions.shape)
    pos += rng.integers(-2, 3, size=pos.shape) @ atoms.cell
    assert_same_list(pos, atoms.cell, [1, 1, 1], 6.5)


def test_large_lattice_counts():
    pos, cell = fcc_positions(3.52, 20, 20, 20)     # 32 000 atoms
    rng = np.random.default_rng(611)
    pos = pos + rng.normal(scale=0.05, size=pos.shape)
    assert_same_list(pos, cell, [1, 1, 1], 6.5)
