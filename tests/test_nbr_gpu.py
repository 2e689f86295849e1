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


# --- block-per-tile build (large systems) ------------------------------------
# The tile kernel must produce the SAME rows, in the same order, as the
# thread-per-atom kernels; TAB_NBR_MODE selects the kernel for A/B runs.
def _export_mode(mode, pos, cell, pbc, rc, types=None, monkeypatch=None):
    monkeypatch.setenv('TAB_NBR_MODE', mode)
    nl, i, j, S = gpu_list(pos, cell, pbc, rc, types)
    return nl.sizes(), nl.counts().cpu().numpy(), i, j, S


def _assert_modes_equal(pos, cell, pbc, rc, types, monkeypatch):
    a = _export_mode('thread', pos, cell, pbc, rc, types, monkeypatch)
    b = _export_mode('tile', pos, cell, pbc, rc, types, monkeypatch)
    assert a[0] == b[0]
    for u, v in zip(a[1:], b[1:]):
        np.testing.assert_array_equal(u, v)      # same entries, same order


def test_tile_kernel_matches_thread_kernel_fcc(monkeypatch):
    pos, cell = fcc_positions(3.52, 21, 20, 19)      # odd bin counts: partial tiles
    rng = np.random.default_rng(611)
    pos = pos + rng.normal(scale=0.05, size=pos.shape)
    _assert_modes_equal(pos, cell, [1, 1, 1], 6.5, None, monkeypatch)
    # and against ASE's list
    monkeypatch.setenv('TAB_NBR_MODE', 'tile')
    assert_same_list(pos, cell, [1, 1, 1], 6.5)


def test_tile_kernel_species_triclinic_mixed_pbc(monkeypatch):
    rng = np.random.default_rng(9)
    cell = np.array([[60.0, 0.0, 0.0], [11.0, 55.0, 0.0], [4.0, -7.0, 48.0]])
    n = 14000
    pos = rng.random((n, 3)) @ cell
    types = rng.integers(0, 3, size=n)
    for pbc in ([1, 1, 1], [1, 0, 1], [0, 0, 0]):
        _assert_modes_equal(pos, cell, pbc, 5.0, types, monkeypatch)
    monkeypatch.setenv('TAB_NBR_MODE', 'tile')
    assert_same_list(pos[:3000] * 0.5, cell * 0.5, [1, 1, 0], 5.0)


def test_tile_kernel_dense_cells_and_thin_box(monkeypatch):
    # > NBT_CAP candidates per cell (segments are split) and > 256 atoms per tile
    rng = np.random.default_rng(2)
    cell = np.diag([16.0, 16.0, 16.0])
    pos = rng.random((9000, 3)) @ cell             # 2.2 atoms / A^3
    _assert_modes_equal(pos, cell, [1, 1, 1], 4.0, None, monkeypatch)
    # a box thinner than the cutoff in y: search range > 1 bin along y
    cell = np.diag([120.0, 5.0, 40.0])
    pos = rng.random((3000, 3)) @ cell
    _assert_modes_equal(pos, cell, [1, 1, 1], 6.0, rng.integers(0, 2, size=3000),
                        monkeypatch)


def test_tile_kernel_boundary_distances(monkeypatch):
    # perfect lattice with rc exactly on a shell: the exact re-test decides
    pos, cell = fcc_positions(3.52, 12, 12, 12)
    rc = 3.52 * np.sqrt(2.0)                        # 4th shell distance
    _assert_modes_equal(pos, cell, [1, 1, 1], rc, None, monkeypatch)
    monkeypatch.setenv('TAB_NBR_MODE', 'tile')
    assert_same_list(pos, cell, [1, 1, 1], rc)


def test_half_width_cells_on_odd_geometries(monkeypatch):
    """TAB_NBR_SUBDIV=2 forces the large-system grid (cells >= rc / 2, tiles of 4 x 4 x 4 cells,
    125-cell neighbourhoods walked row by row) onto the small odd cases above: several species,
    triclinic cells, mixed / no periodicity, > 2048 candidates per box (rows split over
    candidate windows), a box thinner than the cutoff, rc exactly on a shell.  Same rows in the
    same order as the thread-per-atom kernels on the same grid, and ASE's list."""
    monkeypatch.setenv('TAB_NBR_SUBDIV', '2')
    rng = np.random.default_rng(9)
    cell = np.array([[60.0, 0.0, 0.0], [11.0, 55.0, 0.0], [4.0, -7.0, 48.0]])
    n = 14000
    pos = rng.random((n, 3)) @ cell
    types = rng.integers(0, 3, size=n)
    for pbc in ([1, 1, 1], [1, 0, 1], [0, 0, 0]):
        _assert_modes_equal(pos, cell, pbc, 5.0, types, monkeypatch)
    monkeypatch.setenv('TAB_NBR_MODE', 'tile')
    assert_same_list(pos[:3000] * 0.5, cell * 0.5, [1, 1, 0], 5.0)
    assert_same_list(pos[:2500] * 0.4, cell * 0.4, [0, 1, 0], 5.0)
    # dense: 2.2 atoms / A^3, ~9000 candidates per box
    cell = np.diag([16.0, 16.0, 16.0])
    pos = rng.random((9000, 3)) @ cell
    _assert_modes_equal(pos, cell, [1, 1, 1], 4.0, None, monkeypatch)
    monkeypatch.setenv('TAB_NBR_MODE', 'tile')
    assert_same_list(pos[:1500], cell, [1, 1, 1], 4.0)
    # thin box, two species
    cell = np.diag([120.0, 5.0, 40.0])
    pos = rng.random((3000, 3)) @ cell
    _assert_modes_equal(pos, cell, [1, 1, 1], 6.0, rng.integers(0, 2, size=3000), monkeypatch)
    monkeypatch.setenv('TAB_NBR_MODE', 'tile')
    assert_same_list(pos, cell, [1, 1, 1], 6.0)
    # rc exactly on a shell of the perfect lattice
    pos, cell = fcc_positions(3.52, 12, 12, 12)
    rc = 3.52 * np.sqrt(2.0)
    _assert_modes_equal(pos, cell, [1, 1, 1], rc, None, monkeypatch)
    monkeypatch.setenv('TAB_NBR_MODE', 'tile')
    assert_same_list(pos, cell, [1, 1, 1], rc)
