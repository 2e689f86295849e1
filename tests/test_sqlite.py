"""`CoreDatabase` (io/sqlite.py): ASE SQLite files read and written without ASE, against the
reference's own databases -- `qm7m.db` (whole) and the structures of `snap-Ni.db` that attain
the neighbour maxima the reference recorded in the file's metadata
(tests/golden/make_golden.py:copy_databases)."""
import os
import shutil

import numpy as np
import pytest

from oracle import neighbor as onl
from tensoralloy_b200.io.sqlite import CoreDatabase, _get_keypath
from tensoralloy_b200.neighbor import NeighborProperty, NeighborSize

GOLD = os.path.join(os.path.dirname(__file__), 'golden')


def _copy(name, tmp_path):
    dst = str(tmp_path / name)
    shutil.copy(os.path.join(GOLD, name), dst)
    os.chmod(dst, 0o644)
    return dst


def _oracle_sizes(atoms, rc):
    """neighbor.py:50-146 on the oracle's list."""
    i, j, _ = onl.neighbor_list(atoms.positions, atoms.cell, atoms.pbc, rc)[:3]
    numbers = atoms.numbers
    species = sorted(set(numbers.tolist()))
    t = np.searchsorted(species, numbers)
    tc = np.zeros((len(atoms), len(species)), dtype=np.int64)
    np.add.at(tc, (i, t[j]), 1)
    cnt = tc.sum(axis=1)
    ij2k = 0
    for tj in range(len(species)):
        rows = tc[:, tj] > 0
        if rows.any():
            other = tc[rows].copy()
            other[:, tj] -= 1
            ij2k = max(ij2k, int(other.max()))
    return NeighborSize(nij=len(i), nnl=int(tc.max()) if len(i) else 0,
                        nijk=int((cnt * (cnt - 1) // 2).sum()), ij2k=ij2k)


def test_qm7m_database(tmp_path):
    db = CoreDatabase(_copy('qm7m.db', tmp_path))
    assert len(db) == 3
    assert db.max_occurs == {'C': 5, 'H': 8, 'O': 2}
    assert db.has_forces and not db.has_stress and not db.has_periodic_structures
    assert db.get_nij_max(6.0) == 198 and db.get_nnl_max(6.0) == 8
    assert db.get_nijk_max(6.0) == 1217 and db.get_nijk_max(6.0, symmetric=False) == 2434
    assert db.get_nij_max(5.0) is None
    ase = db.get_atomic_static_energy()
    assert ase['C'] == pytest.approx(-0.27244210000000013)
    # the reference's tests/test_neighbor.py:20-36 on the file it reads
    ch4 = db.get_atoms('id=2')
    assert ch4.get_chemical_formula() == 'CH4' and not ch4.pbc.any()
    assert ch4.info['energy'] == -0.66606164 and ch4.info['forces'].shape == (5, 3)
    size = _oracle_sizes(ch4, 6.5)
    assert (size.nij, size.nnl) == (20, 4)
    c2h6 = db.get_atoms(id=3, add_additional_information=True)
    size = _oracle_sizes(c2h6, 6.5)
    assert (size.nij, size.nijk, size.nnl) == (56, 168, 6)
    assert c2h6.info['data'] == {'weights': [1.0, 1.0, 1.0]}
    with pytest.raises(KeyError):
        db.get_atoms(id=4)


def test_snap_subset_reproduces_recorded_maxima(tmp_path):
    db = CoreDatabase(_copy('snap_Ni_subset.db', tmp_path))
    assert len(db) == 7 and db.has_stress and db.has_periodic_structures
    md = db.metadata
    assert md['unit_conversion']['stress'] == pytest.approx(0.0006241509125883258)
    images = [db.get_atoms(id=k) for k in range(1, 8)]
    assert all(a.info['stress'].shape == (6,) for a in images)
    for rc, k_max, props in ((6.5, 2, ('nij', 'nnl')), (6.0, 2, ('nij', 'nnl')),
                             (4.6, 3, ('nij', 'nnl', 'nijk')),
                             (4.5, 3, ('nij', 'nnl', 'nijk', 'ij2k'))):
        sizes = [_oracle_sizes(a, rc) for a in images]
        for p in props:
            recorded = md['neighbors'][str(k_max)][f'{rc * 100:.0f}'][f'{p}_max']
            assert max(s[p] for s in sizes) == recorded, (rc, p)
    assert db.get_nij_max(6.5) == 14494 and db.get_nnl_max(6.5) == 136
    assert db.get_ij2k_max(4.5) == 42 and db.get_nijk_max(4.5) == 93744


def test_write_read_round_trip(tmp_path):
    src = CoreDatabase(_copy('snap_Ni_subset.db', tmp_path))
    path = str(tmp_path / 'new.db')
    db = CoreDatabase(path)
    ids = [db.write(src.get_atoms(id=k)) for k in (2, 5)]
    assert ids == [1, 2] and len(db) == 2
    db.metadata = {'forces': True, 'stress': True, 'periodic': True}
    assert db.max_occurs == {'Ni': max(len(src.get_atoms(id=k)) for k in (2, 5))}
    db.close()
    again = CoreDatabase(path)
    a, b = again.get_atoms(id=2), src.get_atoms(id=5)
    assert np.array_equal(a.positions, b.positions) and np.array_equal(a.cell, b.cell)
    assert np.array_equal(a.pbc, b.pbc) and np.array_equal(a.numbers, b.numbers)
    assert a.info['energy'] == b.info['energy']
    assert np.array_equal(a.info['forces'], b.info['forces'])
    assert np.array_equal(a.info['stress'], b.info['stress'])
    assert again.metadata['max_occurs'] == {'Ni': len(a) if len(a) > len(again.get_atoms(id=1))
                                            else len(again.get_atoms(id=1))}
    # may_update_neighbor_meta: sqlite.py:300-323
    new = NeighborSize(nij=10, nnl=3, nijk=0, ij2k=0)
    assert again.may_update_neighbor_meta(4.0, new)
    assert not again.may_update_neighbor_meta(4.0, new)
    assert again.may_update_neighbor_meta(4.0, NeighborSize(nij=12, nnl=2, nijk=0, ij2k=0))
    assert again.get_nij_max(4.0) == 12 and again.get_nnl_max(4.0) == 3
    assert _get_keypath(3, 4.55, NeighborProperty.nijk) == 'neighbors.3.455.nijk_max'
    ds = again.to_dataset()
    assert len(ds) == 2 and ds.has_stress and ds.labels(0)[1].shape == (len(ds[0]), 3)


@pytest.mark.gpu
def test_update_neighbor_meta_on_gpu(tmp_path):
    """The GPU list reproduces the maxima the reference stored in snap-Ni.db."""
    db = CoreDatabase(_copy('snap_Ni_subset.db', tmp_path))
    db.metadata = {k: v for k, v in db.metadata.items() if k != 'neighbors'}
    assert db.get_nij_max(6.5) is None
    assert db.get_nij_max(6.5, allow_calculation=True) == 14494
    assert db.get_nnl_max(6.5) == 136
    size = db.update_neighbor_meta(4.5, nijk=True, ij2k=True)
    assert (size.nij, size.nnl, size.nijk, size.ij2k) == (4554, 43, 93744, 42)
    assert db.metadata['neighbors']['2']['450']['nijk_max'] == 0
    from tensoralloy_b200.neighbor import find_neighbor_size_of_atoms
    q = CoreDatabase(_copy('qm7m.db', tmp_path))
    s = find_neighbor_size_of_atoms(q.get_atoms(id=3), 6.5, find_nijk=True)
    assert (s.nij, s.nijk, s.nnl) == (56, 168, 6)
