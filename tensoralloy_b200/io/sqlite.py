"""
`CoreDatabase`: the reference's dataset store (tensoralloy/io/sqlite.py:35-324), an ASE SQLite3
database (`ase.db.sqlite`, schema version 8/9) with a JSON `metadata` record in the `information`
table.  ASE is not installed here, so the file format is read and written directly with the
standard library's sqlite3:

* table `systems`: one row per structure; arrays are little-endian BLOBs (`numbers` int32,
  `positions` / `cell` / `forces` / `stress` float64), `pbc` is the bit mask
  pbc[0] + 2 pbc[1] + 4 pbc[2], `key_value_pairs` and `data` are JSON texts;
* table `information`: (name, value) rows, 'version' and 'metadata'.

Mirrored interface: `len`, `get_atoms(id=...)` / `get_atoms('id=3')`, `metadata` (get / set,
written back), `max_occurs`, `has_forces`, `has_stress`, `has_periodic_structures`,
`get_atomic_static_energy`, `get_nij_max / get_nijk_max / get_nnl_max / get_ij2k_max`,
`update_neighbor_meta`, `may_update_neighbor_meta`; plus `write(atoms)` (the reference fills a
database through ASE in `io/read.py`) and `to_dataset()` for the trainers of this package.
Neighbour sizes come from the GPU list (`tensoralloy_b200.neighbor`).
"""
import json
import os
import sqlite3
import time
from collections import Counter

import numpy as np

from tensoralloy_b200.atoms import Atoms
from tensoralloy_b200.neighbor import NeighborProperty, NeighborSize

_SYSTEMS = """CREATE TABLE IF NOT EXISTS systems (
    id INTEGER PRIMARY KEY AUTOINCREMENT,
    unique_id TEXT UNIQUE, ctime REAL, mtime REAL, username TEXT,
    numbers BLOB, positions BLOB, cell BLOB, pbc INTEGER,
    initial_magmoms BLOB, initial_charges BLOB, masses BLOB, tags BLOB, momenta BLOB,
    constraints TEXT, calculator TEXT, calculator_parameters TEXT,
    energy REAL, free_energy REAL, forces BLOB, stress BLOB, dipole BLOB, magmoms BLOB,
    magmom REAL, charges BLOB, key_value_pairs TEXT, data TEXT,
    natoms INTEGER, fmax REAL, smax REAL, volume REAL, mass REAL, charge REAL)"""
_INFORMATION = "CREATE TABLE IF NOT EXISTS information (name TEXT, value TEXT)"
# seconds per (Julian) year / the year-2000 epoch: ASE stores times as years since 2000
_YEAR = 31557600.0
_T2000 = 946681200.0


def _get_keypath(k_max: int, rc: float, prop) -> str:
    """sqlite.py:27-32: 'neighbors.{k_max}.{100 rc}.{prop}_max'."""
    name = prop.name if hasattr(prop, 'name') else str(prop)
    return f"neighbors.{k_max}.{rc * 100.0:.0f}.{name}_max"


def _nested_get(dct, keypath):
    for key in keypath.split('.'):
        if not isinstance(dct, dict) or key not in dct:
            return None
        dct = dct[key]
    return dct


def _nested_set(dct, keypath, value):
    keys = keypath.split('.')
    for key in keys[:-1]:
        dct = dct.setdefault(key, {})
    dct[keys[-1]] = value


def _blob(array, dtype):
    if array is None:
        return None
    return np.ascontiguousarray(array, dtype=np.dtype(dtype).newbyteorder('<')).tobytes()


def _array(blob, dtype, shape=None):
    if blob is None:
        return None
    a = np.frombuffer(blob, dtype=np.dtype(dtype).newbyteorder('<')).astype(dtype)
    return a if shape is None else a.reshape(shape)


class CoreDatabase:
    def __init__(self, filename, serial=True):
        self.filename = str(filename)
        self._serial = serial
        new = not os.path.exists(self.filename)
        self._con = sqlite3.connect(self.filename)
        if new:
            self._con.execute(_SYSTEMS)
            self._con.execute(_INFORMATION)
            self._con.execute("INSERT INTO information VALUES ('version', '8')")
            self._con.commit()
        self._metadata = {}
        tables = {r[0] for r in self._con.execute(
            "SELECT name FROM sqlite_master WHERE type='table'")}
        if 'systems' not in tables:
            raise IOError(f"{self.filename} is not an ASE database (no table 'systems')")
        if 'information' in tables:
            row = self._con.execute(
                "SELECT value FROM information WHERE name='metadata'").fetchone()
            if row:
                self._metadata = json.loads(row[0])

    def __str__(self):
        return f"CoreDataBase@{self.filename}"

    def __len__(self):
        return int(self._con.execute("SELECT COUNT(*) FROM systems").fetchone()[0])

    def close(self):
        self._con.close()

    # -- structures ---------------------------------------------------------------------
    @staticmethod
    def _selection_id(selection, kwargs):
        if 'id' in kwargs:
            return int(kwargs['id'])
        if isinstance(selection, (int, np.integer)):
            return int(selection)
        if isinstance(selection, str) and selection.replace(' ', '').startswith('id='):
            return int(selection.replace(' ', '')[3:])
        raise KeyError(f"only selections by id are supported (got {selection!r}, {kwargs})")

    def get_atoms(self, selection=None, attach_calculator=False,
                  add_additional_information=False, **kwargs) -> Atoms:
        """sqlite.py:58-77.  The labels of the row (energy, forces, stress: what ASE attaches
        as a `SinglePointCalculator`) are always placed in `atoms.info`; key-value pairs and
        `data` join them with `add_additional_information`."""
        aid = self._selection_id(selection, kwargs)
        row = self._con.execute(
            "SELECT numbers, positions, cell, pbc, energy, free_energy, forces, stress, "
            "key_value_pairs, data FROM systems WHERE id=?", (aid,)).fetchone()
        if row is None:
            raise KeyError(f"no row with id = {aid}")
        numbers, positions, cell, pbc, energy, free_energy, forces, stress, kvp, data = row
        numbers = _array(numbers, np.int32)
        n = len(numbers)
        pbc = int(pbc or 0)
        info = {}
        if energy is not None:
            info['energy'] = float(energy)
        if free_energy is not None:
            info['free_energy'] = float(free_energy)
        if forces is not None:
            info['forces'] = _array(forces, np.float64, (n, 3))
        if stress is not None:
            s = _array(stress, np.float64)
            if s.size == 9:
                s = s.reshape(3, 3)[[0, 1, 2, 1, 0, 0], [0, 1, 2, 2, 2, 1]]
            info['stress'] = s
        if add_additional_information:
            if kvp:
                info['key_value_pairs'] = json.loads(kvp)
            if data and data != 'null':
                try:
                    info['data'] = json.loads(data)
                except (TypeError, ValueError):
                    pass                # newer ASE versions store a binary blob: labels suffice
                else:
                    for key in ('etemperature', 'eentropy'):
                        if isinstance(info['data'], dict) and key in info['data']:
                            info[key] = float(np.atleast_1d(info['data'][key])[0])
        cell = np.zeros((3, 3)) if cell is None else _array(cell, np.float64, (3, 3))
        return Atoms(numbers=numbers, positions=_array(positions, np.float64, (n, 3)),
                     cell=cell, pbc=[bool(pbc & 1), bool(pbc & 2), bool(pbc & 4)], info=info)

    def write(self, atoms, key_value_pairs=None, data=None) -> int:
        """Append one labelled structure (labels from `atoms.info`); returns its id."""
        info = getattr(atoms, 'info', {})
        forces = info.get('forces')
        stress = info.get('stress')
        pbc = np.asarray(atoms.pbc, dtype=bool).reshape(3)
        now = (time.time() - _T2000) / _YEAR
        uid = '%032x' % int.from_bytes(os.urandom(16), 'big')
        fmax = float(np.sqrt((np.asarray(forces) ** 2).sum(axis=1)).max()) \
            if forces is not None and len(atoms) else None
        cur = self._con.execute(
            "INSERT INTO systems (unique_id, ctime, mtime, username, numbers, positions, cell, "
            "pbc, calculator, calculator_parameters, energy, free_energy, forces, stress, "
            "key_value_pairs, data, natoms, fmax, smax, volume) "
            "VALUES (?,?,?,?,?,?,?,?,?,?,?,?,?,?,?,?,?,?,?,?)",
            (uid, now, now, os.environ.get('USER', 'root'),
             _blob(atoms.numbers, np.int32), _blob(atoms.positions, np.float64),
             _blob(atoms.cell, np.float64), int(pbc[0]) + 2 * int(pbc[1]) + 4 * int(pbc[2]),
             'unknown', '{}',
             None if 'energy' not in info else float(info['energy']),
             None if 'free_energy' not in info else float(info['free_energy']),
             _blob(forces, np.float64),
             None if stress is None else _blob(np.asarray(stress).reshape(-1), np.float64),
             json.dumps(key_value_pairs or {}), json.dumps(data or {}), len(atoms), fmax,
             None if stress is None else float(np.abs(stress).max()),
             float(atoms.get_volume())))
        self._con.commit()
        return int(cur.lastrowid)

    def to_dataset(self):
        """Every structure with its labels as an in-memory `io.read.Dataset`."""
        from tensoralloy_b200.io.read import Dataset
        images = [self.get_atoms(id=k, add_additional_information=True)
                  for k in range(1, 1 + len(self))]
        ds = Dataset.from_images(images, extxyz=self._metadata.get('extxyz', True))
        ds.metadata.update({k: v for k, v in self._metadata.items() if k != 'max_occurs'})
        return ds

    # -- metadata -----------------------------------------------------------------------
    @property
    def metadata(self) -> dict:
        return self._metadata.copy()

    @metadata.setter
    def metadata(self, dct):
        self._metadata = dict(dct)
        self._write_metadata()

    def _write_metadata(self):
        """sqlite.py:94-111."""
        md = json.dumps(self._metadata)
        self._con.execute(_INFORMATION)
        cur = self._con.execute("SELECT COUNT(*) FROM information WHERE name='metadata'")
        if cur.fetchone()[0]:
            self._con.execute("UPDATE information SET value=? WHERE name='metadata'", [md])
        else:
            self._con.execute("INSERT INTO information VALUES (?, ?)", ('metadata', md))
        self._con.commit()

    @property
    def max_occurs(self) -> Counter:
        if 'max_occurs' not in self._metadata:
            self._find_max_occurs()
        return Counter(self._metadata.get('max_occurs'))

    has_forces = property(lambda self: self._metadata.get('forces', True))
    has_stress = property(lambda self: self._metadata.get('stress', True))
    has_periodic_structures = property(lambda self: self._metadata.get('periodic', True))

    def _find_max_occurs(self):
        max_occurs = Counter()
        for aid in range(1, 1 + len(self)):
            for element, n in Counter(self.get_atoms(id=aid).get_chemical_symbols()).items():
                max_occurs[element] = max(max_occurs[element], n)
        self._metadata['max_occurs'] = dict(max_occurs)
        self._write_metadata()

    def get_atomic_static_energy(self, allow_calculation=False):
        """sqlite.py:156-168; the fit itself (sqlite.py:326-375): least squares of the total
        energies on the element counts."""
        key = 'atomic_static_energy'
        dct = self._metadata.get(key, {})
        if not dct and allow_calculation:
            elements = sorted(self.max_occurs.keys())
            n = len(self)
            counts = np.zeros((n, len(elements)))
            y = np.zeros(n)
            for aid in range(1, 1 + n):
                atoms = self.get_atoms(id=aid)
                c = Counter(atoms.get_chemical_symbols())
                counts[aid - 1] = [c[e] for e in elements]
                y[aid - 1] = atoms.info['energy']
            x = np.linalg.lstsq(counts, y, rcond=None)[0]
            dct = {e: float(v) for e, v in zip(elements, x)}
            self._metadata[key] = dct
            self._write_metadata()
        return dct

    # -- neighbour sizes ----------------------------------------------------------------
    def _get_neighbor_property(self, rc, prop, allow_calculation=False):
        """sqlite.py:170-199."""
        val = None
        for k_max in (3, 2):
            val = _nested_get(self._metadata, _get_keypath(k_max, rc, prop))
            if val is not None:
                return val
        if allow_calculation:
            nijk = prop in (NeighborProperty.nijk, NeighborProperty.ij2k)
            ij2k = prop == NeighborProperty.ij2k
            val = self.update_neighbor_meta(rc=rc, nijk=nijk, ij2k=ij2k)[prop]
        return val

    def get_nij_max(self, rc, allow_calculation=False):
        return self._get_neighbor_property(rc, NeighborProperty.nij, allow_calculation)

    def get_nijk_max(self, rc, allow_calculation=False, symmetric=True):
        value = self._get_neighbor_property(rc, NeighborProperty.nijk, allow_calculation)
        return value * (2 - int(symmetric))

    def get_nnl_max(self, rc, allow_calculation=False):
        return self._get_neighbor_property(rc, NeighborProperty.nnl, allow_calculation)

    def get_ij2k_max(self, rc, allow_calculation=False):
        return self._get_neighbor_property(rc, NeighborProperty.ij2k, allow_calculation)

    def update_neighbor_meta(self, rc, nijk=False, ij2k=False, n_jobs=-1,
                             verbose=False) -> NeighborSize:
        """sqlite.py:234-298: the maxima over the database, stored under both k_max keys."""
        from tensoralloy_b200.neighbor import find_neighbor_size_of_atoms
        results = [find_neighbor_size_of_atoms(self.get_atoms(id=aid), rc, find_ij2k=ij2k,
                                               find_nijk=nijk)
                   for aid in range(1, 1 + len(self))]
        maxvals = NeighborSize(**{prop.name: max(r[prop] for r in results)
                                  for prop in NeighborProperty})
        k_max = 3 if (nijk or ij2k) else 2
        for prop in NeighborProperty:
            _nested_set(self._metadata, _get_keypath(k_max, rc, prop), maxvals[prop])
            if ij2k or nijk:
                _nested_set(self._metadata, _get_keypath(2, rc, prop),
                            0 if prop == NeighborProperty.nijk else maxvals[prop])
        self._write_metadata()
        return maxvals

    def may_update_neighbor_meta(self, rc, newval: NeighborSize, angular=False) -> bool:
        """sqlite.py:300-323."""
        k_max = 2 + int(angular)
        updated = False
        for prop in NeighborProperty:
            keypath = _get_keypath(k_max, rc, prop)
            val = _nested_get(self._metadata, keypath)
            if val is None or newval[prop] > val:
                updated = True
                _nested_set(self._metadata, keypath,
                            newval[prop] if val is None else max(val, newval[prop]))
        if updated:
            self._write_metadata()
        return updated


def connect(filename) -> CoreDatabase:
    return CoreDatabase(filename)
