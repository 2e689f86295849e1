"""
Labelled structures from `extxyz` / `xyz` files -- the front of the reference's input
pipeline, tensoralloy/io/read.py:43-246 (`read_file`, `_read_extxyz`), which parses with
`ase.io.extxyz.read_xyz` and stores into SQLite (`io/sqlite.py`).  Here the result stays in
memory (`Dataset`: a list of `Atoms` with their labels plus the reference's metadata dict)
and feeds the trainers directly (`Dataset.fill(trainer)`); the SQLite / TFRecord stores are
out of scope.

Kept from the reference: unit conversion of energy / forces / stress to eV, eV/A, eV/A^3
(read.py:137-155); `fmax` filtering (:119-121); a zero cell becomes a cube of
20 + 5 (N // 50) A (:125-130); plain xyz frames get zero forces (:141-144); `max_occurs`,
`periodic`, `stress`, `unit_conversion` metadata (:178-187); `etemperature` / `eentropy`
/ `free_energy` are carried in `atoms.info` (atoms_utils.py).
Stress follows ASE's extxyz reader: a 3x3 `stress` is stored as the Voigt vector
(xx, yy, zz, yz, xz, xy); a `virial` becomes stress = -virial / volume.
"""
import re
import shlex
from collections import Counter
from os.path import splitext

import numpy as np

from tensoralloy_b200.atoms import Atoms
from tensoralloy_b200.io.units import get_conversion_units

XYZ_FORMATS = ('normal', 'xyz', 'extxyz', 'stepmax')
HARTREE = 27.211386024367243            # ase.units.Hartree (CODATA 2014)


def cellpar_to_cell(cellpar):
    """ase.geometry.cellpar_to_cell with its default orientation (a along x, b in the xy
    plane): [a, b, c, alpha, beta, gamma] (A, degrees) -> rows = lattice vectors."""
    a, b, c, alpha, beta, gamma = [float(x) for x in cellpar]

    def cosd(x):            # exact zeros at 90 degrees, as ASE does
        return 0.0 if abs(abs(x) - 90.0) < 1e-12 else np.cos(np.radians(x))

    ca, cb, cg = cosd(alpha), cosd(beta), cosd(gamma)
    sg = 1.0 if abs(abs(gamma) - 90.0) < 1e-12 else np.sin(np.radians(gamma))
    cy = (ca - cb * cg) / sg
    cz = np.sqrt(max(1.0 - cb * cb - cy * cy, 0.0))
    return np.array([[a, 0.0, 0.0], [b * cg, b * sg, 0.0], [c * cb, c * cy, c * cz]])


def cell_to_cellpar(cell):
    """ase.geometry.cell_to_cellpar: lengths and angles (degrees)."""
    cell = np.asarray(cell, dtype=np.float64).reshape(3, 3)
    L = np.linalg.norm(cell, axis=1)
    ang = []
    for i, j in ((1, 2), (0, 2), (0, 1)):
        ll = L[i] * L[j]
        ang.append(np.degrees(np.arccos(np.clip(cell[i] @ cell[j] / ll, -1, 1)))
                   if ll > 1e-16 else 90.0)
    return np.array(list(L) + ang)
_VOIGT = ((0, 0), (1, 1), (2, 2), (1, 2), (0, 2), (0, 1))
_KV = re.compile(r'(\w[\w\-:.]*)\s*=\s*("[^"]*"|\{[^}]*\}|\S+)')


class Dataset:
    """In-memory stand-in for the reference's `CoreDatabase` as far as the training step
    needs it: `len`, iteration / indexing over labelled `Atoms`, `.metadata`,
    `.max_occurs`."""

    def __init__(self, images, metadata):
        self.images = list(images)
        self.metadata = dict(metadata)

    @classmethod
    def from_images(cls, images, extxyz=True):
        """A dataset over labelled `Atoms` from any reader (e.g. `io.vasp.read_vasp_xml`):
        every `atoms.info` holds at least `energy`; missing forces become zeros; the
        metadata dict is rebuilt as `read_file` would (read.py:178-187)."""
        images = list(images)
        max_occurs, periodic = Counter(), False
        use_stress = bool(images) and all('stress' in a.info for a in images)
        for k, atoms in enumerate(images):
            if 'energy' not in atoms.info:
                raise ValueError(f"structure {k} has no energy label")
            if 'forces' not in atoms.info:
                atoms.info['forces'] = np.zeros_like(atoms.positions)
            periodic = bool(np.any(atoms.pbc)) or periodic
            for symbol, cnt in Counter(atoms.get_chemical_symbols()).items():
                max_occurs[symbol] = max(max_occurs[symbol], cnt)
        return cls(images, {'max_occurs': dict(max_occurs), 'extxyz': extxyz, 'forces': True,
                            'stress': use_stress, 'periodic': periodic,
                            'unit_conversion': {'energy': 1.0, 'forces': 1.0, 'stress': 1.0}})

    max_occurs = property(lambda self: Counter(self.metadata['max_occurs']))
    has_stress = property(lambda self: bool(self.metadata.get('stress')))
    has_periodic_structures = property(lambda self: bool(self.metadata.get('periodic')))

    def __len__(self):
        return len(self.images)

    def __iter__(self):
        return iter(self.images)

    def __getitem__(self, k):
        return self.images[k]

    def labels(self, k):
        """(energy, forces [N,3], stress Voigt [6] or zeros) of structure k."""
        info = self.images[k].info
        stress = info.get('stress')
        return info['energy'], info['forces'], np.zeros(6) if stress is None else stress

    def fill(self, trainer, indices=None):
        """Queue the structures in a trainer (`AtomicNNTrainer`, `EamTrainer`,
        `GrapFilterTrainer`: `add_structure(atoms, energy, forces, stress)`)."""
        for k in (range(len(self)) if indices is None else indices):
            trainer.add_structure(self.images[k], *self.labels(k))
        return trainer


    # -- TFRecord files (train/dataset/dataset.py:168-400) -----------------------------
    def record_signature(self, transformer):
        """dataset.py:260-278: 'k{2|3}-rc{rcut:.2f}-fp{32|64}'."""
        from tensoralloy_b200.precision import get_float_dtype
        bits = 32 if get_float_dtype().as_numpy_dtype == np.float32 else 64
        return f"k{3 if transformer.angular else 2}-rc{transformer.rcut:.2f}-fp{bits}"

    def to_records(self, savedir, transformer, name='dataset', test_size=0.2, seed=611,
                   write='all', neighbor_lists=None):
        """Split into a training and a test subset (sklearn `train_test_split`, or the given
        1-based test ids) and write each as one TFRecord file of encoded examples, named as the
        reference names them: `{name}-{test|train}-{signature}-{size}.universal.tfrecords`
        (dataset.py:280-342).  `transformer`: a `BatchUniversalTransformer` with `nij_max`
        (and `nijk_max`) set.  Returns {'test': path, 'train': path} of the files written."""
        import os
        from tensoralloy_b200.transformer.tfrecord import TFRecordWriter
        assert write in ('all', 'eval', 'train')
        ids = list(range(1, 1 + len(self)))
        if isinstance(test_size, (list, tuple)):
            test = list(test_size)
            train = [x for x in ids if x not in test]
        else:
            from sklearn.model_selection import train_test_split
            train, test = train_test_split(ids, random_state=seed, test_size=test_size)
        os.makedirs(savedir, exist_ok=True)
        sig = self.record_signature(transformer)
        out = {}
        for key, subset in (('test', test), ('train', train)):
            if write != 'all' and write != {'test': 'eval', 'train': 'train'}[key]:
                continue
            path = os.path.join(savedir,
                                f"{name}-{key}-{sig}-{len(subset)}.universal.tfrecords")
            with TFRecordWriter(path) as writer:
                for k in subset:
                    nl = None if neighbor_lists is None else neighbor_lists[k - 1]
                    writer.write(transformer.encode(self.images[k - 1],
                                                    neighbor_list=nl).SerializeToString())
            out[key] = path
        return out

    @classmethod
    def from_records(cls, filename, transformer):
        """The labelled structures of one TFRecord file (written by `to_records` or by the
        reference with the same transformer settings)."""
        from tensoralloy_b200.transformer.tfrecord import read_tfrecords
        images = [transformer.decode_atoms(transformer.decode_protobuf(rec))
                  for rec in read_tfrecords(filename)]
        return cls.from_images(images)


def _parse_header(line):
    out = {}
    for key, val in _KV.findall(line):
        if val[:1] in '"{':
            val = val[1:-1]
        out[key] = val
    return out


def _value(text):
    toks = text.split()
    try:
        vals = [float(t) for t in toks]
    except ValueError:
        return text
    if len(vals) == 1:
        return vals[0]
    return np.array(vals)


def _columns(spec):
    """`Properties=species:S:1:pos:R:3:forces:R:3` -> [(name, type, ncols)]."""
    f = spec.split(':')
    if len(f) % 3:
        raise ValueError(f"malformed Properties specification: '{spec}'")
    return [(f[k], f[k + 1], int(f[k + 2])) for k in range(0, len(f), 3)]


def _frames(fp):
    while True:
        line = fp.readline()
        if not line:
            return
        if not line.strip():
            continue
        n = int(line.split()[0])
        comment = fp.readline().rstrip('\n')
        rows = [fp.readline().split() for _ in range(n)]
        if rows and not rows[-1]:
            raise ValueError("truncated xyz frame")
        yield n, comment, rows


def _atoms_from_frame(n, comment, rows, extxyz):
    info = {}
    cell, pbc = np.zeros((3, 3)), False
    if extxyz:
        head = _parse_header(comment)
        cols = _columns(head.pop('Properties', 'species:S:1:pos:R:3'))
        if 'Lattice' in head:
            # ASE: the nine numbers are the cell vectors in Fortran order = rows a, b, c
            cell = np.array(head.pop('Lattice').split(), dtype=np.float64).reshape(3, 3)
            pbc = True
        if 'pbc' in head:
            pbc = [t.upper().startswith('T') for t in shlex.split(head.pop('pbc'))]
        for k, v in head.items():
            info[k] = _value(v)
    elif extxyz is None:
        # STEPMAX xyz (io/xyz.py:18-31, read.py:101-112): energy in Hartree, six cell
        # parameters and the label 'Cartesian'
        cols = [('species', 'S', 1), ('pos', 'R', 3)]
        f = comment.split()
        if len(f) != 8 or f[-1].lower() != 'cartesian':
            raise ValueError("stepmax xyz: expected 'energy a b c alpha beta gamma Cartesian'")
        info['energy'] = float(f[0]) * HARTREE
        cell, pbc = cellpar_to_cell(f[1:7]), True
    else:
        cols = [('species', 'S', 1), ('pos', 'R', 3)]
        info['energy'] = float(comment.split()[0])
    data, c0 = {}, 0
    for name, typ, nc in cols:
        block = [r[c0:c0 + nc] for r in rows]
        data[name] = [b[0] for b in block] if typ == 'S' else \
            np.array(block, dtype=np.float64 if typ == 'R' else np.int64)
        c0 += nc
    symbols = data.get('species')
    if symbols is None and 'Z' in data:
        from tensoralloy_b200.atoms import chemical_symbols
        symbols = [chemical_symbols[int(z)] for z in data['Z'].reshape(-1)]
    atoms = Atoms(symbols, data['pos'].reshape(n, 3), cell, pbc)
    # ASE's reader takes a per-atom `stress:R:1` column of a six-atom frame as the Voigt
    # stress (any array named `stress` of length 6); the reference's SNAP/Ni fixture relies
    # on it (io/tests/test_read.py:55-72)
    if 'stress' in data and 'stress' not in info and np.size(data['stress']) == 6:
        info['stress'] = np.asarray(data['stress'], dtype=np.float64).reshape(6)
    forces = data.get('forces', data.get('force'))
    if forces is not None:
        info['forces'] = np.asarray(forces, dtype=np.float64).reshape(n, 3)
    for key in ('stress', 'virial'):
        if key in info and np.size(info[key]) == 9:
            info[key] = np.asarray(info[key], dtype=np.float64).reshape(3, 3)
    atoms.info = info
    return atoms


def _read_xyz_like(filename, units, extxyz, num_examples=None, fmax=None):
    to_eV, to_eV_A, to_eV_A3 = get_conversion_units(units)
    images, max_occurs = [], Counter()
    use_stress, periodic = None, False
    with open(filename) as fp:
        for k, (n, comment, rows) in enumerate(_frames(fp)):
            if num_examples is not None and k >= num_examples:
                break
            atoms = _atoms_from_frame(n, comment, rows, extxyz)
            info = atoms.info
            if 'energy' not in info:
                raise ValueError(f"{filename}: frame {k} has no energy")
            if fmax is not None and 'forces' in info and np.abs(info['forces']).max() > fmax:
                continue
            if np.abs(atoms.cell).sum() < 1e-8:
                atoms.cell = np.eye(3) * (20.0 + (len(atoms) // 50) * 5.0)
            info['energy'] = float(info['energy']) * to_eV
            if extxyz is True and 'forces' in info:
                info['forces'] = info['forces'] * to_eV_A
            else:
                info['forces'] = np.zeros_like(atoms.positions)
            if 'stress' not in info and 'virial' in info:
                info['stress'] = -np.asarray(info['virial']) / atoms.get_volume()
            if use_stress is None:
                use_stress = 'stress' in info
            if use_stress:
                if 'stress' not in info:
                    raise ValueError(f"{filename}: frame {k} has no stress tensor")
                s = np.asarray(info['stress'], dtype=np.float64)
                if s.shape == (3, 3):
                    s = np.array([s[a, b] for a, b in _VOIGT])
                info['stress'] = s * to_eV_A3
            periodic = bool(np.any(atoms.pbc)) or periodic
            for symbol, cnt in Counter(atoms.get_chemical_symbols()).items():
                max_occurs[symbol] = max(max_occurs[symbol], cnt)
            images.append(atoms)
    metadata = {'max_occurs': dict(max_occurs), 'extxyz': extxyz is True, 'forces': True,
                'stress': bool(use_stress), 'periodic': periodic,
                'unit_conversion': {'energy': to_eV, 'forces': to_eV_A, 'stress': to_eV_A3}}
    return Dataset(images, metadata)


def read_file(filename, units=None, num_examples=None, file_type=None, verbose=False,
              append_to=None, fmax=None):
    """read.py:190-246.  `file_type`: 'extxyz', 'xyz' / 'normal' (second line = energy),
    'stepmax' (energy in Hartree + cell parameters); 'db' (SQLite) is refused -- the stores
    are out of scope."""
    if file_type is None:
        file_type = splitext(filename)[1][1:]
    if units is None:
        units = {'energy': 'eV', 'forces': 'eV/Angstrom'}
    if append_to is not None:
        raise NotImplementedError("appending to an SQLite database is out of scope")
    if file_type == 'extxyz':
        return _read_xyz_like(filename, units, True, num_examples, fmax)
    if file_type in ('xyz', 'normal'):
        return _read_xyz_like(filename, units, False, num_examples, fmax)
    if file_type == 'stepmax':
        return _read_xyz_like(filename, units, None, num_examples, fmax)
    if file_type == 'db':
        raise NotImplementedError("SQLite databases are out of scope (DESIGN.md 7)")
    raise ValueError("Unknown file type: {}".format(file_type))
