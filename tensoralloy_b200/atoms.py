"""
Minimal stand-in for `ase.Atoms` -- ASE is a third-party dependency of the
reference (requirements.txt:3) and is not installed in this image.  The host-side
mirror accepts either a real `ase.Atoms` (duck-typed) or this class; only the
accessors the hot path uses are provided (same names and meaning as ASE).
"""
from collections import Counter

import numpy as np

# chemical symbols up to Z = 96 (enough for every reference fixture)
chemical_symbols = [
    'X', 'H', 'He', 'Li', 'Be', 'B', 'C', 'N', 'O', 'F', 'Ne', 'Na', 'Mg', 'Al',
    'Si', 'P', 'S', 'Cl', 'Ar', 'K', 'Ca', 'Sc', 'Ti', 'V', 'Cr', 'Mn', 'Fe', 'Co',
    'Ni', 'Cu', 'Zn', 'Ga', 'Ge', 'As', 'Se', 'Br', 'Kr', 'Rb', 'Sr', 'Y', 'Zr',
    'Nb', 'Mo', 'Tc', 'Ru', 'Rh', 'Pd', 'Ag', 'Cd', 'In', 'Sn', 'Sb', 'Te', 'I',
    'Xe', 'Cs', 'Ba', 'La', 'Ce', 'Pr', 'Nd', 'Pm', 'Sm', 'Eu', 'Gd', 'Tb', 'Dy',
    'Ho', 'Er', 'Tm', 'Yb', 'Lu', 'Hf', 'Ta', 'W', 'Re', 'Os', 'Ir', 'Pt', 'Au',
    'Hg', 'Tl', 'Pb', 'Bi', 'Po', 'At', 'Rn', 'Fr', 'Ra', 'Ac', 'Th', 'Pa', 'U',
    'Np', 'Pu', 'Am', 'Cm']
atomic_numbers = {s: z for z, s in enumerate(chemical_symbols)}

# ase.units (CODATA 2014, ASE >= 3.21 default): eV / A^3 -> GPa
GPa = 1.0 / 160.21766208


# Covalent radii (Angstrom) by atomic number, Z = 1 .. 96: B. Cordero et al., "Covalent radii
# revisited", Dalton Trans. 2008, 2832-2838 (C sp3; Mn, Fe, Co low spin) -- the table behind
# `ase.data.covalent_radii`, which the reference indexes in nn/atomic/grap.py:621-628.  ASE is not
# installed here, so the values are stated, not imported (unpinned against ASE); index 0 and
# elements beyond Cm hold ASE's placeholder 0.2.
covalent_radii = np.array([
    0.20,
    0.31, 0.28,
    1.28, 0.96, 0.84, 0.76, 0.71, 0.66, 0.57, 0.58,
    1.66, 1.41, 1.21, 1.11, 1.07, 1.05, 1.02, 1.06,
    2.03, 1.76, 1.70, 1.60, 1.53, 1.39, 1.39, 1.32, 1.26, 1.24, 1.32, 1.22,
    1.22, 1.20, 1.19, 1.20, 1.20, 1.16,
    2.20, 1.95, 1.90, 1.75, 1.64, 1.54, 1.47, 1.46, 1.42, 1.39, 1.45, 1.44,
    1.42, 1.39, 1.39, 1.38, 1.39, 1.40,
    2.44, 2.15,
    2.07, 2.04, 2.03, 2.01, 1.99, 1.98, 1.98, 1.96, 1.94, 1.92, 1.92, 1.89, 1.90, 1.87, 1.87,
    1.75, 1.70, 1.62, 1.51, 1.44, 1.41, 1.36, 1.36, 1.32,
    1.45, 1.46, 1.48, 1.40, 1.50, 1.50,
    2.60, 2.21,
    2.15, 2.06, 2.00, 1.96, 1.90, 1.87, 1.80, 1.69,
])


class Atoms:
    """Positions (A), cell (rows = lattice vectors), pbc and chemical symbols."""

    def __init__(self, symbols=None, positions=None, cell=None, pbc=False,
                 numbers=None, info=None):
        if symbols is None and numbers is not None:
            symbols = [chemical_symbols[int(z)] for z in numbers]
        if isinstance(symbols, str):
            symbols = _parse_formula(symbols)
        self._symbols = list(symbols)
        self.numbers = np.array([atomic_numbers[s] for s in self._symbols], dtype=np.int64)
        self.positions = np.array(positions, dtype=np.float64).reshape(-1, 3)
        if len(self._symbols) != len(self.positions):
            raise ValueError("symbols and positions differ in length")
        cell = np.zeros((3, 3)) if cell is None else np.asarray(cell, dtype=np.float64)
        if cell.shape == (3,):
            cell = np.diag(cell)
        self.cell = cell.reshape(3, 3).copy()
        pbc = np.asarray(pbc, dtype=bool).reshape(-1)
        self.pbc = np.repeat(pbc, 3) if pbc.size == 1 else pbc.copy()
        self.info = dict(info or {})
        self.calc = None

    def __len__(self):
        return len(self._symbols)

    def copy(self):
        new = Atoms.__new__(Atoms)      # no per-atom symbol work
        new._symbols = list(self._symbols)
        new.numbers = self.numbers.copy()
        new.positions = self.positions.copy()
        new.cell = self.cell.copy()
        new.pbc = self.pbc.copy()
        new.info = dict(self.info)
        new.calc = None
        return new

    def get_chemical_symbols(self):
        return list(self._symbols)

    def get_atomic_numbers(self):
        return self.numbers.copy()

    def get_positions(self):
        return self.positions.copy()

    def set_positions(self, positions):
        self.positions = np.array(positions, dtype=np.float64).reshape(-1, 3)

    def get_cell(self, complete=False):
        return self.cell.copy()

    def set_cell(self, cell, scale_atoms=False):
        cell = np.asarray(cell, dtype=np.float64).reshape(3, 3)
        if scale_atoms:
            m = np.linalg.solve(self.cell, cell)
            self.positions = self.positions @ m
        self.cell = cell.copy()

    def get_pbc(self):
        return self.pbc.copy()

    def get_volume(self):
        return float(abs(np.linalg.det(self.cell)))

    def get_chemical_formula(self, mode='hill'):
        if mode == 'reduce':
            out, prev, cnt = [], None, 0
            for s in self._symbols + [None]:
                if s == prev:
                    cnt += 1
                else:
                    if prev is not None:
                        out.append(prev + (str(cnt) if cnt > 1 else ''))
                    prev, cnt = s, 1
            return ''.join(out)
        c = Counter(self._symbols)
        return ''.join(k + (str(c[k]) if c[k] > 1 else '') for k in sorted(c))

    def repeat(self, rep):
        if isinstance(rep, int):
            rep = (rep, rep, rep)
        pos, sym = [], []
        for a in range(rep[0]):
            for b in range(rep[1]):
                for c in range(rep[2]):
                    pos.append(self.positions + np.array([a, b, c], float) @ self.cell)
                    sym.extend(self._symbols)
        cell = self.cell * np.asarray(rep, dtype=float)[:, None]
        return Atoms(sym, np.concatenate(pos), cell, self.pbc, info=dict(self.info))

    __mul__ = repeat

    def rattle(self, stdev=0.001, seed=42):
        rng = np.random.RandomState(seed)
        self.positions = self.positions + rng.normal(scale=stdev,
                                                     size=self.positions.shape)

    # ASE calculator plumbing -------------------------------------------------
    def set_calculator(self, calc):
        self.calc = calc

    def get_potential_energy(self):
        return self.calc.get_potential_energy(self)

    def get_forces(self):
        return self.calc.get_forces(self)

    def get_stress(self, voigt=True):
        return self.calc.get_stress(self, voigt=voigt)


def _parse_formula(formula):
    import re
    out = []
    for sym, cnt in re.findall(r'([A-Z][a-z]?)(\d*)', formula):
        out.extend([sym] * (int(cnt) if cnt else 1))
    return out


def bulk_fcc(symbol, a, repeat=(1, 1, 1), cubic=True):
    """fcc conventional (cubic, 4-atom) cell repeated: the stand-in for
    `ase.build.bulk(symbol, cubic=True) * repeat` used by the reference tests.
    Atom order follows ASE: cell-major over (a, b, c) repeats."""
    assert cubic
    base = np.array([[0, 0, 0], [0, .5, .5], [.5, 0, .5], [.5, .5, 0]]) * a
    atoms = Atoms([symbol] * 4, base, np.eye(3) * a, True)
    return atoms.repeat(repeat)


def fcc_positions(a, nx, ny, nz, dtype=np.float64):
    """Vectorised fcc generator for large synthetic lattices (bench / tests)."""
    base = np.array([[0, 0, 0], [0, .5, .5], [.5, 0, .5], [.5, .5, 0]], dtype=dtype)
    ix, iy, iz = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz),
                             indexing='ij')
    cells = np.stack([ix.ravel(), iy.ravel(), iz.ravel()], axis=1).astype(dtype)
    pos = (cells[:, None, :] + base[None, :, :]).reshape(-1, 3) * a
    cell = np.diag([nx * a, ny * a, nz * a]).astype(dtype)
    return pos, cell


def bulk_hcp(symbol, a, c, repeat=(1, 1, 1)):
    """hcp primitive cell (2 atoms) repeated; stand-in for
    `ase.build.bulk(symbol, 'hcp', a=a, c=c) * repeat` (the Be crystal of the
    reference's nn/constraint/data.py:42-50)."""
    cell = np.array([[a, 0.0, 0.0],
                     [-0.5 * a, 0.5 * np.sqrt(3.0) * a, 0.0],
                     [0.0, 0.0, c]])
    scaled = np.array([[0.0, 0.0, 0.0], [1.0 / 3.0, 2.0 / 3.0, 0.5]])
    atoms = Atoms([symbol] * 2, scaled @ cell, cell, True)
    return atoms.repeat(repeat)
