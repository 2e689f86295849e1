"""
LAMMPS setfl (eam/alloy) and ADP setfl tables: reader and writer.
Mirror of the reference's tensoralloy/io/lammps.py:60-296 (`Spline`, `SetFL`,
`_read_setfl` :107-221, `read_eam_alloy_setfl`, `read_adp_setfl`,
`write_adp_setfl` :280) with the same conventions:
  * the phi column of the file stores r * phi(r) -> divided out on reading
    (lammps.py:200-201; the r = 0 entry is kept as read);
  * ADP u(r), w(r) are stored raw (lammps.py:196-199);
  * x grids are  k * dx, k = 0 .. n-1  (lammps.py:95-103);
  * splines are natural cubic (second derivative zero at both ends).
"""
from dataclasses import dataclass
from typing import Dict, List

import numpy as np


@dataclass
class Spline:
    bc_start: float
    bc_end: float
    x: np.ndarray
    y: np.ndarray
    natural_boundary: bool

    def coefficients(self):
        """Per-interval (c0, c1, c2, c3) of the natural cubic spline through (x, y):
        f(x) = c0 + c1 t + c2 t^2 + c3 t^3, t = x - x_k.  Shape [n-1, 4]."""
        from scipy.interpolate import CubicSpline
        bc = 'natural' if self.natural_boundary else ((1, self.bc_start), (1, self.bc_end))
        cs = CubicSpline(self.x, self.y, bc_type=bc)
        return np.ascontiguousarray(cs.c[::-1].T)        # scipy stores highest power first


@dataclass
class SetFL:
    elements: List[str]
    rho: Dict[str, Spline]
    phi: Dict[str, Spline]
    embed: Dict[str, Spline]
    dipole: Dict[str, Spline]
    quadrupole: Dict[str, Spline]
    nr: int
    dr: float
    nrho: int
    drho: float
    rcut: float
    atomic_masses: List[float]
    lattice_constants: List[float]
    lattice_types: List[str]


def _spline(n, dx, y):
    return Spline(0.0, 0.0, np.linspace(0.0, n * dx, n, endpoint=False),
                  np.asarray(y, dtype=np.float64), True)


def _read_setfl(filename, is_adp=False) -> SetFL:
    with open(filename) as fp:
        lines = fp.read().split('\n')
    head = lines[3].split()
    n_el = int(head[0])
    elements = head[1:1 + n_el]
    v = lines[4].split()
    nrho, drho, nr, dr, rcut = int(v[0]), float(v[1]), int(v[2]), float(v[3]), float(v[4])
    tokens: List[str] = []
    for line in lines[5:]:
        tokens.extend(line.split())
    pos = 0
    rho, frho, masses, lattice, ltypes = {}, {}, [], [], []
    for el in elements:
        masses.append(float(tokens[pos + 1]))
        lattice.append(float(tokens[pos + 2]))
        ltypes.append(tokens[pos + 3])
        pos += 4
        frho[el] = _spline(nrho, drho, np.array(tokens[pos:pos + nrho], dtype=np.float64))
        pos += nrho
        rho[el] = _spline(nr, dr, np.array(tokens[pos:pos + nr], dtype=np.float64))
        pos += nr
    keys = [f"{elements[i]}{elements[j]}" for i in range(n_el) for j in range(i, n_el)]
    # the file lists pairs as (i, j <= i); the reference keys them el_i el_j with
    # j >= i in the same sequence (lammps.py:150-158)
    phi, dipole, quadrupole = {}, {}, {}
    r = np.linspace(0.0, nr * dr, nr, endpoint=False)
    for key in keys:
        y = np.array(tokens[pos:pos + nr], dtype=np.float64)
        pos += nr
        y[1:] = y[1:] / r[1:]
        phi[key] = _spline(nr, dr, y)
    if is_adp:
        for target in (dipole, quadrupole):
            for key in keys:
                target[key] = _spline(nr, dr, np.array(tokens[pos:pos + nr],
                                                       dtype=np.float64))
                pos += nr
    return SetFL(elements=elements, rho=rho, phi=phi, embed=frho, dipole=dipole,
                 quadrupole=quadrupole, nr=nr, dr=dr, nrho=nrho, drho=drho, rcut=rcut,
                 atomic_masses=masses, lattice_constants=lattice, lattice_types=ltypes)


def read_eam_alloy_setfl(filename) -> SetFL:
    return _read_setfl(filename, is_adp=False)


def read_adp_setfl(filename) -> SetFL:
    return _read_setfl(filename, is_adp=True)


def write_setfl(filename, setfl: SetFL, comments=("", "", ""), is_adp=False):
    """Writes the tables back in setfl layout (one value per line, %20.16e), the
    format produced by the reference's export_to_setfl (alloy.py:377-379)."""
    from tensoralloy_b200.atoms import atomic_numbers
    out = [str(c) for c in comments[:3]]
    out.append(f"{len(setfl.elements)} " + " ".join(setfl.elements))
    out.append(f"{setfl.nrho} {setfl.drho:.16e} {setfl.nr} {setfl.dr:.16e} "
               f"{setfl.rcut:.16e}")
    for k, el in enumerate(setfl.elements):
        out.append(f"{atomic_numbers.get(el, 0)} {setfl.atomic_masses[k]:.16e} "
                   f"{setfl.lattice_constants[k]:.16e} {setfl.lattice_types[k]}")
        out.extend(f"{v: 20.16e}" for v in setfl.embed[el].y)
        out.extend(f"{v: 20.16e}" for v in setfl.rho[el].y)
    r = np.linspace(0.0, setfl.nr * setfl.dr, setfl.nr, endpoint=False)
    n = len(setfl.elements)
    keys = [f"{setfl.elements[i]}{setfl.elements[j]}" for i in range(n)
            for j in range(i, n)]
    for key in keys:
        y = setfl.phi[key].y.copy()
        y[1:] = y[1:] * r[1:]
        out.extend(f"{v: 20.16e}" for v in y)
    if is_adp:
        for table in (setfl.dipole, setfl.quadrupole):
            for key in keys:
                out.extend(f"{v: 20.16e}" for v in table[key].y)
    with open(filename, 'w') as fp:
        fp.write("\n".join(out) + "\n")


def write_fs_setfl(filename, elements, nrho, drho, nr, dr, rcut, embed, rho, phi,
                   atomic_masses, lattice_constants, lattice_types, comments=("", "", "")):
    """eam/fs (Finnis-Sinclair) setfl layout, as written by the reference's
    `EamFsNN.export_to_setfl` (fs.py:205-..., atsim `writeSetFLFinnisSinclair`): per element
    a header, F(rho) and then ONE density table per partner element (the density at a centre
    of that element from a neighbour of the partner); then r * phi for the pairs (i, j <= i).
    embed[el] [nrho]; rho[el_centre + el_neighbour] [nr]; phi[sorted pair key] [nr]."""
    from tensoralloy_b200.atoms import atomic_numbers
    out = [str(c) for c in comments[:3]]
    out.append(f"{len(elements)} " + " ".join(elements))
    out.append(f"{nrho} {drho:.16e} {nr} {dr:.16e} {rcut:.16e}")
    for k, el in enumerate(elements):
        out.append(f"{atomic_numbers.get(el, 0)} {atomic_masses[k]:.16e} "
                   f"{lattice_constants[k]:.16e} {lattice_types[k]}")
        out.extend(f"{v: 20.16e}" for v in embed[el])
        for other in elements:
            out.extend(f"{v: 20.16e}" for v in rho[f"{el}{other}"])
    r = np.linspace(0.0, nr * dr, nr, endpoint=False)
    for i in range(len(elements)):
        for j in range(i + 1):
            key = "".join(sorted([elements[i], elements[j]]))
            y = np.asarray(phi[key], dtype=np.float64) * r
            y[0] = 0.0 if not np.isfinite(y[0]) else y[0]
            out.extend(f"{v: 20.16e}" for v in y)
    with open(filename, 'w') as fp:
        fp.write("\n".join(out) + "\n")
