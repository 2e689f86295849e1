"""
Labelled structures from a VASP `vasprun.xml` -- the data format on the input side of the
finite-temperature models, mirror of tensoralloy/io/vasp.py:56-315 (`read_vasp_xml`, itself an
extension of ASE's reader).  Only what the training data needs is parsed: species, and per
ionic step the cell, positions, forces, stress, energies, the electron entropy and the
kinetic energy; k-points, eigenvalues, dipoles and constraints are not read.

Conventions kept from the reference:
  energy        `e_fr_energy` of the ionic step + (e_0_energy - e_fr_energy) of its last SCF
                step = E(sigma -> 0) (vasp.py:196-221); with `finite_temperature=True` the
                internal energy U = F + sigma * S (vasp.py:290-294)
  free_energy   `e_fr_energy` of the ionic step
  eentropy      | -(e_fr_energy - e_wo_entrp) / sigma | of the last SCF step, 0 when
                sigma ~ 0 (vasp.py:211-218)
  etemperature  SIGMA of the INCAR block (eV)
  stress        -0.1 GPa * (kbar tensor), Voigt order xx yy zz yz xz xy (vasp.py:250-258)
  kinetic       `energy/i[@name="kinetic"]` of MD runs
`etemperature` / `eentropy` are stored whenever SIGMA is known (the reference's own test,
io/tests/test_vasp.py:21-28, reads them without `finite_temperature=True`).
"""
import gzip
import xml.etree.ElementTree as ET

import numpy as np

from tensoralloy_b200 import atoms_utils
from tensoralloy_b200.atoms import GPa, Atoms


def _rows(node):
    return np.array([[float(x) for x in v.text.split()] for v in node])


def _energy(node, name):
    return float(node.find(f'i[@name="{name}"]').text)


def read_vasp_xml(filename='vasprun.xml', index=-1, finite_temperature=False):
    """Generator of `Atoms` (labels in `atoms.info`: energy, free_energy, forces, stress).
    `index`: int, slice or list of ints over the ionic steps.  `.gz` files are opened
    transparently."""
    opener = gzip.open if str(filename).endswith('.gz') else open
    with opener(filename, 'rb') as fp:
        root = ET.parse(fp).getroot()
    species = [rc[0].text.strip()
               for rc in root.find("atominfo/array[@name='atoms']/set")]
    natoms = len(species)
    sigma = root.find("incar/i[@name='SIGMA']")
    sigma = float(sigma.text) if sigma is not None else None
    calculations = [c for c in root.findall('calculation') if c.find('energy') is not None]
    if isinstance(index, int):
        steps = [calculations[index]]
    elif isinstance(index, (list, tuple)):
        steps = [calculations[k] for k in index]
    else:
        steps = calculations[index]
    for step in steps:
        lastscf = step.findall('scstep/energy')[-1]
        delta = _energy(lastscf, 'e_0_energy') - _energy(lastscf, 'e_fr_energy')
        eentropy = _energy(lastscf, 'e_fr_energy') - _energy(lastscf, 'e_wo_entrp')
        eentropy = 0.0 if (sigma is None or abs(sigma) < 1e-6) else abs(-eentropy / sigma)
        free_energy = _energy(step.find('energy'), 'e_fr_energy')
        energy = free_energy + delta
        if finite_temperature:
            if sigma is None:
                raise ValueError("For finite temperature calculations `ISIGMA` should be "
                                 "non-zero")
            energy = free_energy + eentropy * sigma
        cell = _rows(step.find('structure/crystal/varray[@name="basis"]'))
        scaled = _rows(step.find('structure/varray[@name="positions"]'))
        if scaled.shape != (natoms, 3):
            raise ValueError("vasprun.xml: positions do not match atominfo")
        atoms = Atoms(species, scaled @ cell, cell, True)
        info = {'energy': energy, 'free_energy': free_energy}
        node = step.find('varray[@name="forces"]')
        if node is not None:
            info['forces'] = _rows(node)
        node = step.find('varray[@name="stress"]')
        if node is not None:
            s = _rows(node) * (-0.1 * GPa)
            info['stress'] = s.reshape(9)[[0, 4, 8, 5, 2, 1]]
        atoms.info = info
        if sigma is not None:
            atoms_utils.set_electron_temperature(atoms, sigma)
            atoms_utils.set_electron_entropy(atoms, eentropy)
        kin = step.find('energy/i[@name="kinetic"]')
        if kin is not None:
            atoms_utils.set_kinetic_energy(atoms, float(kin.text))
        yield atoms
