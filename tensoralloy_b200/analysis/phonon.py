"""
PhononCalculator -- the Hessian caller of the reference (tensoralloy/analysis/phonon.py:
283-592): vibrational frequencies from the analytic GPU Hessian
(`tab_eam_hessian`, csrc/hessian.cu) and phonopy-format force constants.

  get_frequencies_and_normal_modes   phonon.py:288-320
  get_force_constants                the analytic branch of get_phonon_spectrum
                                     (phonon.py:519-528: Hessian of the supercell ->
                                     VirtualAtomMap.reverse_map_hessian(phonopy_format=True))
  get_phonon_frequencies             the band-structure numbers phonopy would produce from
                                     those force constants (dynamical matrix with phonopy's
                                     shortest-vector multiplicities); phonopy / seekpath are
                                     not installed in this image, so plotting and automatic
                                     band paths stay with phonopy (get_phonon_spectrum raises
                                     ImportError without it).
"""
import numpy as np

from tensoralloy_b200.atoms import atomic_numbers
from tensoralloy_b200.calculator import TensorAlloyCalculator

# phonopy.units (CODATA as shipped with phonopy 2.x): sqrt(eV / (amu A^2)) / 2 pi in THz,
# and THz -> cm^-1
VaspToTHz = 15.633302
THzToCm = 33.356410
VaspToCm = VaspToTHz * THzToCm

# standard atomic weights (ase.data.atomic_masses, IUPAC 2016) of the elements the
# reference fixtures use; others must be passed explicitly
_MASSES = {'H': 1.008, 'Li': 6.94, 'Be': 9.0121831, 'B': 10.81, 'C': 12.011, 'N': 14.007,
           'O': 15.999, 'Mg': 24.305, 'Al': 26.9815385, 'Si': 28.085, 'Ti': 47.867,
           'V': 50.9415, 'Cr': 51.9961, 'Fe': 55.845, 'Co': 58.933194, 'Ni': 58.6934,
           'Cu': 63.546, 'Zr': 91.224, 'Nb': 92.90637, 'Mo': 95.95, 'Pd': 106.42,
           'Ag': 107.8682, 'Ta': 180.94788, 'W': 183.84, 'Pt': 195.084, 'Au': 196.966569,
           'Pu': 244.0}


def get_masses(atoms):
    if hasattr(atoms, 'get_masses'):
        try:
            return np.asarray(atoms.get_masses(), dtype=np.float64)
        except Exception:      # the stand-in Atoms has no mass table
            pass
    try:
        return np.array([_MASSES[s] for s in atoms.get_chemical_symbols()])
    except KeyError as exc:
        raise ValueError(f"no atomic mass for element {exc}") from None


class PhononCalculator(TensorAlloyCalculator):
    """A `TensorAlloyCalculator` with the phonon helpers of the reference."""

    def get_frequencies_and_normal_modes(self, atoms=None, masses=None):
        """phonon.py:288-320.  NB the reference multiplies the EIGENVALUES of the
        mass-weighted Hessian by VaspToCm (no square root, :318-319); the same numbers are
        returned here.  `get_phonon_frequencies` returns physical frequencies."""
        atoms = atoms if atoms is not None else self.atoms
        hessian = self.get_hessian(atoms)
        m = get_masses(atoms) if masses is None else np.asarray(masses, dtype=np.float64)
        inv = np.repeat(m ** -0.5, 3)
        mwh = hessian * inv[:, None] * inv[None, :]
        eigvals, eigvecs = np.linalg.eigh(mwh)
        return eigvals * VaspToCm, eigvecs.transpose()

    def get_force_constants(self, atoms, supercell=(1, 1, 1)):
        """Force constants [N, N, 3, 3] (phonopy layout, eV/A^2) of `atoms * supercell`
        from ONE analytic Hessian evaluation (phonon.py:514-528).  Returns (fc, supercell
        atoms); atom order = ase `Atoms.__mul__` order."""
        if not np.all(np.asarray(atoms.pbc)):
            raise ValueError("The PBC of `atoms` are all False.")
        if 'hessian' not in self.predict_properties:
            raise ValueError("this model cannot predict hessians.")
        sup = atoms * tuple(int(x) for x in supercell)
        self.calculate(sup, properties=['hessian'])
        vap = self.transformer.get_vap_transformer(sup)
        fc = vap.reverse_map_hessian(self.results['hessian'], phonopy_format=True)
        return fc.astype(np.float64), sup

    def get_phonon_frequencies(self, atoms, qpoints, supercell=(4, 4, 4), masses=None,
                               use_wavenumber=False):
        """Frequencies [nq, 3 n_prim] (THz, or cm^-1) at `qpoints` (reduced coordinates of
        the reciprocal cell of `atoms`, which is taken as the primitive cell) from the
        force constants of `atoms * supercell`.  Dynamical matrix as phonopy builds it:
            D_ab(q) = 1/sqrt(m_a m_b) sum_{b' in images of b} Phi(a, b') <exp(i q.r)>
        where <.> averages over the equivalent shortest vectors between a and b' across
        supercell images.  Imaginary modes are returned as negative numbers."""
        fc, sup = self.get_force_constants(atoms, supercell)
        n_prim = len(atoms)
        reps = np.asarray(supercell, dtype=int)
        n_cells = int(np.prod(reps))
        m = get_masses(atoms) if masses is None else np.asarray(masses, dtype=np.float64)
        cell = np.asarray(atoms.cell, dtype=np.float64)
        scell = np.asarray(sup.cell, dtype=np.float64)
        # ase repeat order: image-major, atoms of the unit cell inside; image 0 = home cell
        pos = np.asarray(sup.positions)
        home = pos[:n_prim]
        # candidate lattice translations of the supercell for the shortest-vector search
        tr = np.array([[i, j, k] for i in (-1, 0, 1) for j in (-1, 0, 1)
                       for k in (-1, 0, 1)], dtype=np.float64) @ scell
        qcart = 2.0 * np.pi * np.asarray(qpoints, dtype=np.float64) @ np.linalg.inv(cell).T
        dyn = np.zeros((len(qcart), 3 * n_prim, 3 * n_prim), dtype=np.complex128)
        for a in range(n_prim):
            d = pos[None, :, :] + tr[:, None, :] - home[a][None, None, :]   # [27, N, 3]
            dist = np.linalg.norm(d, axis=2)
            dmin = dist.min(axis=0)
            mask = dist <= dmin[None, :] + 1e-5                            # equivalent images
            mult = mask.sum(axis=0)
            for qi, q in enumerate(qcart):
                phase = (np.exp(1j * (d @ q)) * mask).sum(axis=0) / mult      # [N]
                for b in range(n_prim):
                    sel = np.arange(b, n_prim * n_cells, n_prim)
                    blk = np.tensordot(phase[sel], fc[a, sel], axes=(0, 0))
                    dyn[qi, 3 * a:3 * a + 3, 3 * b:3 * b + 3] = blk / np.sqrt(m[a] * m[b])
        dyn = 0.5 * (dyn + dyn.conj().transpose(0, 2, 1))
        ev = np.linalg.eigvalsh(dyn)
        freq = np.sign(ev) * np.sqrt(np.abs(ev)) * VaspToTHz
        return freq * THzToCm if use_wavenumber else freq

    def get_phonon_spectrum(self, *args, **kwargs):
        """phonon.py:428-592 plots the band structure through phonopy (+ seekpath)."""
        try:
            import phonopy  # noqa: F401
        except ImportError as exc:
            raise ImportError(
                "get_phonon_spectrum draws the bands with phonopy, which is not installed; "
                "get_force_constants / get_phonon_frequencies provide the numbers") from exc
        raise NotImplementedError("phonopy plotting is delegated to the reference tool")
