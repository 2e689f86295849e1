"""
UniversalTransformer: the descriptor interface of the reference
(tensoralloy/transformer/universal.py:236-918, base.py:27-587) on top of the GPU
neighbour lists of libtab200.

What changed under the hood:
  * `ase.neighborlist.neighbor_list` + the Python index-map loops
    (universal.py:46-233) -> one GPU cell-list build (csrc/nbr.cu);
  * the dense padded g[4, T, Nvap, nnl_max] tensors (universal.py:583-620) are
    never materialised: models consume the lists directly.
What is kept: constructor signature, `as_dict`, k-body term bookkeeping, VAP
handling, and -- for callers that want the reference wire format --
`get_np_feed_dict`, which rebuilds the reference's arrays (`g2.v2g_map`,
`g2.ilist`, ...) from the GPU list.
"""
from collections import Counter
from typing import Dict, List

import numpy as np

from tensoralloy_b200 import _lib
from tensoralloy_b200.atoms import chemical_symbols
from tensoralloy_b200.precision import get_float_dtype
from tensoralloy_b200.transformer.vap import VirtualAtomMap
from tensoralloy_b200.utils import (get_elements_from_kbody_term,
                                    get_kbody_terms)


class DeviceFeatures:
    """What crosses host -> device for one structure: the replacement of the
    reference's feed dict (universal.py:728-785)."""

    def __init__(self, atoms, vap, types, nbr, d_pos, cell, volume, pbc):
        self.atoms = atoms
        self.vap = vap
        self.types = types          # int32 [N] host, index into sorted elements
        self.nbr = nbr              # _lib.NeighborList (built)
        self.d_pos = d_pos          # cuda float64 [N,3]
        self.cell = cell            # float64 [3,3] host
        self.volume = volume
        self.pbc = pbc
        self.n_atoms = len(types)


class BatchDeviceFeatures:
    """A batch of independent structures on the device: ONE neighbour handle
    (`tab_nbr_build_batch`) instead of the reference's padded batch tensors
    (BatchUniversalTransformer, universal.py:921-1388)."""

    def __init__(self, images, vaps, types, offsets, nbr, d_pos, cells, volumes,
                 transformer=None):
        self.images = images
        self.vaps = vaps            # None: built on demand by structure()
        self._transformer = transformer
        self.types = types          # int32 [N] host, all structures back to back
        self.offsets = offsets      # int32 [B+1]
        self.nbr = nbr
        self.d_pos = d_pos
        self.cells = cells          # [B,3,3] the structures' own lattices
        self.volumes = volumes      # [B]
        self.n_struct = len(images)
        self.n_atoms = int(offsets[-1])

    def structure(self, s):
        """Per-structure view with the attributes `BasicNN._finalize` reads."""
        lo, hi = int(self.offsets[s]), int(self.offsets[s + 1])
        f = DeviceFeatures.__new__(DeviceFeatures)
        vap = self.vaps[s] if self.vaps is not None else \
            self._transformer.get_vap_transformer(self.images[s])
        f.atoms, f.vap, f.types = self.images[s], vap, self.types[lo:hi]
        f.nbr, f.d_pos, f.cell = None, None, self.cells[s]
        f.volume, f.pbc, f.n_atoms = float(self.volumes[s]), None, hi - lo
        return f


class UniversalTransformer:
    """See module docstring.  Signature: universal.py:243-244."""

    def __init__(self, elements: List[str], rcut, acut=None, angular=False,
                 periodic=True, symmetric=True, use_computed_dists=True):
        for element in elements:
            if element not in chemical_symbols:
                raise ValueError(f"{element} is not a valid chemical symbol!")
        if angular and acut is None:
            acut = rcut
        all_kbody_terms, kbody_terms_for_element, elements = get_kbody_terms(
            elements, angular=angular, symmetric=symmetric)
        self._all_kbody_terms = all_kbody_terms
        self._kbody_terms_for_element = kbody_terms_for_element
        self._max_nr_terms = max(
            len([x for x in kbody_terms_for_element[e]
                 if len(get_elements_from_kbody_term(x)) == 2]) for e in elements)
        self._max_na_terms = max(
            len([x for x in kbody_terms_for_element[e]
                 if len(get_elements_from_kbody_term(x)) == 3]) for e in elements)
        self._rcut = rcut
        self._acut = acut
        self._elements = elements
        self._n_elements = len(elements)
        self._periodic = periodic
        self._angular = angular
        self._symmetric = symmetric
        self._use_computed_dists = use_computed_dists
        self._vap_transformers: Dict[str, VirtualAtomMap] = {}
        self._nbr = None
        self._batch_nbr = None
        self._zlut = None
        self._types_cache = (None, None)

    # -- reference properties (universal.py:323-445) -----------------------
    def as_dict(self) -> Dict:
        return {'class': self.__class__.__name__, 'elements': self._elements,
                'rcut': self._rcut, 'acut': self._acut,
                'angular': self._angular, 'periodic': self._periodic,
                'symmetric': self._symmetric,
                'use_computed_dists': self._use_computed_dists}

    descriptor = property(lambda self: "universal")
    rc = property(lambda self: self._rcut)
    rcut = property(lambda self: self._rcut)
    acut = property(lambda self: self._acut)
    elements = property(lambda self: self._elements)
    n_elements = property(lambda self: self._n_elements)
    periodic = property(lambda self: self._periodic)
    angular = property(lambda self: self._angular)
    symmetric = property(lambda self: self._symmetric)
    use_computed_dists = property(lambda self: self._use_computed_dists)
    all_kbody_terms = property(lambda self: self._all_kbody_terms)
    kbody_terms_for_element = property(lambda self: self._kbody_terms_for_element)
    max_nr_terms = property(lambda self: self._max_nr_terms)
    max_na_terms = property(lambda self: self._max_na_terms)

    # -- VAP (base.py:199-226) ----------------------------------------------
    def get_vap_transformer(self, atoms) -> VirtualAtomMap:
        types = self.get_types(atoms)
        # same cache key semantics as get_chemical_formula(mode='reduce'):
        # the run-length encoding of the symbol sequence
        change = np.flatnonzero(np.diff(types)) + 1
        starts = np.concatenate(([0], change))
        lengths = np.diff(np.concatenate((starts, [len(types)])))
        key = (tuple(types[starts].tolist()), tuple(lengths.tolist()))
        if key not in self._vap_transformers:
            symbols = atoms.get_chemical_symbols()
            counter = Counter(symbols)
            max_occurs = Counter()
            for element in self._elements:
                max_occurs[element] = max(1, counter[element])
            self._vap_transformers[key] = VirtualAtomMap(max_occurs, symbols)
        return self._vap_transformers[key]

    def get_types(self, atoms) -> np.ndarray:
        """Element index (into the sorted element list) of every atom."""
        cached_atoms, cached = self._types_cache
        if cached_atoms is atoms and cached is not None and len(cached) == len(atoms):
            return cached
        numbers = getattr(atoms, 'numbers', None)
        if numbers is not None:
            # through the atomic numbers: no per-atom Python strings (1 M atoms: 0.2 s -> 2 ms)
            lut = self._z_lut()
            numbers = np.asarray(numbers)
            if numbers.size and (numbers.min() < 0 or numbers.max() >= len(lut)):
                raise ValueError("atomic numbers outside the periodic table")
            types = lut[numbers]
            if (types < 0).any():
                bad = sorted({chemical_symbols[z] for z in np.unique(numbers[types < 0])})
                raise ValueError(f"elements {bad} are not supported by this transformer")
            types = types.astype(np.int32)
            self._types_cache = (atoms, types)
            return types
        symbols = np.asarray(atoms.get_chemical_symbols())
        els = np.asarray(self._elements)
        idx = np.searchsorted(els, symbols)
        idx = np.clip(idx, 0, len(els) - 1)
        if not np.array_equal(els[idx], symbols):
            bad = sorted(set(symbols[els[idx] != symbols].tolist()))
            raise ValueError(f"elements {bad} are not supported by this transformer")
        types = idx.astype(np.int32)
        self._types_cache = (atoms, types)
        return types

    # -- device path ----------------------------------------------------------
    def _cell_and_pbc(self, atoms):
        """Binning frame handed to the library: (cell, pbc, origin).  Periodic
        structures use their own cell; non-periodic directions (or a missing
        cell) get a bounding extent with its own origin -- only the binning frame
        changes, S stays 0 along non-periodic directions."""
        cell = np.asarray(atoms.get_cell(complete=True), dtype=np.float64).reshape(3, 3)
        pbc = np.asarray(atoms.get_pbc(), dtype=bool).reshape(3)
        if not self._periodic:
            pbc = np.zeros(3, dtype=bool)
        origin = np.zeros(3)
        if abs(np.linalg.det(cell)) < 1e-12 or not pbc.all():
            cell, origin = _bounding_cell(np.asarray(atoms.positions), cell, pbc,
                                          max(self._rcut, self._acut or 0.0))
        return cell, pbc, origin

    def get_device_features(self, atoms, rc=None) -> DeviceFeatures:
        """Build the GPU neighbour lists of `atoms` (replaces get_feed_dict)."""
        import torch
        if self._nbr is None:
            self._nbr = _lib.NeighborList()
        types = self.get_types(atoms)
        cell, pbc, origin = self._cell_and_pbc(atoms)
        d_pos = torch.as_tensor(np.ascontiguousarray(atoms.positions, dtype=np.float64)
                                ).to('cuda', non_blocking=True)
        d_types = torch.as_tensor(types).to('cuda', non_blocking=True)
        if rc is None:
            rc = max(self._rcut, self._acut) if (self._angular and self._acut) \
                else self._rcut
        self._nbr.build_dd(d_pos, d_types, len(types), cell, origin, pbc, rc)
        real_cell = np.asarray(atoms.get_cell(complete=True), dtype=np.float64)
        return DeviceFeatures(atoms, self.get_vap_transformer(atoms), types,
                              self._nbr, d_pos, real_cell.reshape(3, 3),
                              float(atoms.get_volume()), pbc)

    def get_batch_features(self, images, rc=None, nbr=None) -> BatchDeviceFeatures:
        """Neighbour lists of a list of structures in one device handle: one H2D copy
        of the concatenated positions / types, one `tab_nbr_build_batch`.  The host side
        is vectorised over the batch (no per-atom Python work)."""
        import torch
        if nbr is None:
            if self._batch_nbr is None:
                self._batch_nbr = _lib.NeighborList()
            nbr = self._batch_nbr
        nb = len(images)
        lens = np.fromiter((len(a) for a in images), dtype=np.int64, count=nb)
        offsets = np.zeros(nb + 1, dtype=np.int32)
        np.cumsum(lens, out=offsets[1:])
        numbers = np.concatenate([a.numbers for a in images])
        types = self._z_lut()[numbers]
        if (types < 0).any():
            from tensoralloy_b200.atoms import chemical_symbols
            bad = sorted({chemical_symbols[z] for z in np.unique(numbers[types < 0])})
            raise ValueError(f"elements {bad} are not supported by this transformer")
        pos = np.concatenate([a.positions for a in images]).astype(np.float64, copy=False)
        real_cells = np.stack([np.asarray(a.cell, dtype=np.float64).reshape(3, 3)
                               for a in images])
        pbcs = np.stack([np.asarray(a.pbc, dtype=bool).reshape(3) for a in images])
        if not self._periodic:
            pbcs = np.zeros_like(pbcs)
        det = np.linalg.det(real_cells)
        cells = real_cells.copy()
        rmax = max(self._rcut, self._acut or 0.0)
        for s in np.flatnonzero((np.abs(det) < 1e-12) | ~pbcs.all(axis=1)):
            lo, hi = offsets[s], offsets[s + 1]
            cells[s], _ = _bounding_cell(pos[lo:hi], real_cells[s], pbcs[s], rmax)
        d_pos = torch.as_tensor(np.ascontiguousarray(pos)).to('cuda', non_blocking=True)
        d_types = torch.as_tensor(types).to('cuda', non_blocking=True)
        if rc is None:
            rc = max(self._rcut, self._acut) if (self._angular and self._acut) \
                else self._rcut
        nbr.build_batch(d_pos, d_types, offsets, cells, pbcs, rc)
        return BatchDeviceFeatures(list(images), None, types, offsets, nbr, d_pos,
                                   real_cells, np.abs(det), transformer=self)

    def _z_lut(self):
        """atomic number -> element index of this transformer (-1 = unsupported)."""
        if self._zlut is None:
            from tensoralloy_b200.atoms import atomic_numbers
            lut = np.full(128, -1, dtype=np.int32)
            for k, el in enumerate(self._elements):
                lut[atomic_numbers[el]] = k
            self._zlut = lut
        return self._zlut

    # -- reference wire format (universal.py:46-233,728-785,851-893) ----------------
    def get_np_feed_dict(self, atoms):
        """The reference's PREDICT-mode feed dict, rebuilt from the GPU list: every key of
        universal.py:851-893 incl. the angular ones (`g4.v2g_map`, `g4.ilist / jlist / klist`,
        `g4.n1 / n2 / n3`, `ij2k_max`).  Slots follow the canonical pair order
        (transformer/wire_format.py)."""
        import torch
        from tensoralloy_b200.transformer import wire_format as wf
        np_dtype = get_float_dtype().as_numpy_dtype
        feats = self.get_device_features(atoms)
        vap = feats.vap
        i, j, S = feats.nbr.export()
        torch.cuda.synchronize()
        return self._feed_dict_from_list(atoms, vap, feats.types, feats.cell, feats.volume,
                                         i.cpu().numpy(), j.cpu().numpy(), S.cpu().numpy(),
                                         np_dtype)

    def _feed_dict_from_list(self, atoms, vap, types, cell, volume, i, j, S, np_dtype):
        """Host part of `get_np_feed_dict` (pure numpy; the CPU tests drive it with the
        oracle's neighbour list)."""
        from tensoralloy_b200.transformer import wire_format as wf
        pos = np.asarray(atoms.positions, dtype=np.float64)
        arr = wf.build_feed_arrays(i, j, S, pos, cell, types, vap.local_to_gsl_array,
                                   self._elements, self._kbody_terms_for_element,
                                   self._rcut, self._acut, self._angular, self._symmetric)
        v2g = arr["g2.v2g_map"]
        feed = dict()
        if self._use_computed_dists:
            feed["positions"] = vap.map_positions(pos).astype(np_dtype)
            feed["cell"] = np.asarray(cell, dtype=np.float64).astype(np_dtype)
            feed["volume"] = np_dtype(volume)
        feed["n_atoms_vap"] = np.int32(vap.max_vap_natoms)
        feed["nnl_max"] = np.int32(v2g[:, 2].max() + 1 if len(v2g) else 0)
        feed["atom_masks"] = vap.atom_masks.astype(np_dtype)
        feed["etemperature"] = np_dtype(atoms.info.get('etemperature', 0.0))
        feed["row_splits"] = np.int32([1] + [vap.max_occurs[e] for e in self._elements])
        feed["g2.v2g_map"] = v2g
        if self._use_computed_dists:
            feed["g2.ilist"] = arr["g2.ilist"]
            feed["g2.jlist"] = arr["g2.jlist"]
            feed["g2.n1"] = arr["g2.n1"].astype(np_dtype)
        else:
            feed["g2.rij"] = np.concatenate((arr["g2.d"][:, None], arr["g2.D"]),
                                            axis=1).T.astype(np_dtype)
        if self._angular:
            g4 = arr["g4.v2g_map"]
            feed["ij2k_max"] = np.int32(g4[:, 3].max() + 1 if len(g4) else 0)
            feed["g4.v2g_map"] = g4
            if self._use_computed_dists:
                for key in ("g4.ilist", "g4.jlist", "g4.klist"):
                    feed[key] = arr[key]
                for key in ("g4.n1", "g4.n2", "g4.n3"):
                    feed[key] = arr[key].astype(np_dtype)
            else:
                raise NotImplementedError(
                    "g4.rijk (use_computed_dists=False with angular terms)")
        return feed

    def get_feed_dict(self, atoms):
        """Kept for API compatibility: there are no TF placeholders here, the
        numpy feed dict is returned keyed by name."""
        return self.get_np_feed_dict(atoms)

    def get_constant_features(self, atoms):
        """universal.py:910-918: in this build the 'features' handed to
        `BasicNN.build` are the device-side lists."""
        return self.get_device_features(atoms)


def _bounding_cell(positions, cell, pbc, rc):
    """Frame for structures that are not periodic in every direction.  Along a
    periodic direction the lattice vector is kept; along a non-periodic one the
    frame spans the atoms (plus a margin) starting at `origin`."""
    cell = np.array(cell, dtype=np.float64)
    pos = np.asarray(positions, dtype=np.float64)
    out = cell.copy()
    origin = np.zeros(3)
    if not pbc.any():
        lo = pos.min(axis=0) - 0.5
        hi = pos.max(axis=0) + 0.5
        return np.diag(np.maximum(hi - lo, 1.0)), lo
    for k in range(3):
        if pbc[k]:
            continue
        v = cell[k]
        if not np.any(v):
            others = [cell[m] for m in range(3) if m != k and np.any(cell[m])]
            v = np.cross(others[0], others[1]) if len(others) == 2 else np.eye(3)[k]
        v = v / np.linalg.norm(v)
        proj = pos @ v
        lo, hi = proj.min() - 0.5, proj.max() + 0.5
        out[k] = v * max(hi - lo, 1.0)
        origin = origin + v * lo
    return out, origin


class BatchUniversalTransformer(UniversalTransformer):
    """Mirror of the reference's mini-batch transformer (transformer/universal.py:921-1388):
    same constructor, `as_dict`, size properties and `as_descriptor_transformer`.

    The reference needs the maxima (`max_occurs`, `nij_max`, `nnl_max`, ...) to PAD every
    structure of a batch to `[B, N + 1, 3]` / `[B, nij_max, .]` tensors.  Here a batch is one
    extended atom array with one sliced neighbour table (`get_batch_features`,
    csrc/nbr_batch.cuh) and nothing is padded; the maxima are kept as the CONTRACT of the
    dataset: `get_batch_features` refuses a batch larger than `batch_size`, a structure that
    exceeds `max_occurs` and one with more than `nij_max` pairs (the reference fails late, in
    `scatter_nd` or in the VirtualAtomMap).  The TFRecord codec (`encode`, `decode_protobuf`,
    universal.py:1205-1330) writes and reads the reference's record layout through
    transformer/tfrecord.py (no TensorFlow): the padded TRAIN-mode maps exist only there."""

    def __init__(self, max_occurs, rcut, acut=None, angular=False, periodic=True,
                 symmetric=True, nij_max=None, nijk_max=None, nnl_max=None, ij2k_max=None,
                 batch_size=None, use_forces=True, use_stress=False):
        max_occurs = Counter({k: int(v) for k, v in dict(max_occurs).items()})
        super().__init__(elements=sorted(max_occurs.keys()), rcut=rcut, acut=acut,
                         angular=angular, periodic=periodic, symmetric=symmetric,
                         use_computed_dists=True)
        self._nij_max = nij_max
        self._nijk_max = nijk_max
        self._nnl_max = nnl_max
        self._ij2k_max = ij2k_max
        self._batch_size = batch_size
        self._use_forces = use_forces
        self._use_stress = use_stress
        self._max_occurs = max_occurs
        self._max_n_atoms = sum(max_occurs.values())

    def as_dict(self) -> Dict:
        return {'class': self.__class__.__name__, 'max_occurs': dict(self._max_occurs),
                'rcut': self._rcut, 'acut': self._acut, 'angular': self._angular,
                'nij_max': self._nij_max, 'nijk_max': self._nijk_max,
                'nnl_max': self._nnl_max, 'ij2k_max': self._ij2k_max,
                'batch_size': self._batch_size, 'use_forces': self._use_forces,
                'use_stress': self._use_stress}

    batch_size = property(lambda self: self._batch_size)
    nij_max = property(lambda self: self._nij_max)
    nijk_max = property(lambda self: self._nijk_max)
    nnl_max = property(lambda self: self._nnl_max)
    ij2k_max = property(lambda self: self._ij2k_max)
    max_occurs = property(lambda self: self._max_occurs)
    max_n_atoms = property(lambda self: self._max_n_atoms)
    use_forces = property(lambda self: self._use_forces)
    use_stress = property(lambda self: self._use_stress)

    def as_descriptor_transformer(self) -> UniversalTransformer:
        """universal.py:1040-1048."""
        return UniversalTransformer(elements=sorted(self._max_occurs.keys()),
                                    rcut=self._rcut, acut=self._acut, angular=self._angular,
                                    periodic=self._periodic, symmetric=self._symmetric)

    def get_g_shape(self, features=None, angular=False):
        """Shape of the reference's padded descriptor tensor (universal.py:1112-1127)."""
        if angular:
            return [self._batch_size, self._max_na_terms, self._max_n_atoms + 1,
                    self._nnl_max, self._ij2k_max]
        return [self._batch_size, self._max_nr_terms, self._max_n_atoms + 1,
                self._nnl_max, 1]

    @staticmethod
    def get_row_split_axis():
        return 2

    def get_row_split_sizes(self, _=None):
        """universal.py:1137-1144: the virtual atom, then `max_occurs` per element."""
        return [1] + [self._max_occurs[e] for e in self._elements]

    def check_occurs(self, images):
        """Every structure must fit the declared `max_occurs` (the reference's
        VirtualAtomMap asserts the same, vap.py:52-60)."""
        lut = self._z_lut()
        n_el = len(self._elements)
        limit = np.array([self._max_occurs[e] for e in self._elements])
        for k, atoms in enumerate(images):
            types = lut[np.asarray(atoms.numbers)]
            if (types < 0).any():
                continue            # reported by get_batch_features with the symbols
            counts = np.bincount(types, minlength=n_el)
            if (counts > limit).any():
                e = self._elements[int(np.argmax(counts > limit))]
                raise ValueError(f"structure {k}: {counts[self._elements.index(e)]} atoms of "
                                 f"{e} exceed max_occurs[{e}] = {self._max_occurs[e]}")

    def get_batch_features(self, images, rc=None, nbr=None) -> BatchDeviceFeatures:
        if self._batch_size is not None and len(images) > self._batch_size:
            raise ValueError(f"{len(images)} structures exceed batch_size = "
                             f"{self._batch_size}")
        self.check_occurs(images)
        feats = super().get_batch_features(images, rc=rc, nbr=nbr)
        if self._nij_max is not None:
            # pairs per structure from the per-atom row lengths (caller order)
            counts = feats.nbr.counts().cpu().numpy().astype(np.int64)
            csum = np.concatenate(([0], np.cumsum(counts)))
            nij = csum[feats.offsets[1:]] - csum[feats.offsets[:-1]]
            if nij.size and int(nij.max()) > self._nij_max:
                raise ValueError(f"structure {int(nij.argmax())}: nij = {int(nij.max())} "
                                 f"exceeds nij_max = {self._nij_max}")
        return feats

    # -- TFRecord codec (universal.py:1177-1330, base.py:365-437) ---------------------
    def get_dataset_vap(self, atoms) -> VirtualAtomMap:
        """The map for the DATASET's `max_occurs` (the reference's batch-mode
        `get_vap_transformer`, base.py:206-226), cached by the run-length encoding of the symbol
        sequence.  (`get_vap_transformer` stays the per-structure map: evaluation results of
        this package are never padded.)"""
        key = ('batch', atoms.get_chemical_formula(mode='reduce'))
        if key not in self._vap_transformers:
            self._vap_transformers[key] = VirtualAtomMap(self._max_occurs,
                                                         atoms.get_chemical_symbols())
        return self._vap_transformers[key]

    def _neighbor_list(self, atoms, rc):
        """(i, j, S) of `atoms` within rc from the GPU list builder."""
        import torch
        saved = self._types_cache
        feats = UniversalTransformer.get_device_features(self, atoms, rc=rc)
        i, j, S = feats.nbr.export()
        torch.cuda.synchronize()
        self._types_cache = saved
        return i.cpu().numpy(), j.cpu().numpy(), S.cpu().numpy()

    def get_metadata(self, atoms, vap=None, neighbor_list=None):
        """TRAIN-mode metadata of one structure (universal.py:1050-1086): the maps carry a
        leading batch-index column (6 columns; filled by the batching step) and every array is
        padded with zero rows to `nij_max` / `nijk_max`.  `neighbor_list` = (i, j, S) within
        max(rcut, acut) replaces the GPU list builder (CPU tests)."""
        from tensoralloy_b200.transformer import wire_format as wf
        np_dtype = get_float_dtype().as_numpy_dtype
        vap = vap or self.get_dataset_vap(atoms)
        rmax = max(self._rcut, self._acut) if (self._angular and self._acut) else self._rcut
        i, j, S = neighbor_list if neighbor_list is not None else \
            self._neighbor_list(atoms, rmax)
        types = self._z_lut()[np.asarray(atoms.numbers)]
        arr = wf.build_feed_arrays(i, j, S, np.asarray(atoms.positions, dtype=np.float64),
                                   np.asarray(atoms.get_cell(complete=True), dtype=np.float64),
                                   types, vap.local_to_gsl_array, self._elements,
                                   self._kbody_terms_for_element, self._rcut, self._acut,
                                   self._angular, self._symmetric)

        def padded(a, n_max, what):
            n = a.shape[0]
            n_max = n if n_max is None else int(n_max)
            if n > n_max:
                raise ValueError(f"{what} = {n} exceeds {what}_max = {n_max}")
            out = np.zeros((n_max,) + a.shape[1:], dtype=a.dtype)
            out[:n] = a
            return out

        def train_map(v2g, n_max, what):
            out = np.zeros((v2g.shape[0], 6), dtype=np.int32)
            out[:, 1:] = v2g
            return padded(out, n_max, what)

        meta = {"g2.v2g_map": train_map(arr["g2.v2g_map"], self._nij_max, 'nij'),
                "g2.ilist": padded(arr["g2.ilist"], self._nij_max, 'nij'),
                "g2.jlist": padded(arr["g2.jlist"], self._nij_max, 'nij'),
                "g2.n1": padded(arr["g2.n1"].astype(np_dtype), self._nij_max, 'nij')}
        if self._angular:
            meta["g4.v2g_map"] = train_map(arr["g4.v2g_map"], self._nijk_max, 'nijk')
            for key in ("g4.ilist", "g4.jlist", "g4.klist"):
                meta[key] = padded(arr[key], self._nijk_max, 'nijk')
            for key in ("g4.n1", "g4.n2", "g4.n3"):
                meta[key] = padded(arr[key].astype(np_dtype), self._nijk_max, 'nijk')
        return meta

    def encode(self, atoms, neighbor_list=None):
        """One structure -> `Example` (tfrecord.py; `.SerializeToString()` is the record the
        reference's `tf.train.Example` yields): universal.py:1219-1230 + base.py:383-437.
        Labels come from `atoms.info` ('energy', 'forces', 'stress' in eV/A^3 Voigt,
        'etemperature', 'eentropy'), missing ones count as zero (the reference attaches a
        zero `SinglePointCalculator`)."""
        from tensoralloy_b200.transformer.tfrecord import Example, bytes_feature, int64_feature
        from tensoralloy_b200.io.units import _UNITS
        GPa = _UNITS['GPa']
        np_dtype = get_float_dtype().as_numpy_dtype
        vap = self.get_dataset_vap(atoms)
        info = getattr(atoms, 'info', {})
        n = len(atoms)

        def raw(a):
            return bytes_feature(np.ascontiguousarray(a, dtype=np_dtype).tobytes())

        energy = np.atleast_1d(float(info.get('energy', 0.0))).astype(np_dtype)
        etemp = np.atleast_1d(float(info.get('etemperature', 0.0))).astype(np_dtype)
        eentropy = np.atleast_1d(float(info.get('eentropy', 0.0))).astype(np_dtype)
        feats = {
            'positions': raw(vap.map_positions(np.asarray(atoms.positions))),
            'cell': raw(atoms.get_cell(complete=True)),
            'n_atoms_vap': int64_feature(n),
            'volume': raw(np.atleast_1d(atoms.get_volume())),
            'energy': raw(energy),
            'free_energy': raw(energy - etemp * eentropy),
            'atom_masks': raw(vap.atom_masks),
            'eentropy': raw(eentropy),
            'etemperature': raw(etemp),
        }
        if self._use_forces:
            forces = np.asarray(info.get('forces', np.zeros((n, 3))), dtype=np.float64)
            feats['forces'] = raw(vap.map_forces(forces.reshape(n, 3)))
        if self._use_stress:
            stress = np.asarray(info.get('stress', np.zeros(6)), dtype=np.float64).reshape(6)
            stress = stress.astype(np_dtype)
            feats['stress'] = raw(stress)
            feats['total_pressure'] = raw(np.atleast_1d(-stress[:3].mean() / GPa))
        meta = self.get_metadata(atoms, vap=vap, neighbor_list=neighbor_list)
        g2 = np.concatenate((meta["g2.v2g_map"], meta["g2.ilist"][:, None],
                             meta["g2.jlist"][:, None]), axis=1).astype(np.int32)
        feats['g2.indices'] = bytes_feature(g2.tobytes())
        feats['g2.shifts'] = bytes_feature(meta["g2.n1"].tobytes())
        if self._angular:
            g4 = np.concatenate((meta["g4.v2g_map"], meta["g4.ilist"][:, None],
                                 meta["g4.jlist"][:, None], meta["g4.klist"][:, None]),
                                axis=1).astype(np.int32)
            feats['g4.indices'] = bytes_feature(g4.tobytes())
            feats['g4.shifts'] = bytes_feature(np.concatenate(
                (meta["g4.n1"], meta["g4.n2"], meta["g4.n3"]), axis=1).tobytes())
        return Example(feats)

    def decode_protobuf(self, example_proto):
        """Serialised `Example` (bytes) -> the dict of arrays of universal.py:1232-1319 (numpy
        instead of tf.Tensor).  The sizes are the transformer's (`max_n_atoms`, `nij_max`,
        `nijk_max`): a record of another size raises, as `set_shape` does."""
        from tensoralloy_b200.transformer.tfrecord import Example
        if self._nij_max is None or (self._angular and self._nijk_max is None):
            raise ValueError("decode_protobuf needs nij_max (and nijk_max for angular "
                             "transformers)")
        np_dtype = get_float_dtype().as_numpy_dtype
        feats = Example.FromString(example_proto).features
        n1 = self._max_n_atoms + 1

        def raw(key, shape, dtype=np_dtype):
            if key not in feats or feats[key].kind != 'bytes_list' or \
                    len(feats[key].value) != 1:
                raise KeyError(f"feature '{key}' is missing from the example")
            a = np.frombuffer(feats[key].value[0], dtype=dtype)
            if a.size != int(np.prod(shape, dtype=np.int64)):
                raise ValueError(f"feature '{key}': {a.size} values, expected shape "
                                 f"{tuple(shape)}")
            return a.reshape(shape).copy()

        out = {
            'positions': raw('positions', (n1, 3)),
            'n_atoms_vap': np.int64(feats['n_atoms_vap'].value[0]),
            'energy': raw('energy', (1,))[0],
            'cell': raw('cell', (3, 3)),
            'volume': raw('volume', (1,))[0],
            'atom_masks': raw('atom_masks', (n1,)),
            'etemperature': raw('etemperature', (1,))[0],
            'eentropy': raw('eentropy', (1,))[0],
            'free_energy': raw('free_energy', (1,))[0],
        }
        if self._use_forces:
            out['forces'] = raw('forces', (n1, 3))
        if self._use_stress:
            out['stress'] = raw('stress', (6,))
            out['total_pressure'] = raw('total_pressure', (1,))[0]
        g2 = raw('g2.indices', (self._nij_max, 8), np.int32)
        out['g2.v2g_map'] = g2[:, :6]
        out['g2.ilist'] = g2[:, 6]
        out['g2.jlist'] = g2[:, 7]
        out['g2.n1'] = raw('g2.shifts', (self._nij_max, 3))
        if self._angular:
            g4 = raw('g4.indices', (self._nijk_max, 9), np.int32)
            out['g4.v2g_map'] = g4[:, :6]
            out['g4.ilist'], out['g4.jlist'], out['g4.klist'] = g4[:, 6], g4[:, 7], g4[:, 8]
            sh = raw('g4.shifts', (self._nijk_max, 9))
            out['g4.n1'], out['g4.n2'], out['g4.n3'] = sh[:, 0:3], sh[:, 3:6], sh[:, 6:9]
        return out

    def decode_atoms(self, decoded):
        """A decoded example back to a labelled `Atoms` (GSL order undone): what the trainers
        of this package consume (`add_structure(atoms, energy, forces, stress)`).  The element
        of every real atom follows from its GSL slot (row splits = `max_occurs` per element in
        sorted order)."""
        from tensoralloy_b200.atoms import Atoms
        mask = decoded['atom_masks'] > 0
        symbols_gsl = [None]
        for e in self._elements:
            symbols_gsl.extend([e] * self._max_occurs[e])
        rows = np.flatnonzero(mask)
        info = {'energy': float(decoded['energy']),
                'etemperature': float(decoded['etemperature']),
                'eentropy': float(decoded['eentropy'])}
        if 'forces' in decoded:
            info['forces'] = np.asarray(decoded['forces'], dtype=np.float64)[rows]
        if 'stress' in decoded:
            info['stress'] = np.asarray(decoded['stress'], dtype=np.float64)
        cell = np.asarray(decoded['cell'], dtype=np.float64)
        return Atoms([symbols_gsl[r] for r in rows],
                     np.asarray(decoded['positions'], dtype=np.float64)[rows], cell,
                     pbc=self._periodic and abs(np.linalg.det(cell)) > 1e-12, info=info)
