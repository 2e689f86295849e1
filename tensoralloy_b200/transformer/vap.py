"""
VirtualAtomMap: local (ASE) order <-> element-sorted 'GSL' order with a virtual
atom at index 0.  Mirror of the reference's transformer/vap.py:18-197 -- same
attributes and results, but built with vectorised numpy (the reference's Python
loops take seconds at 10^6 atoms).
"""
from collections import Counter
from typing import List

import numpy as np


class VirtualAtomMap:
    REAL_ATOM_START = 1

    def __init__(self, max_occurs: Counter, symbols: List[str]):
        self._symbols = list(symbols)
        self._max_occurs = max_occurs
        self._max_vap_natoms = int(sum(max_occurs.values()) + 1)
        istart = VirtualAtomMap.REAL_ATOM_START
        elements = sorted(max_occurs.keys())
        offsets = np.concatenate(
            ([0], np.cumsum([max_occurs[e] for e in elements])[:-1])).astype(np.int64)
        n = len(self._symbols)
        sym = np.asarray(self._symbols)
        el_index = np.searchsorted(np.asarray(elements), sym)
        # rank of each atom among the atoms of its own element (stable)
        order = np.argsort(el_index, kind='stable')
        sorted_el = el_index[order]
        first = np.concatenate(([True], sorted_el[1:] != sorted_el[:-1]))
        group_start = np.maximum.accumulate(np.where(first, np.arange(n), 0))
        rank = np.empty(n, dtype=np.int64)
        rank[order] = np.arange(n) - group_start
        # l2g[k] = GSL index of local index k (both 1-based, 0 = virtual atom)
        self._l2g = np.concatenate(([0], offsets[el_index] + rank + istart))
        self._g2l = np.full(self._max_vap_natoms, -1, dtype=np.int64)
        self._g2l[self._l2g[1:]] = np.arange(n)
        mask = np.zeros(self._max_vap_natoms, dtype=bool)
        mask[self._l2g[1:]] = True
        self._mask = mask
        self._vap_symbols = None
        self._elements_sorted = elements
        self._identity = bool(np.array_equal(self._l2g, np.arange(n + 1))
                              and self._max_vap_natoms == n + 1)

    # dict views with the reference's names (vap.py:50-51)
    @property
    def local_to_gsl_map(self):
        return _ArrayMap(self._l2g)

    @property
    def gsl_to_local_map(self):
        return _ArrayMap(self._g2l, default=-1)

    @property
    def local_to_gsl_array(self):
        """int64 [N+1] array form of `local_to_gsl_map` (index 0 -> 0)."""
        return self._l2g

    @property
    def is_identity(self):
        """True when GSL order == local order (single-species, no padding)."""
        return self._identity

    @property
    def vap_symbols(self):
        if self._vap_symbols is None:
            out = ["X"]
            for el in self._elements_sorted:
                out.extend([el] * self._max_occurs[el])
            self._vap_symbols = out
        return self._vap_symbols

    @property
    def symbols(self):
        return self._symbols

    @property
    def max_vap_natoms(self):
        return self._max_vap_natoms

    @property
    def max_occurs(self) -> Counter:
        return self._max_occurs

    @property
    def atom_masks(self) -> np.ndarray:
        return self._mask

    def map_array(self, array: np.ndarray, reverse=False):
        """vap.py:94-133."""
        array = np.asarray(array)
        rank = np.ndim(array)
        if rank == 2:
            array = array[np.newaxis, ...]
        elif rank <= 1 or rank > 3:
            raise ValueError("The rank should be 2 or 3")
        n = len(self._symbols)
        if not reverse:
            if array.shape[1] != n:
                shape = (array.shape[0], n, array.shape[2])
                raise ValueError(f"The shape should be {shape}")
            array = np.insert(array, 0, np.asarray(0, dtype=array.dtype), axis=1)
            indices = self._g2l + self.REAL_ATOM_START     # -1 -> 0 (virtual row)
        else:
            indices = self._l2g[1:]
        output = array[:, indices]
        if rank == 2:
            output = np.squeeze(output, axis=0)
        return output

    def map_positions(self, positions, reverse=False):
        return self.map_array(positions, reverse=reverse)

    def map_forces(self, forces, reverse=False):
        return self.map_array(forces, reverse=reverse)

    def reverse_map_hessian(self, hessian: np.ndarray, phonopy_format=False):
        """[Np,3,Np,3] (GSL) -> [3N,3N] or [N,N,3,3] (local); vap.py:143-197."""
        hessian = np.asarray(hessian)
        if hessian.ndim != 4 or hessian.shape[1] != 3 or hessian.shape[3] != 3:
            raise ValueError(
                "The input array should be a 4D matrix of shape [Np, 3, Np, 3]")
        idx = self._l2g[1:]
        n = len(self._symbols)
        h = hessian[idx][:, :, idx, :]          # [N,3,N,3]
        if phonopy_format:
            return np.ascontiguousarray(h.transpose(0, 2, 1, 3))
        return np.ascontiguousarray(h.reshape(n * 3, n * 3))


class _ArrayMap:
    """Read-only dict-like view over an index array (keeps the reference's
    `vap.local_to_gsl_map[i]` spelling without building a Python dict)."""

    def __init__(self, arr, default=None):
        self._arr = arr
        self._default = default

    def __getitem__(self, k):
        return int(self._arr[k])

    def get(self, k, default=None):
        if 0 <= k < len(self._arr):
            return int(self._arr[k])
        return default

    def __len__(self):
        return len(self._arr)

    def items(self):
        return ((i, int(v)) for i, v in enumerate(self._arr))

    def keys(self):
        return range(len(self._arr))

    def values(self):
        return (int(v) for v in self._arr)
