"""
Neighbour sizes (`nij`, `nnl`, `nijk`, `ij2k`) of a structure -- mirror of the
reference's tensoralloy/neighbor.py:34-146, computed from the GPU neighbour lists
(the reference calls ASE's neighbor_list and loops in Python).

  nij   number of directed pairs within rc
  nnl   max over atoms and neighbour species of the per-species neighbour count
        (neighbor.py:88-95)
  nijk  sum_i n_i (n_i - 1) / 2                                  (neighbor.py:121,141)
  ij2k  max over (centre, neighbour entry j, species k) of the number of OTHER
        row entries of species k                                  (neighbor.py:122-135)
"""
from dataclasses import dataclass
from enum import Enum
from typing import Union

import numpy as np


class NeighborProperty(Enum):
    nij = 0
    nnl = 1
    nijk = 2
    ij2k = 3


@dataclass(frozen=True)
class NeighborSize:
    nnl: int
    nij: int
    nijk: int
    ij2k: int

    def __getitem__(self, item: Union[str, NeighborProperty]):
        if isinstance(item, NeighborProperty):
            item = item.name
        return self.__dict__[item]


def find_neighbor_size_of_atoms(atoms, rc: float, find_nijk=False,
                                find_ij2k=False) -> NeighborSize:
    import torch
    from tensoralloy_b200.transformer import UniversalTransformer
    elements = sorted(set(atoms.get_chemical_symbols()))
    clf = UniversalTransformer(elements, rcut=rc)
    feats = clf.get_device_features(atoms)
    i, j, _ = feats.nbr.export()
    torch.cuda.synchronize()
    i = i.cpu().numpy().astype(np.int64)
    j = j.cpu().numpy().astype(np.int64)
    types = feats.types.astype(np.int64)
    n, nel = len(atoms), len(elements)
    nij = len(i)
    tc = np.zeros((n, nel), dtype=np.int64)
    np.add.at(tc, (i, types[j]), 1)
    nnl = int(tc.max()) if nij else 0
    nijk = ij2k = 0
    if find_ij2k or find_nijk:
        cnt = tc.sum(axis=1)
        nijk = int((cnt * (cnt - 1) // 2).sum())
    if find_ij2k and nij:
        # for a neighbour entry of species tj: others of species tk = tc[tk] - (tj == tk)
        present = tc > 0
        best = 0
        for tj in range(nel):
            rows = present[:, tj]
            if not rows.any():
                continue
            other = tc[rows].copy()
            other[:, tj] -= 1
            best = max(best, int(other.max()))
        ij2k = best
    return NeighborSize(nnl=nnl, nij=nij, nijk=nijk, ij2k=ij2k)
