"""
The reference's feed-dict "wire format" (transformer/universal.py:46-233, 728-785, 851-893)
rebuilt from a neighbour list: `g2.v2g_map / ilist / jlist / n1`, and for angular transformers
`g4.v2g_map / ilist / jlist / klist / n1 / n2 / n3`, `ij2k_max`.

Vectorised numpy (the reference: Python loops over pairs and triples).  The slot numbers the
reference assigns (`v2g_map[:, 2]`, `[:, 3]`) follow the order of ASE's neighbour list INSIDE a
row, which no reference test pins; here the list is first brought into the canonical order
(i, j, Sx, Sy, Sz) and the slots follow that order.  tests/test_wire_format.py compares with the
loop restatement in oracle/wire_format.py on the same canonical list.
"""
from typing import Dict, List

import numpy as np

from tensoralloy_b200.utils import get_elements_from_kbody_term


def canonical_order(i, j, S):
    """Permutation that sorts the directed pairs by (i, j, Sx, Sy, Sz)."""
    return np.lexsort((S[:, 2], S[:, 1], S[:, 0], j, i))


def _running_count(keys):
    """out[p] = number of earlier entries with the same key (list order kept)."""
    n = len(keys)
    if n == 0:
        return np.zeros(0, dtype=np.int32)
    order = np.argsort(keys, kind='stable')
    ks = keys[order]
    first = np.concatenate(([True], ks[1:] != ks[:-1]))
    start = np.maximum.accumulate(np.where(first, np.arange(n), 0))
    out = np.empty(n, dtype=np.int32)
    out[order] = (np.arange(n) - start).astype(np.int32)
    return out


def interaction_tables(elements: List[str], kbody_terms_for_element: Dict[str, List[str]]):
    """radial[c, a] = index of the term c-a among the terms of centre c; angular[c, a, b] =
    index of the term c-a-b minus the number of elements (universal.py:807-817); -1 = the
    transformer has no such term (symmetric transformers only hold a <= b)."""
    n = len(elements)
    radial = np.full((n, n), -1, dtype=np.int32)
    angular = np.full((n, n, n), -1, dtype=np.int32)
    pos = {e: k for k, e in enumerate(elements)}
    for c, element in enumerate(elements):
        for idx, term in enumerate(kbody_terms_for_element[element]):
            parts = get_elements_from_kbody_term(term)
            if len(parts) == 2:
                radial[c, pos[parts[1]]] = idx
            else:
                angular[c, pos[parts[1]], pos[parts[2]]] = idx - n
    return radial, angular


def radial_metadata(i, j, types, l2g, radial):
    """(v2g_map [nij,5], ilist, jlist) of a canonically ordered list (i, j local, 0-based):
    universal.py:46-112 in PREDICT mode (iaxis = 0)."""
    nij = len(i)
    tlist = radial[types[i], types[j]].astype(np.int32)
    ilist = l2g[i + 1].astype(np.int32)
    jlist = l2g[j + 1].astype(np.int32)
    n_terms = int(radial.max()) + 1
    inc = _running_count(ilist.astype(np.int64) * n_terms + tlist)
    v2g = np.zeros((nij, 5), dtype=np.int32)
    v2g[:, 0] = tlist
    v2g[:, 1] = ilist
    v2g[:, 2] = inc
    v2g[:, 4] = 1
    return v2g, ilist, jlist


def angular_metadata(i, j, S, types, ilist, jlist, inc, angular, symmetric=True):
    """Triples of a canonically ordered radial list: universal.py:115-233.

    For every centre (in list order) all position pairs a < b of its row (symmetric) or all
    ordered pairs a != b (not symmetric).  v2g_map = [angular term, centre, radial slot of the
    'reference pair', running count of that (centre, term, reference pair), 1]; the reference
    pair is (i, j) when symbol_j < symbol_k, else (i, k) (symmetric), always (i, j) otherwise.
    Returns v2g_map, ilist, jlist, klist, n1, n2, n3 (integer shifts)."""
    nij = len(i)
    z = np.zeros(0, dtype=np.int32)
    z3 = np.zeros((0, 3), dtype=S.dtype)
    empty = (np.zeros((0, 5), dtype=np.int32), z, z, z, z3, z3, z3)
    if nij == 0:
        return empty
    # rows of the list: the entries of one centre are contiguous (sorted by i)
    starts = np.flatnonzero(np.concatenate(([True], i[1:] != i[:-1])))
    counts = np.diff(np.concatenate((starts, [nij])))
    pa, pb = [], []
    for n in np.unique(counts):
        rows = starts[counts == n]
        if symmetric:
            a, b = np.triu_indices(n, 1)
        else:
            a, b = np.nonzero(~np.eye(n, dtype=bool))
        if len(a):
            pa.append((rows[:, None] + a[None, :]).ravel())
            pb.append((rows[:, None] + b[None, :]).ravel())
    if not pa:
        return empty
    pa = np.concatenate(pa)
    pb = np.concatenate(pb)
    order = np.lexsort((pb, pa))          # centres in list order, then (a, b) row-major
    pa, pb = pa[order], pb[order]
    tc, ta, tb = types[i[pa]], types[j[pa]], types[j[pb]]
    if symmetric:
        # element indices follow the sorted element list: index order == symbol order
        lo, hi = np.minimum(ta, tb), np.maximum(ta, tb)
        index = angular[tc, lo, hi]
        ref = np.where(ta < tb, pa, pb)
    else:
        index = angular[tc, ta, tb]
        ref = pa
    nijk = len(pa)
    centre = ilist[pa]
    # running count per (centre, term, reference pair); the reference pair is identified by
    # its position in the radial list (one position = one (i, j, S))
    n_index = int(angular.max()) + 1
    key = (centre.astype(np.int64) * n_index + index) * (nij + 1) + ref
    v2g = np.zeros((nijk, 5), dtype=np.int32)
    v2g[:, 0] = index
    v2g[:, 1] = centre
    v2g[:, 2] = inc[ref]
    v2g[:, 3] = _running_count(key)
    v2g[:, 4] = 1
    n1, n2 = S[pa], S[pb]
    return v2g, centre.astype(np.int32), jlist[pa].astype(np.int32), \
        jlist[pb].astype(np.int32), n1, n2, n2 - n1


def pair_distances(positions, cell, i, j, S):
    D = positions[j] - positions[i] + S.astype(np.float64) @ np.asarray(cell, dtype=np.float64)
    return np.sqrt((D * D).sum(axis=1)), D


def build_feed_arrays(i, j, S, positions, cell, types, l2g, elements,
                      kbody_terms_for_element, rcut, acut=None, angular=False,
                      symmetric=True):
    """All list-derived entries of the feed dict from ONE neighbour list built with the
    radius max(rcut, acut).  Returns a dict with the integer arrays (shifts as int; the caller
    casts n1 / n2 / n3 to the float dtype as the reference does, universal.py:76)."""
    i = np.asarray(i, dtype=np.int64)
    j = np.asarray(j, dtype=np.int64)
    S = np.asarray(S, dtype=np.int64).reshape(-1, 3)
    order = canonical_order(i, j, S)
    i, j, S = i[order], j[order], S[order]
    radial, ang = interaction_tables(elements, kbody_terms_for_element)
    d, D = pair_distances(np.asarray(positions, dtype=np.float64), cell, i, j, S)
    in_r = d < rcut
    ri, rj, rS = i[in_r], j[in_r], S[in_r]
    v2g, ilist, jlist = radial_metadata(ri, rj, types, l2g, radial)
    out = {"g2.v2g_map": v2g, "g2.ilist": ilist, "g2.jlist": jlist, "g2.n1": rS,
           "g2.d": d[in_r], "g2.D": D[in_r]}
    if angular:
        acut = rcut if acut is None else acut
        if np.round(acut - rcut, 2) == 0.0:
            # universal.py:828-830: the radial list (and its slots) serve the triples
            ai, aj, aS, a_il, a_jl, a_inc = ri, rj, rS, ilist, jlist, v2g[:, 2]
        else:
            in_a = d < acut
            ai, aj, aS = i[in_a], j[in_a], S[in_a]
            a_v2g, a_il, a_jl = radial_metadata(ai, aj, types, l2g, radial)
            a_inc = a_v2g[:, 2]
        g4, il, jl, kl, n1, n2, n3 = angular_metadata(ai, aj, aS, types, a_il, a_jl, a_inc,
                                                      ang, symmetric)
        out.update({"g4.v2g_map": g4, "g4.ilist": il, "g4.jlist": jl, "g4.klist": kl,
                    "g4.n1": n1, "g4.n2": n2, "g4.n3": n3})
    return out
