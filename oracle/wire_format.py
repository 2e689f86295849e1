"""
oracle/wire_format.py -- TEST INFRASTRUCTURE (see oracle/__init__.py).

Loop restatement of the reference's feature-dict metadata builders
    get_radial_metadata    tensoralloy/transformer/universal.py:46-112
    get_angular_metadata   tensoralloy/transformer/universal.py:115-233
    get_metadata           tensoralloy/transformer/universal.py:786-849
in PREDICT mode (iaxis = 0).  One pair / one triple at a time, dictionaries for the running
counters and for the (i, j, S) -> radial slot map, exactly as the reference does it (the
reference keys that map by Szudzik pairings of (i, j) and of the shift; a tuple serves the same
purpose).  The neighbour list is handed in (ASE's own order inside a row is not pinned by any
reference test: callers pass a canonically sorted list).
"""
from collections import Counter

import numpy as np


def _split_term(term):
    out = []
    for ch in term:
        if ch.isupper():
            out.append(ch)
        else:
            out[-1] += ch
    return out


def interactions(elements, kbody_terms_for_element):
    """universal.py:807-817."""
    radial, angular = {}, {}
    n = len(elements)
    for element in elements:
        for idx, term in enumerate(kbody_terms_for_element[element]):
            if len(_split_term(term)) == 2:
                radial[term] = idx
            else:
                angular[term] = idx - n
    return radial, angular


def radial_metadata(symbols, ilist0, jlist0, shifts, local_to_gsl, radial):
    """ilist0 / jlist0: local 0-based indices, rows sorted by i.  Returns
    (v2g_map, ilist, jlist, slot_of) with slot_of[(i_gsl, j_gsl, Sx, Sy, Sz)] = slot."""
    nij = len(ilist0)
    v2g = np.zeros((nij, 5), dtype=np.int32)
    ilist = np.zeros(nij, dtype=np.int32)
    jlist = np.zeros(nij, dtype=np.int32)
    counters = {}
    slot_of = {}
    for p in range(nij):
        a, b = int(ilist0[p]), int(jlist0[p])
        term = radial[f"{symbols[a]}{symbols[b]}"]
        gi, gj = local_to_gsl[a + 1], local_to_gsl[b + 1]
        ilist[p], jlist[p] = gi, gj
        cnt = counters.setdefault(gi, Counter())
        slot = cnt[term]
        cnt[term] += 1
        slot_of[(gi, gj) + tuple(int(x) for x in shifts[p])] = slot
        v2g[p] = (term, gi, slot, 0, 1 if gi > 0 else 0)
    return v2g, ilist, jlist, slot_of


def angular_metadata(symbols, gsl_to_local, ilist, jlist, shifts, slot_of, angular,
                     symmetric=True):
    """universal.py:136-232."""
    rows, order = {}, []
    for p, gi in enumerate(ilist):
        gi = int(gi)
        if gi not in rows:
            rows[gi] = []
            order.append(gi)
        rows[gi].append(p)
    out = []
    counters = {}
    for gi in order:
        plist = rows[gi]
        si = symbols[gsl_to_local[gi]]
        for a in range(len(plist)):
            pj = plist[a]
            gj = int(jlist[pj])
            sj = symbols[gsl_to_local[gj]]
            for b in (range(a + 1, len(plist)) if symmetric else range(len(plist))):
                pk = plist[b]
                gk = int(jlist[pk])
                sk = symbols[gsl_to_local[gk]]
                key_j = (gi, gj) + tuple(int(x) for x in shifts[pj])
                key_k = (gi, gk) + tuple(int(x) for x in shifts[pk])
                if symmetric:
                    term = f"{si}{''.join(sorted([sj, sk]))}"
                    ref = key_j if sj < sk else key_k
                else:
                    if key_j == key_k:
                        continue
                    term = f"{si}{sj}{sk}"
                    ref = key_j
                index = angular[term]
                c = counters.setdefault(gi, {}).setdefault(index, Counter())
                out.append((index, gi, slot_of[ref], c[ref], 1, gi, gj, gk,
                            tuple(shifts[pj]), tuple(shifts[pk])))
                c[ref] += 1
    n = len(out)
    v2g = np.zeros((n, 5), dtype=np.int32)
    il = np.zeros(n, dtype=np.int32)
    jl = np.zeros(n, dtype=np.int32)
    kl = np.zeros(n, dtype=np.int32)
    n1 = np.zeros((n, 3), dtype=np.int64)
    n2 = np.zeros((n, 3), dtype=np.int64)
    for q, row in enumerate(out):
        v2g[q] = row[0:5]
        il[q], jl[q], kl[q] = row[5:8]
        n1[q], n2[q] = row[8], row[9]
    return v2g, il, jl, kl, n1, n2, n2 - n1


def feed_metadata(symbols, positions, cell, nl, elements, kbody_terms_for_element,
                  local_to_gsl, gsl_to_local, rcut, acut=None, angular=False, symmetric=True):
    """universal.py:786-849 on the neighbour list `nl` = (i, j, S, d, D) built with the
    radius max(rcut, acut) and canonically sorted."""
    i, j, S, d = nl[0], nl[1], nl[2], nl[3]
    radial, ang = interactions(elements, kbody_terms_for_element)
    keep = d < rcut
    v2g, ilist, jlist, slot_of = radial_metadata(symbols, i[keep], j[keep], S[keep],
                                                 local_to_gsl, radial)
    out = {"g2.v2g_map": v2g, "g2.ilist": ilist, "g2.jlist": jlist, "g2.n1": S[keep]}
    if angular:
        acut = rcut if acut is None else acut
        if np.round(acut - rcut, 2) == 0.0:
            a_il, a_jl, a_S, a_slot = ilist, jlist, S[keep], slot_of
        else:
            ka = d < acut
            _, a_il, a_jl, a_slot = radial_metadata(symbols, i[ka], j[ka], S[ka],
                                                    local_to_gsl, radial)
            a_S = S[ka]
        g4 = angular_metadata(symbols, gsl_to_local, a_il, a_jl, a_S, a_slot, ang, symmetric)
        for key, val in zip(("g4.v2g_map", "g4.ilist", "g4.jlist", "g4.klist", "g4.n1",
                             "g4.n2", "g4.n3"), g4):
            out[key] = val
    return out
