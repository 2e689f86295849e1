"""
oracle/neighbor.py -- TEST INFRASTRUCTURE (see oracle/__init__.py).

Restatement of the neighbour search the reference delegates to ASE:
    ilist, jlist, S, d, D = ase.neighborlist.neighbor_list('ijSdD', atoms, rc)
called at transformer/universal.py:58 and neighbor.py:84 of the reference.

ASE (>= 3.21, requirements.txt:3) is a third-party dependency that is NOT under
/root/reference and not installed here.  Its published semantics, restated:
  * all DIRECTED pairs (i, j, S), S an integer cell-shift vector, with
        D = positions[j] - positions[i] + S @ cell
        d = sqrt(D.D)                       (float64, no eps)
        d <  rc                             (strict)
    the self pair (i == j, S == 0) excluded, images i == j with S != 0 kept;
  * S is relative to the positions AS GIVEN (atoms outside the cell are not
    wrapped in the result);
  * pairs are sorted by i; the order inside a row is an implementation detail
    and is not pinned by any reference test -> compare after canonical sort.
  * non-periodic directions get S == 0 only.

Two searches are provided; both decide membership with exactly the arithmetic
above on the original positions, so they agree bit for bit:
  * `neighbor_list_brute`  -- O(N^2 * images), for tiny cells (tests)
  * `neighbor_list`        -- KD-tree candidate search + exact filter (default)
"""
import numpy as np


def _complete_cell(cell, pbc):
    """Non-periodic directions may have zero lattice vectors; complete them to
    an invertible matrix the way ase.geometry.complete_cell does (only used to
    compute scaled coordinates / face distances)."""
    cell = np.array(cell, dtype=np.float64).reshape(3, 3)
    missing = [i for i in range(3) if not np.any(cell[i])]
    if not missing:
        return cell
    out = cell.copy()
    if len(missing) == 3:
        return np.eye(3)
    if len(missing) == 1:
        i = missing[0]
        a, b = out[(i + 1) % 3], out[(i + 2) % 3]
        v = np.cross(a, b)
        out[i] = v / np.linalg.norm(v)
        return out
    # two missing
    k = [i for i in range(3) if i not in missing][0]
    a = out[k]
    e = np.eye(3)[np.argmin(np.abs(a))]
    v1 = np.cross(a, e)
    v1 /= np.linalg.norm(v1)
    v2 = np.cross(a, v1)
    v2 /= np.linalg.norm(v2)
    out[missing[0]] = v1
    out[missing[1]] = v2
    return out


def face_distances(cell):
    """Perpendicular distance between opposite faces of the cell for each
    lattice direction (ASE: 1 / |b_c| with b_c the reciprocal vectors)."""
    rec = np.linalg.pinv(cell).T
    n = np.linalg.norm(rec, axis=1)
    return np.where(n > 0, 1.0 / np.where(n > 0, n, 1.0), 1.0)


def _exact_filter(positions, cell, i, j, S, rc):
    """The ASE membership test, on the positions as given."""
    D = positions[j] - positions[i] + S.astype(np.float64).dot(cell)
    d = np.sqrt(np.sum(D * D, axis=1))
    keep = d < rc
    keep &= ~((i == j) & np.all(S == 0, axis=1))
    return i[keep], j[keep], S[keep], d[keep], D[keep]


def _sort_by_i(i, j, S, d, D):
    order = np.lexsort((S[:, 2], S[:, 1], S[:, 0], j, i))
    return i[order], j[order], S[order], d[order], D[order]


def _image_ranges(cell_c, pbc, rc):
    fd = face_distances(cell_c)
    return [int(np.ceil(rc / fd[c])) if pbc[c] else 0 for c in range(3)]


def neighbor_list_brute(positions, cell, pbc, rc):
    """O(N^2) search over all images.  Returns (i, j, S, d, D) canonically
    sorted by (i, j, Sx, Sy, Sz)."""
    positions = np.asarray(positions, dtype=np.float64)
    cell = np.asarray(cell, dtype=np.float64).reshape(3, 3)
    pbc = np.asarray(pbc, dtype=bool).reshape(3)
    n = len(positions)
    cc = _complete_cell(cell, pbc)
    # wrap to find candidate images, then express S w.r.t. the given positions
    scaled = np.linalg.solve(cc.T, positions.T).T
    s0 = np.where(pbc, np.floor(scaled), 0).astype(np.int64)
    rng = _image_ranges(cc, pbc, rc)
    out_i, out_j, out_S = [], [], []
    ii, jj = np.meshgrid(np.arange(n), np.arange(n), indexing='ij')
    ii = ii.ravel()
    jj = jj.ravel()
    for sx in range(-rng[0] - 1, rng[0] + 2):
        if not pbc[0] and sx != 0:
            continue
        for sy in range(-rng[1] - 1, rng[1] + 2):
            if not pbc[1] and sy != 0:
                continue
            for sz in range(-rng[2] - 1, rng[2] + 2):
                if not pbc[2] and sz != 0:
                    continue
                Sw = np.array([sx, sy, sz], dtype=np.int64)
                S = Sw[None, :] - s0[jj] + s0[ii]
                out_i.append(ii)
                out_j.append(jj)
                out_S.append(S)
    i = np.concatenate(out_i)
    j = np.concatenate(out_j)
    S = np.concatenate(out_S)
    # the same (i, j, S) can be generated only once because Sw <-> S is 1:1
    return _sort_by_i(*_exact_filter(positions, cell, i, j, S, rc))


def neighbor_list(positions, cell, pbc, rc, workers=-1):
    """KD-tree candidate search followed by the exact ASE membership test.
    Returns (i, j, S, d, D) canonically sorted by (i, j, Sx, Sy, Sz)."""
    from scipy.spatial import cKDTree
    positions = np.asarray(positions, dtype=np.float64)
    cell = np.asarray(cell, dtype=np.float64).reshape(3, 3)
    pbc = np.asarray(pbc, dtype=bool).reshape(3)
    n = len(positions)
    cc = _complete_cell(cell, pbc)
    scaled = np.linalg.solve(cc.T, positions.T).T
    s0 = np.where(pbc, np.floor(scaled), 0).astype(np.int64)
    wrapped = positions - s0.astype(np.float64).dot(cell)
    scaled_w = scaled - s0
    fd = face_distances(cc)
    margin = (rc * (1.0 + 1e-9) + 1e-6) / fd  # in scaled units
    rng = _image_ranges(cc, pbc, rc)
    ext_pos, ext_owner, ext_S = [], [], []
    for sx in range(-rng[0], rng[0] + 1):
        for sy in range(-rng[1], rng[1] + 1):
            for sz in range(-rng[2], rng[2] + 1):
                Sw = np.array([sx, sy, sz], dtype=np.int64)
                sc = scaled_w + Sw
                ok = np.ones(n, dtype=bool)
                for c in range(3):
                    if pbc[c]:
                        ok &= (sc[:, c] > -margin[c]) & (sc[:, c] < 1 + margin[c])
                idx = np.nonzero(ok)[0]
                if len(idx) == 0:
                    continue
                ext_pos.append(wrapped[idx] + Sw.astype(np.float64).dot(cell))
                ext_owner.append(idx)
                ext_S.append(np.broadcast_to(Sw, (len(idx), 3)))
    ext_pos = np.concatenate(ext_pos)
    ext_owner = np.concatenate(ext_owner)
    ext_S = np.concatenate(ext_S)
    tree = cKDTree(ext_pos)
    hits = tree.query_ball_point(wrapped, rc * (1.0 + 1e-9) + 1e-6,
                                 workers=workers)
    counts = np.fromiter((len(h) for h in hits), dtype=np.int64, count=n)
    i = np.repeat(np.arange(n, dtype=np.int64), counts)
    e = np.fromiter((x for h in hits for x in h), dtype=np.int64,
                    count=int(counts.sum()))
    j = ext_owner[e]
    S = ext_S[e] - s0[j] + s0[i]
    return _sort_by_i(*_exact_filter(positions, cell, i, j, S, rc))


def canonical_sort(i, j, S):
    """Canonical order used by every neighbour-list comparison."""
    S = np.asarray(S).reshape(-1, 3)
    order = np.lexsort((S[:, 2], S[:, 1], S[:, 0], j, i))
    return np.asarray(i)[order], np.asarray(j)[order], S[order]
