"""Distinct 128-byte lines touched by one warp-level neighbour gather, for the thread-per-atom
row layout (L = 1) and the lane-split layouts (L lanes per atom), on a rattled fcc Ni lattice in
the library's cell-sorted order (dev tool, CPU only).  The L1 data stage of B200 serves a gather
in one wavefront per distinct line: the L = 1 figure (21.5) reproduces the 22.5 wavefronts per
gather that ncu measured on `k_eam_rho` (profiles/r01k).  usage: python tools/sim_gather_lines.py"""
import numpy as np
from scipy.spatial import cKDTree

rng = np.random.default_rng(611)
a, nc, rc = 3.52, 24, 6.5
base = np.array([[0, 0, 0], [.5, .5, 0], [.5, 0, .5], [0, .5, .5]])
g = np.stack(np.meshgrid(*[np.arange(nc)] * 3, indexing='ij'), -1).reshape(-1, 3)
pos = ((g[:, None, :] + base[None]) * a).reshape(-1, 3)
pos += rng.normal(scale=0.05, size=pos.shape)
box = a * nc
pos %= box
n = len(pos)
nb = int(box // (rc * (1 + 1e-6)))
c = np.floor(pos / (box / nb)).astype(int) % nb
B = 2
tl = (nb + B - 1) // B
rank = ((c[:, 2] // B * tl + c[:, 1] // B) * tl + c[:, 0] // B) * 8 + \
    ((c[:, 2] % B) * B + c[:, 1] % B) * B + c[:, 0] % B
order = np.lexsort((np.arange(n), rank))
pos, c = pos[order], c[order]
nbrs = cKDTree(pos, boxsize=box).query_ball_point(pos, rc)
rows = []
for i in range(n):
    js = np.array([j for j in nbrs[i] if j != i])
    d = (c[js] - c[i] + nb // 2) % nb - nb // 2
    rows.append(js[np.lexsort((js, d[:, 0], d[:, 1], d[:, 2]))])   # the library's cell walk


def lines(L, per_line, warps=300, first=5000):
    apw = 32 // L
    tot = steps = 0
    for w in range(warps):
        rr = [rows[i] for i in range(first + w * apw, first + (w + 1) * apw)]
        for k in range((max(len(x) for x in rr) + L - 1) // L):
            idx = np.concatenate([x[k * L:(k + 1) * L] for x in rr])
            tot += len(np.unique(idx // per_line))
            steps += 1
    return tot / steps


print(f"{n} atoms, {np.mean([len(r) for r in rows]):.1f} neighbours per atom")
for per_line, name in ((4, 'Atom4 32 B'), (8, 'Rec16 16 B')):
    print(name, ' '.join(f"L={L}: {lines(L, per_line):.1f}" for L in (1, 2, 4, 8, 16, 32)))
