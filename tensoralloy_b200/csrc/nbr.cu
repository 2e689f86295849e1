// nbr.cu -- GPU cell list, periodic ghost images and ELL neighbour lists.
//
// Replaces ase.neighborlist.neighbor_list as called by the reference at
// tensoralloy/transformer/universal.py:58 and tensoralloy/neighbor.py:84, and the
// Python index-map loops of universal.py:69-106.
//
// Pipeline (all on the caller's stream):
//   k_bin          scaled coords -> owned-cell rank (+ wrap shift), histogram
//   scan           cell_count -> cell_start
//   k_scatter      atoms -> cell-sorted permutation (then k_rank_in_cell makes the
//                  order inside a cell = caller index order: deterministic)
//   k_gather_owned wrapped positions in sorted order -> Atom4 records
//   k_ext_cells    extended (ghost-padded) cell table; ghost counts
//   scan           ghost counts -> ghost starts
//   k_fill_ghosts  ghost records = source record + S.h
//   k_count / k_nbr_warp<false>  neighbours per atom (thread- or warp-per-atom with
//                  __ballot_sync compaction), k_slice_stats: slice widths, nij, nnl_max
//   scan           slice widths -> slice_ptr
//   k_fill         ELL entries, 32 atoms per slice, entry k of lane l at
//                  col[(slice_ptr[s] + k) * 32 + l]  (coalesced for thread-per-atom
//                  consumers; no per-pair shift vectors: ghosts carry them)
//
// Domain decomposition: the caller may append HALO atoms (received from other
// ranks, already shifted into this rank's frame) after its owned atoms.  Owned and
// halo atoms form two cell-sorted groups ([0,n) and [n,n_loc) of the extended
// array); neighbour rows exist for owned atoms only; periodic images are generated
// for both groups.
//
// Membership is ASE's: D = pos[j] - pos[i] + S.cell, sqrt(D.D) < rc in float64.
// The fast test uses the pre-shifted ghost records; candidates whose d^2 is
// within 1e-9 relative of rc^2 are re-decided with exactly ASE's expression on
// the caller's positions, so the list is bit-exact.
#include <math.h>
#include <stdlib.h>

#include "tab_internal.h"


// ---------------------------------------------------------------------------
// cell indexing
// ---------------------------------------------------------------------------
__host__ __device__ inline int owned_rank(const Grid &g, int cx, int cy, int cz) {
    const int B = g.tb;
    const int tx = cx / B, ty = cy / B, tz = cz / B;
    const int lx = cx % B, ly = cy % B, lz = cz % B;
    return ((tz * g.tl[1] + ty) * g.tl[0] + tx) * (B * B * B) + (lz * B + ly) * B + lx;
}

__host__ __device__ inline void owned_unrank(const Grid &g, int rank, int &cx,
                                             int &cy, int &cz) {
    const int B = g.tb;
    const int local = rank % (B * B * B);
    int t = rank / (B * B * B);
    const int tx = t % g.tl[0];
    t /= g.tl[0];
    const int ty = t % g.tl[1];
    const int tz = t / g.tl[1];
    cx = tx * B + local % B;
    cy = ty * B + (local / B) % B;
    cz = tz * B + local / (B * B);
}

__host__ __device__ inline int floordiv(int a, int b) {
    int q = a / b;
    if ((a % b != 0) && ((a < 0) != (b < 0))) --q;
    return q;
}

// ---------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------
__global__ void k_bin(int n, int n_owned, const double *__restrict__ pos,
                      const int *__restrict__ types, Grid g,
                      int *__restrict__ cell_of, int *__restrict__ s0,
                      uint32_t *__restrict__ cell_count,
                      unsigned long long *__restrict__ stats) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (types) {
        const int t = types[i];
        if (t > 0) atomicMax(&stats[3], (unsigned long long)t);
    }
    const double x = pos[3 * i] - g.origin[0], y = pos[3 * i + 1] - g.origin[1],
                 z = pos[3 * i + 2] - g.origin[2];
    int c[3], sh[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const double s = x * g.hinv[k] + y * g.hinv[3 + k] + z * g.hinv[6 + k];
        int b = (int)floor(s * (double)g.nb[k]);
        if (g.pbc[k]) {
            sh[k] = floordiv(b, g.nb[k]);
            b -= sh[k] * g.nb[k];
        } else {
            sh[k] = 0;
            // atoms more than one bin outside a non-periodic frame: stats[4] (the
            // tile kernel's distance form wants bounded coordinates)
            if (b < -1 || b > g.nb[k]) atomicMax(&stats[4], 1ull);
            b = min(max(b, 0), g.nb[k] - 1);
        }
        c[k] = b;
    }
    // slot = group * n_slots + rank: owned atoms sort before halo atoms
    const int slot = owned_rank(g, c[0], c[1], c[2]) + (i >= n_owned ? g.n_slots : 0);
    cell_of[i] = slot;
    s0[i] = tab_pack_shift(sh[0], sh[1], sh[2]);
    atomicAdd(&cell_count[slot], 1u);
}

__global__ void k_scatter(int n, const int *__restrict__ cell_of,
                          const uint32_t *__restrict__ cell_start,
                          uint32_t *__restrict__ cell_fill,
                          int *__restrict__ perm) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int rank = cell_of[i];
    const uint32_t slot = atomicAdd(&cell_fill[rank], 1u);
    perm[cell_start[rank] + slot] = i;
}

// order inside a cell = ascending caller index (removes the atomics' race order):
// every atom counts the members of its cell with a smaller index -> its rank.
__global__ void k_rank_in_cell(int n, const int *__restrict__ cell_of,
                               const uint32_t *__restrict__ cell_start,
                               const uint32_t *__restrict__ cell_count,
                               const int *__restrict__ perm_in,
                               int *__restrict__ perm_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int c = cell_of[i];
    const uint32_t start = cell_start[c], cnt = cell_count[c];
    uint32_t rank = 0;
    for (uint32_t k = 0; k < cnt; ++k) rank += perm_in[start + k] < i ? 1u : 0u;
    perm_out[start + rank] = i;
}

// fixed-point frame of the float32 records (tab_internal.h Rec16); inv_delta == 0: unused
struct QFrame {
    double ox, oy, oz, inv_delta;
};

__device__ __forceinline__ Rec16 make_rec16(const QFrame &q, double x, double y, double z) {
    Rec16 r;
    r.qx = __double2int_rn((x - q.ox) * q.inv_delta);
    r.qy = __double2int_rn((y - q.oy) * q.inv_delta);
    r.qz = __double2int_rn((z - q.oz) * q.inv_delta);
    r.w = 0.f;
    return r;
}

// REF = 1: build (store the reference positions of the displacement check);
// REF = 2: refresh (largest squared displacement since the build -> disp[0], float bits)
template <int REF>
__global__ void __launch_bounds__(256)
k_gather_owned(int n, const double *__restrict__ pos, const int *__restrict__ types, Grid g,
               const int *__restrict__ perm, const int *__restrict__ s0,
               Atom4 *__restrict__ atoms, uint8_t *__restrict__ types_ext,
               double *__restrict__ pos_ref, unsigned int *__restrict__ disp, QFrame qf,
               Rec16 *__restrict__ rec16) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    float d2 = 0.f;
    if (idx < n) {
        const int i = perm[idx];
        double x = pos[3 * i], y = pos[3 * i + 1], z = pos[3 * i + 2];
        const int s = s0[i];
        if (s != tab_pack_shift(0, 0, 0)) {
            int a, b, c;
            tab_unpack_shift(s, a, b, c);
            x -= a * g.h[0] + b * g.h[3] + c * g.h[6];
            y -= a * g.h[1] + b * g.h[4] + c * g.h[7];
            z -= a * g.h[2] + b * g.h[5] + c * g.h[8];
        }
        Atom4 r;
        r.x = x;
        r.y = y;
        r.z = z;
        r.w = 0.0;
        atoms[idx] = r;
        if (rec16) rec16[idx] = make_rec16(qf, x, y, z);
        if (types_ext) types_ext[idx] = types ? (uint8_t)types[i] : (uint8_t)0;
        if (REF == 1 && pos_ref) {
            pos_ref[3 * (size_t)idx] = x;
            pos_ref[3 * (size_t)idx + 1] = y;
            pos_ref[3 * (size_t)idx + 2] = z;
        }
        if (REF == 2 && pos_ref) {
            const double ex = x - pos_ref[3 * (size_t)idx], ey = y - pos_ref[3 * (size_t)idx + 1],
                         ez = z - pos_ref[3 * (size_t)idx + 2];
            // rounded UP: the check must never under-report
            d2 = __double2float_ru(ex * ex + ey * ey + ez * ez);
        }
    }
    if (REF == 2 && disp) {
        __shared__ float s_mx[8];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) d2 = fmaxf(d2, __shfl_xor_sync(0xffffffffu, d2, d));
        if ((threadIdx.x & 31) == 0) s_mx[threadIdx.x >> 5] = d2;
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int w = 1; w < 8; ++w) d2 = fmaxf(d2, s_mx[w]);
            // non-negative floats order like their bit patterns; NaN (0x7fc00000) sorts on top
            if (d2 != 0.f) atomicMax(disp, __float_as_uint(d2));
        }
    }
}

// one thread per extended cell.  Table entry = {start_a, count_a, start_b,
// count_b}: interior cells -> owned range + halo range; ghost cells get the total
// count of their source cell (their start is filled in by k_fill_ghosts).
__global__ void k_ext_cells(Grid g, const uint32_t *__restrict__ cell_start,
                            const uint32_t *__restrict__ cell_count,
                            uint4 *__restrict__ ext_tab,
                            uint32_t *__restrict__ gcount) {
    const int lin = blockIdx.x * blockDim.x + threadIdx.x;
    if (lin >= g.n_ecells) return;
    int e[3];
    e[0] = lin % g.ne[0];
    e[1] = (lin / g.ne[0]) % g.ne[1];
    e[2] = lin / (g.ne[0] * g.ne[1]);
    int c[3];
    bool ghost = false;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        c[k] = e[k] - g.g[k];
        if (c[k] < 0 || c[k] >= g.nb[k]) {
            ghost = true;
            c[k] -= floordiv(c[k], g.nb[k]) * g.nb[k];
        }
    }
    const int rank = owned_rank(g, c[0], c[1], c[2]);
    const uint32_t ca = cell_count[rank], cb = cell_count[g.n_slots + rank];
    if (ghost) {
        gcount[lin] = ca + cb;
        ext_tab[lin] = make_uint4(0u, ca + cb, 0u, 0u);
    } else {
        gcount[lin] = 0;
        ext_tab[lin] = make_uint4(cell_start[rank], ca, cell_start[g.n_slots + rank], cb);
    }
}

// one warp per extended cell; ghosts only.  Ghost records of a cell = images of
// the owned atoms of the source cell followed by images of its halo atoms.
__global__ void k_fill_ghosts(Grid g, int n_loc,
                              const uint32_t *__restrict__ cell_start,
                              const uint32_t *__restrict__ cell_count,
                              const uint32_t *__restrict__ gstart,
                              const uint32_t *__restrict__ gcount,
                              uint4 *__restrict__ ext_tab,
                              Atom4 *__restrict__ atoms,
                              uint8_t *__restrict__ types_ext,
                              int *__restrict__ ghost_owner,
                              int *__restrict__ ghost_S, int refresh_only, QFrame qf,
                              Rec16 *__restrict__ rec16) {
    // 32 lanes per cell of full width, 8 per half-width cell (~4 atoms)
    const int lpc = g.tb == 4 ? 8 : 32;
    const int lin = (blockIdx.x * blockDim.x + threadIdx.x) / lpc;
    const int lane = threadIdx.x & (lpc - 1);
    if (lin >= g.n_ecells) return;
    const uint32_t cnt = gcount[lin];
    if (cnt == 0) return;
    int e[3], c[3], S[3];
    e[0] = lin % g.ne[0];
    e[1] = (lin / g.ne[0]) % g.ne[1];
    e[2] = lin / (g.ne[0] * g.ne[1]);
    bool ghost = false;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        c[k] = e[k] - g.g[k];
        S[k] = 0;
        if (c[k] < 0 || c[k] >= g.nb[k]) {
            ghost = true;
            S[k] = floordiv(c[k], g.nb[k]);
            c[k] -= S[k] * g.nb[k];
        }
    }
    if (!ghost) return;
    const int rank = owned_rank(g, c[0], c[1], c[2]);
    const uint32_t src_a = cell_start[rank], cnt_a = cell_count[rank];
    const uint32_t src_b = cell_start[g.n_slots + rank];
    const uint32_t dst = (uint32_t)n_loc + gstart[lin];
    if (lane == 0 && !refresh_only) ext_tab[lin] = make_uint4(dst, cnt, 0u, 0u);
    const double sx = S[0] * g.h[0] + S[1] * g.h[3] + S[2] * g.h[6];
    const double sy = S[0] * g.h[1] + S[1] * g.h[4] + S[2] * g.h[7];
    const double sz = S[0] * g.h[2] + S[1] * g.h[5] + S[2] * g.h[8];
    const int packed = tab_pack_shift(S[0], S[1], S[2]);
    for (uint32_t k = lane; k < cnt; k += lpc) {
        const uint32_t src = k < cnt_a ? src_a + k : src_b + (k - cnt_a);
        Atom4 a = atoms[src];
        a.x += sx;
        a.y += sy;
        a.z += sz;
        atoms[dst + k] = a;
        if (rec16) rec16[dst + k] = make_rec16(qf, a.x, a.y, a.z);
        if (!refresh_only) {
            types_ext[dst + k] = types_ext[src];
            ghost_owner[dst + k - n_loc] = (int)src;
            ghost_S[dst + k - n_loc] = packed;
        }
    }
}

// per-step refresh of the ghost records: one thread per ghost atom (source index and image
// shift were stored by k_fill_ghosts at build time)
__global__ void k_refresh_ghosts(int n_ghost, int n_loc, Grid g,
                                 const int *__restrict__ ghost_owner,
                                 const int *__restrict__ ghost_S, Atom4 *__restrict__ atoms,
                                 QFrame qf, Rec16 *__restrict__ rec16) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_ghost) return;
    int S[3];
    tab_unpack_shift(ghost_S[k], S[0], S[1], S[2]);
    Atom4 a = atoms[ghost_owner[k]];
    a.x += S[0] * g.h[0] + S[1] * g.h[3] + S[2] * g.h[6];
    a.y += S[0] * g.h[1] + S[1] * g.h[4] + S[2] * g.h[7];
    a.z += S[0] * g.h[2] + S[1] * g.h[5] + S[2] * g.h[8];
    atoms[n_loc + k] = a;
    if (rec16) rec16[n_loc + k] = make_rec16(qf, a.x, a.y, a.z);
}

struct ExactCtx {
    const double *pos;       // caller positions
    const int *perm;
    const int *s0;           // caller order
    const int *ghost_owner;
    const int *ghost_S;
    int n_loc;               // owned + halo: first library-generated image
};

// ASE's membership expression on the caller's positions (no FMA contraction).
__device__ __noinline__ bool exact_inside(const Grid &g, const ExactCtx &x, int i,
                                          int j) {
    const int oi = x.perm[i];
    int owner = j, Sa = 0, Sb = 0, Sc = 0;
    if (j >= x.n_loc) {
        owner = x.ghost_owner[j - x.n_loc];
        tab_unpack_shift(x.ghost_S[j - x.n_loc], Sa, Sb, Sc);
    }
    const int oj = x.perm[owner];
    int ia, ib, ic, ja, jb, jc;
    tab_unpack_shift(x.s0[oi], ia, ib, ic);
    tab_unpack_shift(x.s0[oj], ja, jb, jc);
    const double S0 = (double)(Sa - ja + ia), S1 = (double)(Sb - jb + ib),
                 S2 = (double)(Sc - jc + ic);
    double D[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const double sh = __dadd_rn(__dadd_rn(__dmul_rn(S0, g.h[k]),
                                              __dmul_rn(S1, g.h[3 + k])),
                                    __dmul_rn(S2, g.h[6 + k]));
        D[k] = __dadd_rn(__dsub_rn(x.pos[3 * oj + k], x.pos[3 * oi + k]), sh);
    }
    const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(D[0], D[0]), __dmul_rn(D[1], D[1])),
                                __dmul_rn(D[2], D[2]));
    return __dsqrt_rn(d2) < g.rc;
}

// Shared traversal of the candidate cells of one owned atom.  F(j) is called
// for every neighbour in deterministic order.
template <typename F>
__device__ __forceinline__ void for_each_neighbor(
    const Grid &g, const ExactCtx &x, int idx, int rank,
    const Atom4 *__restrict__ atoms, const uint4 *__restrict__ ext_tab, F &&f) {
    int cx, cy, cz;
    owned_unrank(g, rank, cx, cy, cz);
    const Atom4 me = atoms[idx];
    const double tol = 1e-9 * g.rc2;
    for (int dz = -g.sr[2]; dz <= g.sr[2]; ++dz) {
        const int ez = cz + dz;
        if (!g.pbc[2] && (ez < 0 || ez >= g.nb[2])) continue;
        for (int dy = -g.sr[1]; dy <= g.sr[1]; ++dy) {
            const int ey = cy + dy;
            if (!g.pbc[1] && (ey < 0 || ey >= g.nb[1])) continue;
            for (int dx = -g.sr[0]; dx <= g.sr[0]; ++dx) {
                const int ex = cx + dx;
                if (!g.pbc[0] && (ex < 0 || ex >= g.nb[0])) continue;
                const int lin = ((ez + g.g[2]) * g.ne[1] + (ey + g.g[1])) * g.ne[0] +
                                (ex + g.g[0]);
                const uint4 t = ext_tab[lin];
#pragma unroll
                for (int part = 0; part < 2; ++part) {
                    const uint32_t start = part ? t.z : t.x;
                    const uint32_t cnt = part ? t.w : t.y;
                    for (uint32_t k = 0; k < cnt; ++k) {
                        const int j = (int)(start + k);
                        if (j == idx) continue;
                        const Atom4 a = atoms[j];
                        const double ddx = a.x - me.x, ddy = a.y - me.y,
                                     ddz = a.z - me.z;
                        const double d2 = ddx * ddx + ddy * ddy + ddz * ddz;
                        bool in = d2 < g.rc2;
                        if (fabs(d2 - g.rc2) <= tol) in = exact_inside(g, x, idx, j);
                        if (in) f(j);
                    }
                }
            }
        }
    }
}

// ---- thread-per-atom variants (large systems: lanes = the 32 atoms of a slice,
//      candidate loads are warp-wide broadcasts)
__global__ void __launch_bounds__(128)
k_count(int n, Grid g, ExactCtx x, const int *__restrict__ cell_of,
        const Atom4 *__restrict__ atoms, const uint4 *__restrict__ ext_tab,
        const uint8_t *__restrict__ types_ext, int n_types,
        int *__restrict__ counts, int *__restrict__ tcounts) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    int cnt = 0;
    const int rank = cell_of[x.perm[idx]];
    if (n_types > 1) {
        int tc[TAB_MAX_ELEMENTS];
        for (int t = 0; t < n_types; ++t) tc[t] = 0;
        for_each_neighbor(g, x, idx, rank, atoms, ext_tab, [&](int j) {
            ++cnt;
            ++tc[types_ext[j]];
        });
        for (int t = 0; t < n_types; ++t) tcounts[(size_t)idx * n_types + t] = tc[t];
    } else {
        for_each_neighbor(g, x, idx, rank, atoms, ext_tab, [&](int) { ++cnt; });
        tcounts[idx] = cnt;
    }
    counts[idx] = cnt;
}

__global__ void __launch_bounds__(128)
k_fill(int n, Grid g, ExactCtx x, const int *__restrict__ cell_of,
       const Atom4 *__restrict__ atoms, const uint8_t *__restrict__ types_ext,
       const uint4 *__restrict__ ext_tab, int n_types,
       const int *__restrict__ tcounts,
       const uint32_t *__restrict__ slice_w,
       const uint32_t *__restrict__ slice_ptr, uint32_t *__restrict__ col, uint32_t pad) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int s = idx >> 5, lane = idx & 31;
    if (s >= (n + 31) / 32) return;
    uint32_t *base = col + ((size_t)slice_ptr[s] * 32u + lane);
    uint32_t k = 0;
    if (idx < n) {
        const int rank = cell_of[x.perm[idx]];
        if (n_types > 1) {
            // rows are grouped by neighbour species (stable inside a species)
            uint32_t off[TAB_MAX_ELEMENTS];
            uint32_t run = 0;
            for (int t = 0; t < n_types; ++t) {
                off[t] = run;
                run += (uint32_t)tcounts[(size_t)idx * n_types + t];
            }
            for_each_neighbor(g, x, idx, rank, atoms, ext_tab, [&](int j) {
                const uint32_t t = types_ext[j];
                base[(size_t)(off[t]++) * 32u] = (uint32_t)j | (t << TAB_COL_TYPE_SHIFT);
            });
            k = run;
        } else {
            for_each_neighbor(g, x, idx, rank, atoms, ext_tab, [&](int j) {
                base[(size_t)k * 32u] =
                    (uint32_t)j | ((uint32_t)types_ext[j] << TAB_COL_TYPE_SHIFT);
                ++k;
            });
        }
    }
    const uint32_t w = slice_w[s];
    for (; k < w; ++k) base[(size_t)k * 32u] = pad;
}

// ---- warp-per-atom variants (small systems: the 32 lanes test 32 candidates at a
//      time, __ballot_sync compaction keeps the deterministic candidate order).
//      FILL = false: count only.
template <bool FILL>
__global__ void __launch_bounds__(128)
k_nbr_warp(int n, Grid g, ExactCtx x, const int *__restrict__ cell_of,
           const Atom4 *__restrict__ atoms, const uint8_t *__restrict__ types_ext,
           const uint4 *__restrict__ ext_tab, int n_types, int *__restrict__ counts,
           int *__restrict__ tcounts, const uint32_t *__restrict__ slice_w,
           const uint32_t *__restrict__ slice_ptr, uint32_t *__restrict__ col, uint32_t pad) {
    const int idx = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (idx >= n) return;
    const uint32_t lt = (1u << lane) - 1u;
    int cx, cy, cz;
    owned_unrank(g, cell_of[x.perm[idx]], cx, cy, cz);
    const Atom4 me = atoms[idx];
    const double tol = 1e-9 * g.rc2;
    uint32_t off[TAB_MAX_ELEMENTS];
    uint32_t total = 0;
    if (FILL) {
        uint32_t run = 0;
        for (int t = 0; t < n_types; ++t) {
            off[t] = run;
            run += (uint32_t)tcounts[(size_t)idx * n_types + t];
        }
    } else {
        for (int t = 0; t < n_types; ++t) off[t] = 0;
    }
    uint32_t *base = FILL ? col + ((size_t)slice_ptr[idx >> 5] * 32u + (idx & 31)) : nullptr;
    for (int dz = -g.sr[2]; dz <= g.sr[2]; ++dz) {
        const int ez = cz + dz;
        if (!g.pbc[2] && (ez < 0 || ez >= g.nb[2])) continue;
        for (int dy = -g.sr[1]; dy <= g.sr[1]; ++dy) {
            const int ey = cy + dy;
            if (!g.pbc[1] && (ey < 0 || ey >= g.nb[1])) continue;
            for (int dx = -g.sr[0]; dx <= g.sr[0]; ++dx) {
                const int ex = cx + dx;
                if (!g.pbc[0] && (ex < 0 || ex >= g.nb[0])) continue;
                const int lin = ((ez + g.g[2]) * g.ne[1] + (ey + g.g[1])) * g.ne[0] +
                                (ex + g.g[0]);
                const uint4 t4 = ext_tab[lin];
#pragma unroll
                for (int part = 0; part < 2; ++part) {
                    const uint32_t start = part ? t4.z : t4.x;
                    const uint32_t cnt = part ? t4.w : t4.y;
                    for (uint32_t k0 = 0; k0 < cnt; k0 += 32) {
                        const uint32_t k = k0 + lane;
                        bool in = false;
                        int j = 0;
                        if (k < cnt) {
                            j = (int)(start + k);
                            if (j != idx) {
                                const Atom4 a = atoms[j];
                                const double ddx = a.x - me.x, ddy = a.y - me.y,
                                             ddz = a.z - me.z;
                                const double d2 = ddx * ddx + ddy * ddy + ddz * ddz;
                                in = d2 < g.rc2;
                                if (fabs(d2 - g.rc2) <= tol) in = exact_inside(g, x, idx, j);
                            }
                        }
                        const uint32_t t = in ? types_ext[j] : 0u;
                        if (n_types == 1) {
                            const uint32_t m = __ballot_sync(0xffffffffu, in);
                            if (FILL && in)
                                base[(size_t)(off[0] + __popc(m & lt)) * 32u] = (uint32_t)j;
                            off[0] += __popc(m);
                        } else {
                            for (int s = 0; s < n_types; ++s) {
                                const uint32_t m = __ballot_sync(0xffffffffu, in && t == (uint32_t)s);
                                if (FILL && in && t == (uint32_t)s)
                                    base[(size_t)(off[s] + __popc(m & lt)) * 32u] =
                                        (uint32_t)j | (t << TAB_COL_TYPE_SHIFT);
                                off[s] += __popc(m);
                            }
                        }
                    }
                }
            }
        }
    }
    if (!FILL) {
        for (int t = 0; t < n_types; ++t) total += off[t];
        if (lane == 0) {
            counts[idx] = (int)total;
            for (int t = 0; t < n_types; ++t) tcounts[(size_t)idx * n_types + t] = (int)off[t];
        }
    } else {
        // pad the rest of this atom's column up to the slice width
        const uint32_t w = slice_w[idx >> 5];
        const uint32_t have = (uint32_t)counts[idx];
        for (uint32_t k = have + lane; k < w; k += 32) base[(size_t)k * 32u] = pad;
    }
}

// ---- block-per-tile variant (large systems), SINGLE PASS.
//      One block = one B x B x B tile of owned cells, one thread = one atom of the
//      tile.  The candidates of the tile's neighbourhood box (tile cells +- sr) are
//      staged ONCE per block in shared memory, cell by cell in ascending
//      (ez, ey, ex) order; every lane then walks the cells of its own cell's
//      neighbourhood and reads the candidates as shared-memory broadcasts.  Per-atom
//      candidate order is identical to for_each_neighbor(), so the rows hold the
//      same entries in the same order as the thread-per-atom kernels (which remain
//      for odd geometries).
//
//      Hits are written straight into rows of a FIXED capacity `wcap` (row of atom
//      idx, entry k at rows[((idx >> 5) * wcap + k) * 32 + (idx & 31)]) while they
//      are counted; counting continues past the capacity, so the host can see an
//      overflow (nnl_max > wcap) and repeat the launch with the exact width.  For a
//      single species these rows ARE the list (slice_ptr[s] = s * wcap); with
//      several species k_rows_by_species regroups them into compact slices.
//
//      Distance test = float32 PRE-FILTER + exact float64 decision.  A candidate is
//      staged as one 16-byte record (x, y, z relative to the box centre, rounded to
//      float32, + the entry bits), so a test is one LDS.128, six FP32 operations and
//      two FSETP.  The float32 d^2 differs from the true d^2 by less than `band`
//      (host-computed from the box size); candidates with |d^2 - rc^2| < band are
//      undecidable in float32: the cell segment they belong to is re-scanned by
//      nbt_scan_exact, which repeats the float64 test of for_each_neighbor()
//      (including ASE's exact expression within 1e-9 rc^2).  The band is ~1e-4 A^2
//      wide, so re-scans touch ~0.1 % of the segments of a random structure.
#define NBT_THREADS 256
#define NBT_CAP 2048      // staged candidates per window (16 B each)
#define NBT_CELLS 512     // box cells whose table entries are cached per batch (8^3 box)
#define NBT_CPT (NBT_CELLS / NBT_THREADS)

// Speculative scan of one cell segment: band candidates count as hits; returns
// true when any candidate fell into the band (the caller then restores kk and
// repeats the segment with nbt_scan_exact).  The caller guarantees room for the
// whole segment in the row.  The decision step is written in PTX so that it stays
// branch-free: 2 FSETP (the second predicated on the first), the self-exclusion,
// a predicated store and a predicated pointer bump.
__device__ __forceinline__ bool nbt_scan(int kbeg, int kend, const float4 *cand, float mx,
                                         float my, float mz, float thr_hi, float thr_lo,
                                         uint32_t e_self, uint32_t *__restrict__ base,
                                         uint32_t &kk) {
    uint32_t near_any = 0;
    uint32_t *ptr = base + (size_t)kk * 32u;
#pragma unroll 4
    for (int k = kbeg; k < kend; ++k) {
        const float4 c = cand[k];
        const float dx = c.x - mx, dy = c.y - my, dz = c.z - mz;
        const float d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
        asm volatile(
            "{\n\t"
            ".reg .pred p, q;\n\t"
            "setp.lt.f32 p, %2, %3;\n\t"             // inside (band included)
            "setp.gt.and.f32 q, %2, %4, p;\n\t"      // ... and in the band
            "@q mov.u32 %1, 1;\n\t"
            "setp.ne.and.u32 p, %5, %6, p;\n\t"      // not the centre itself
            "@p st.global.u32 [%0], %5;\n\t"
            "@p add.u64 %0, %0, 128;\n\t"
            "}"
            : "+l"(ptr), "+r"(near_any)
            : "f"(d2), "f"(thr_hi), "f"(thr_lo), "r"(__float_as_uint(c.w)), "r"(e_self)
            : "memory");
    }
    kk = (uint32_t)((ptr - base) >> 5);
    return near_any != 0u;
}

// float64 re-scan of a segment: exactly the test of for_each_neighbor()
__device__ __noinline__ void nbt_scan_exact(int kbeg, int kend, const float4 *cand,
                                            const Atom4 *__restrict__ atoms, int idx,
                                            uint32_t wcap, uint32_t *__restrict__ base,
                                            uint32_t &kk, const Grid &g, const ExactCtx &x) {
    const Atom4 me = atoms[idx];
    const double tol = 1e-9 * g.rc2;
    for (int k = kbeg; k < kend; ++k) {
        const uint32_t e = __float_as_uint(cand[k].w);
        const int j = (int)(e & TAB_COL_IDX_MASK);
        if (j == idx) continue;
        const Atom4 a = atoms[j];
        const double ddx = a.x - me.x, ddy = a.y - me.y, ddz = a.z - me.z;
        const double d2 = ddx * ddx + ddy * ddy + ddz * ddz;
        bool in = d2 < g.rc2;
        if (fabs(d2 - g.rc2) <= tol) in = exact_inside(g, x, idx, j);
        if (in && kk < wcap) base[(size_t)kk * 32u] = e;
        kk += in ? 1u : 0u;
    }
}

__global__ void __launch_bounds__(NBT_THREADS, 4)
k_nbr_tile(int n, Grid g, ExactCtx x, const Atom4 *__restrict__ atoms,
           const uint8_t *__restrict__ types_ext,
           const uint32_t *__restrict__ cell_start,
           const uint32_t *__restrict__ cell_count,
           const uint4 *__restrict__ ext_tab, uint32_t wcap, float band,
           int *__restrict__ counts, uint32_t *__restrict__ rows) {
    __shared__ float4 cand[NBT_CAP];        // x, y, z relative to the box centre; entry bits
    __shared__ uint4 cell_tab[NBT_CELLS];   // {start_a, count_a, start_b, count_b}
    __shared__ int cell_off[NBT_CELLS + 1]; // candidate offset of the cell in the batch
    __shared__ int warp_tot[NBT_THREADS / 32];
    __shared__ uint32_t tile_start[TAB_TILE_B_MAX * TAB_TILE_B_MAX * TAB_TILE_B_MAX + 1];

    const int B = g.tb, B3 = B * B * B;
    const int tile = blockIdx.x;
    const int tid = threadIdx.x;
    if (tid < B3) tile_start[tid] = cell_start[tile * B3 + tid];
    if (tid == B3 - 1)
        tile_start[B3] = cell_start[tile * B3 + tid] + cell_count[tile * B3 + tid];
    __syncthreads();
    const int a0 = (int)tile_start[0], a1 = (int)tile_start[B3];
    if (a1 <= a0) return;

    int t3[3];
    t3[0] = tile % g.tl[0];
    t3[1] = (tile / g.tl[0]) % g.tl[1];
    t3[2] = tile / (g.tl[0] * g.tl[1]);
    int b0[3], bn[3];
    double ctr[3] = {g.origin[0], g.origin[1], g.origin[2]};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        int lo = t3[k] * B - g.sr[k];
        int hi = min(t3[k] * B + B - 1, g.nb[k] - 1) + g.sr[k];
        if (!g.pbc[k]) {
            lo = max(lo, 0);
            hi = min(hi, g.nb[k] - 1);
        }
        b0[k] = lo;
        bn[k] = hi - lo + 1;
        // box centre in scaled coordinates along k -> Cartesian
        const double sc = 0.5 * (double)(lo + hi + 1) / (double)g.nb[k];
        ctr[0] += sc * g.h[3 * k + 0];
        ctr[1] += sc * g.h[3 * k + 1];
        ctr[2] += sc * g.h[3 * k + 2];
    }
    const int box_cells = bn[0] * bn[1] * bn[2];
    const int sr0 = g.sr[0], sr1 = g.sr[1], sr2 = g.sr[2];
    const float thr_hi = (float)g.rc2 + band, thr_lo = (float)g.rc2 - band;

    for (int chunk = a0; chunk < a1; chunk += NBT_THREADS) {
        const int idx = chunk + tid;
        const bool active = idx < a1;
        int mx = 0, my = 0, mz = 0;       // my cell, relative to the box origin
        float fx = 0.f, fy = 0.f, fz = 0.f;
        uint32_t *base = rows;
        uint32_t e_self = 0;
        if (active) {
            e_self = (uint32_t)idx | ((uint32_t)types_ext[idx] << TAB_COL_TYPE_SHIFT);
            // my cell of the tile: the last one that starts at or before idx (empty cells
            // share their start with the next one)
            int l = 0;
            for (int step = B3 >> 1; step > 0; step >>= 1)
                l += tile_start[l + step] <= (uint32_t)idx ? step : 0;
            mx = t3[0] * B + l % B - b0[0];
            my = t3[1] * B + (l / B) % B - b0[1];
            mz = t3[2] * B + l / (B * B) - b0[2];
            const Atom4 me = atoms[idx];
            fx = (float)(me.x - ctr[0]);      // same rounding as the staged record
            fy = (float)(me.y - ctr[1]);
            fz = (float)(me.z - ctr[2]);
            base = rows + ((size_t)(idx >> 5) * wcap * 32u + (idx & 31));
        }
        uint32_t kk = 0;                  // hits so far = next row slot

        for (int batch = 0; batch < box_cells; batch += NBT_CELLS) {
            const int batch_cells = min(NBT_CELLS, box_cells - batch);
            __syncthreads();              // previous users of the tables / buffers are done
            // every thread owns NBT_CPT box cells (c = tid, tid + NBT_THREADS, ...): table entries
            // (all loads issued before the first use), then an exclusive block scan of the
            // candidate counts, one round per group of NBT_THREADS cells
            uint4 tabs[NBT_CPT];
#pragma unroll
            for (int h = 0; h < NBT_CPT; ++h) {
                const int c = h * NBT_THREADS + tid;
                tabs[h] = make_uint4(0u, 0u, 0u, 0u);
                if (c < batch_cells) {
                    const int q = batch + c;
                    const int qx = q % bn[0], qy = (q / bn[0]) % bn[1], qz = q / (bn[0] * bn[1]);
                    tabs[h] = ext_tab[((b0[2] + qz + g.g[2]) * g.ne[1] + (b0[1] + qy + g.g[1])) *
                                          g.ne[0] + (b0[0] + qx + g.g[0])];
                }
            }
            int base_tot = 0;
#pragma unroll
            for (int h = 0; h < NBT_CPT; ++h) {
                const int c = h * NBT_THREADS + tid;
                const int mine = (int)(tabs[h].y + tabs[h].w);
                cell_tab[c] = tabs[h];
                int incl = mine;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int v = __shfl_up_sync(0xffffffffu, incl, d);
                    if ((tid & 31) >= d) incl += v;
                }
                if ((tid & 31) == 31) warp_tot[tid >> 5] = incl;
                __syncthreads();
                int wbase = 0, all = 0;
#pragma unroll
                for (int w = 0; w < NBT_THREADS / 32; ++w) {
                    wbase += w < (tid >> 5) ? warp_tot[w] : 0;
                    all += warp_tot[w];
                }
                cell_off[c] = base_tot + wbase + incl - mine;   // (= the total past batch_cells)
                base_tot += all;
                __syncthreads();          // warp_tot is reused by the next round
            }
            if (tid == 0) cell_off[NBT_CELLS] = base_tot;
            __syncthreads();
            const int total = cell_off[NBT_CELLS];

            for (int w0 = 0; w0 < total; w0 += NBT_CAP) {
                const int w1 = min(total, w0 + NBT_CAP);
                if (w0 > 0) __syncthreads();          // the previous window has been consumed
                // stage the window: one thread per candidate (a half-width cell holds ~4 atoms:
                // a warp per cell would idle 28 lanes and serialise 512 cells over 8 warps).  The
                // candidate's cell = the last one that starts at or before it (binary search in
                // the offsets; empty cells share their start with the next one).
                for (int u = w0 + tid; u < w1; u += NBT_THREADS) {
                    int lo = 0, hi = batch_cells;
                    while (hi - lo > 1) {
                        const int mid = (lo + hi) >> 1;
                        if (cell_off[mid] <= u) lo = mid;
                        else hi = mid;
                    }
                    const uint4 t = cell_tab[lo];
                    const int o = u - cell_off[lo];
                    const uint32_t j = o < (int)t.y ? t.x + (uint32_t)o
                                                    : t.z + (uint32_t)(o - (int)t.y);
                    const Atom4 a = atoms[j];
                    const uint32_t e = j | ((uint32_t)types_ext[j] << TAB_COL_TYPE_SHIFT);
                    cand[u - w0] = make_float4((float)(a.x - ctr[0]), (float)(a.y - ctr[1]),
                                               (float)(a.z - ctr[2]), __uint_as_float(e));
                }
                __syncthreads();
                if (active) {
                    // the neighbourhood row by row: for fixed (z, y) the cells x - sr .. x + sr
                    // are consecutive box cells, i.e. ONE contiguous run of staged candidates
                    // (clipped to this batch of cells and this window of candidates).  Rows in
                    // ascending (z, y): the candidate order of for_each_neighbor().
                    const int x_lo = max(mx - sr0, 0), x_hi = min(mx + sr0, bn[0] - 1);
                    for (int qz = max(mz - sr2, 0); qz <= min(mz + sr2, bn[2] - 1); ++qz) {
                        for (int qy = max(my - sr1, 0); qy <= min(my + sr1, bn[1] - 1); ++qy) {
                            const int q = (qz * bn[1] + qy) * bn[0];
                            const int c0 = max(q + x_lo - batch, 0);
                            const int c1 = min(q + x_hi + 1 - batch, batch_cells);
                            if (c1 <= c0) continue;
                            const int kbeg = max(cell_off[c0], w0) - w0;
                            const int kend = min(cell_off[c1], w1) - w0;
                            if (kend <= kbeg) continue;
                            const uint32_t kk0 = kk;
                            // fast path needs room for the whole run (rows have slack)
                            if (kk + (uint32_t)(kend - kbeg) > wcap ||
                                nbt_scan(kbeg, kend, cand, fx, fy, fz, thr_hi, thr_lo, e_self,
                                         base, kk)) {
                                kk = kk0;
                                nbt_scan_exact(kbeg, kend, cand, atoms, idx, wcap, base, kk, g, x);
                            }
                        }
                    }
                }
            }
        }
        if (active) counts[idx] = (int)kk;
    }
}

// several species: regroup the fixed-capacity rows of k_nbr_tile by neighbour
// species (stable) into the compact slices; also produces the per-species counts.
__global__ void __launch_bounds__(128)
k_rows_by_species(int n, int n_types, uint32_t wcap, const int *__restrict__ counts,
                  const uint32_t *__restrict__ rows, const uint32_t *__restrict__ slice_ptr,
                  uint32_t *__restrict__ col, int *__restrict__ tcounts) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    const uint32_t *src = rows + ((size_t)(idx >> 5) * wcap * 32u + (idx & 31));
    uint32_t *dst = col + ((size_t)slice_ptr[idx >> 5] * 32u + (idx & 31));
    const int cnt = counts[idx];
    uint32_t off[TAB_MAX_ELEMENTS];
    for (int t = 0; t < n_types; ++t) off[t] = 0;
    for (int k = 0; k < cnt; ++k) ++off[src[(size_t)k * 32u] >> TAB_COL_TYPE_SHIFT];
    uint32_t run = 0;
    for (int t = 0; t < n_types; ++t) {
        const uint32_t c = off[t];
        tcounts[(size_t)idx * n_types + t] = (int)c;
        off[t] = run;
        run += c;
    }
    for (int k = 0; k < cnt; ++k) {
        const uint32_t e = src[(size_t)k * 32u];
        dst[(size_t)(off[e >> TAB_COL_TYPE_SHIFT]++) * 32u] = e;
    }
}

// single species, fixed-capacity rows used in place: slice_ptr[s] = s * wcap
__global__ void k_fixed_slices(int n, int n_slices, uint32_t wcap,
                               const int *__restrict__ counts,
                               uint32_t *__restrict__ slice_ptr, int *__restrict__ tcounts) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_slices) slice_ptr[i] = (uint32_t)i * wcap;
    if (i < n) tcounts[i] = counts[i];
}

// slice widths (max count of the 32 atoms of a slice), nij, nnl_max; pads the
// columns of the non-existent atoms of the last slice
__global__ void k_slice_stats(int n, const int *__restrict__ counts,
                              uint32_t *__restrict__ slice_w,
                              unsigned long long *__restrict__ stats) {
    // grid-stride, one atomic pair per BLOCK (31 k same-address atomics serialise
    // in L2: the per-warp version took 43 us for 1 M atoms)
    __shared__ int s_mx[8];
    __shared__ unsigned long long s_sum[8];
    int bmx = 0;
    unsigned long long bsum = 0;
    const int n_pad = (n + 31) & ~31;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n_pad;
         idx += gridDim.x * blockDim.x) {
        const int cnt = idx < n ? counts[idx] : 0;
        int mx = cnt, sum = cnt;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, d));
            sum += __shfl_xor_sync(0xffffffffu, sum, d);
        }
        if ((threadIdx.x & 31) == 0) slice_w[idx >> 5] = (uint32_t)mx;
        bmx = max(bmx, mx);
        bsum += (unsigned long long)sum;
    }
    if ((threadIdx.x & 31) == 0) {
        s_mx[threadIdx.x >> 5] = bmx;
        s_sum[threadIdx.x >> 5] = bsum;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) {
            bmx = max(bmx, s_mx[w]);
            bsum += s_sum[w];
        }
        atomicAdd(&stats[0], bsum);
        atomicMax(&stats[1], (unsigned long long)bmx);
    }
}

// columns of the padding lanes of the last (partial) slice
__global__ void k_pad_tail(int n, const uint32_t *__restrict__ slice_w,
                           const uint32_t *__restrict__ slice_ptr,
                           uint32_t *__restrict__ col, uint32_t pad) {
    const int lane = threadIdx.x & 31;
    const int s = (n - 1) >> 5;
    if (s * 32 + lane < n) return;
    uint32_t *base = col + ((size_t)slice_ptr[s] * 32u + lane);
    for (uint32_t k = 0; k < slice_w[s]; ++k) base[(size_t)k * 32u] = pad;
}

// Rows that pair loops may read without looking at the counts (eam_fast.cuh): every slot
// past an atom's count up to the slice width holds the SENTINEL entry, and so do `extra` rows
// after the slice (fixed-stride rows: up to the stride; compact rows: only after the last slice,
// the rows after any other slice are the next slice's, initialised the same way).
__global__ void __launch_bounds__(128)
k_pad_sentinel(int n, int n_slices, uint32_t pad, uint32_t extra, uint32_t stride,
               const int *__restrict__ counts, const uint32_t *__restrict__ slice_w,
               const uint32_t *__restrict__ slice_ptr, uint32_t *__restrict__ col) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int s = idx >> 5;
    if (s >= n_slices) return;
    const uint32_t cnt = idx < n ? (uint32_t)counts[idx] : 0u;
    uint32_t hi = slice_w[s] + extra;
    if (s != n_slices - 1) hi = stride ? min(hi, stride) : slice_w[s];
    uint32_t *base = col + ((size_t)slice_ptr[s] * 32u + (idx & 31));
    for (uint32_t k = cnt; k < hi; ++k) base[(size_t)k * 32u] = pad;
}

__global__ void k_scatter_counts(int n, const int *__restrict__ perm,
                                 const int *__restrict__ counts,
                                 int *__restrict__ out) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < n) out[perm[idx]] = counts[idx];
}

__global__ void k_export(int n, int n_loc, const int *__restrict__ perm,
                         const int *__restrict__ s0,
                         const int *__restrict__ counts,
                         const uint32_t *__restrict__ slice_ptr,
                         const uint32_t *__restrict__ col,
                         const int *__restrict__ ghost_owner,
                         const int *__restrict__ ghost_S,
                         const uint32_t *__restrict__ row_ptr,
                         int *__restrict__ out_i, int *__restrict__ out_j,
                         int *__restrict__ out_S) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    const int oi = perm[idx];
    int ia, ib, ic;
    tab_unpack_shift(s0[oi], ia, ib, ic);
    const uint32_t *base = col + ((size_t)slice_ptr[idx >> 5] * 32u + (idx & 31));
    size_t o = row_ptr[oi];
    const int cnt = counts[idx];
    for (int k = 0; k < cnt; ++k, ++o) {
        const int j = (int)(base[(size_t)k * 32u] & TAB_COL_IDX_MASK);
        int owner = j, Sa = 0, Sb = 0, Sc = 0;
        if (j >= n_loc) {
            owner = ghost_owner[j - n_loc];
            tab_unpack_shift(ghost_S[j - n_loc], Sa, Sb, Sc);
        }
        const int oj = perm[owner];
        int ja, jb, jc;
        tab_unpack_shift(s0[oj], ja, jb, jc);
        out_i[o] = oi;
        out_j[o] = oj;
        out_S[3 * o + 0] = Sa - ja + ia;
        out_S[3 * o + 1] = Sb - jb + ib;
        out_S[3 * o + 2] = Sc - jc + ic;
    }
}

// rev[entry of (i -> j, S)] = position inside the row of owner(j) of the reverse
// pair (owner(j) -> image of i with shift -S).  Needed by models whose per-pair
// gradient g_p = dE_i/dD_p is not symmetric in the pair (symmetry functions):
//   F_i = sum_{p in row i} g_p - sum_{p in row i} g_rev(p).
#define TAB_REV_NONE 0xFFFFFFFFu
// one block per atom, one thread per row entry (the serial version took 2 ms for
// a 128-atom structure: 90 x 90 dependent loads in a single thread)
__global__ void __launch_bounds__(128)
k_build_reverse(int n, int n_loc, const int *__restrict__ counts,
                const uint32_t *__restrict__ slice_ptr,
                const uint32_t *__restrict__ col, const int *__restrict__ ghost_owner,
                const int *__restrict__ ghost_S, uint32_t *__restrict__ rev) {
    const int idx = blockIdx.x;
    if (idx >= n) return;
    const size_t base = (size_t)slice_ptr[idx >> 5] * 32u + (idx & 31);
    const int cnt = counts[idx];
    const int zero = tab_pack_shift(0, 0, 0);
    for (int k = threadIdx.x; k < cnt; k += blockDim.x) {
        const int j = (int)(col[base + (size_t)k * 32u] & TAB_COL_IDX_MASK);
        int o = j, S = zero;
        if (j >= n_loc) {
            o = ghost_owner[j - n_loc];
            S = ghost_S[j - n_loc];
        }
        uint32_t found = TAB_REV_NONE;
        if (o < n) {
            int a, b, c;
            tab_unpack_shift(S, a, b, c);
            const int want = tab_pack_shift(-a, -b, -c);
            const size_t obase = (size_t)slice_ptr[o >> 5] * 32u + (o & 31);
            const int ocnt = counts[o];
            for (int q = 0; q < ocnt; ++q) {
                const int e = (int)(col[obase + (size_t)q * 32u] & TAB_COL_IDX_MASK);
                int eo = e, eS = zero;
                if (e >= n_loc) {
                    eo = ghost_owner[e - n_loc];
                    eS = ghost_S[e - n_loc];
                }
                if (eo == idx && eS == want) {
                    found = (uint32_t)q;
                    break;
                }
            }
        }
        rev[base + (size_t)k * 32u] = found;
    }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
static void invert3(const double *h, double *inv, double *det_out) {
    const double a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5],
                 g = h[6], hh = h[7], i = h[8];
    const double det = a * (e * i - f * hh) - b * (d * i - f * g) + c * (d * hh - e * g);
    inv[0] = (e * i - f * hh) / det;
    inv[1] = (c * hh - b * i) / det;
    inv[2] = (b * f - c * e) / det;
    inv[3] = (f * g - d * i) / det;
    inv[4] = (a * i - c * g) / det;
    inv[5] = (c * d - a * f) / det;
    inv[6] = (d * hh - e * g) / det;
    inv[7] = (b * g - a * hh) / det;
    inv[8] = (a * e - b * d) / det;
    *det_out = det;
}

// `subdiv` = 1: cells at least rc wide, 27-cell neighbourhoods, tiles of 2 x 2 x 2 cells;
// `subdiv` = 2: cells at least rc / 2 wide, 125-cell neighbourhoods (15.6 rc^3 of candidates
// instead of 27 rc^3), tiles of 4 x 4 x 4 cells -- the same tile volume
static int setup_grid(Grid &g, int n, const double *h_cell, const double *h_origin,
                      const int *h_pbc, double rc, int subdiv) {
    g.tb = subdiv == 2 ? 4 : 2;
    const int B = g.tb;
    memcpy(g.h, h_cell, 9 * sizeof(double));
    for (int k = 0; k < 3; ++k) g.origin[k] = h_origin ? h_origin[k] : 0.0;
    double det;
    invert3(g.h, g.hinv, &det);
    if (!(fabs(det) > 1e-12)) {
        tab_set_error("cell is singular (det=%g); non-periodic structures must be "
                      "given a bounding cell by the caller", det);
        return TAB_EINVAL;
    }
    g.rc = rc;
    g.rc2 = rc * rc;
    const double rpad = rc * (1.0 + 1e-6);
    // face distance along direction k = 1 / |column k of hinv|
    long long cells = 1;
    double fd[3];
    for (int k = 0; k < 3; ++k) {
        const double nrm = sqrt(g.hinv[k] * g.hinv[k] + g.hinv[3 + k] * g.hinv[3 + k] +
                                g.hinv[6 + k] * g.hinv[6 + k]);
        fd[k] = 1.0 / nrm;
        g.pbc[k] = h_pbc[k] ? 1 : 0;
        int nb = (int)floor(fd[k] * subdiv / rpad);
        if (nb < 1) nb = 1;
        g.nb[k] = nb;
        cells *= nb;
    }
    // bound the table for dilute systems
    const long long cap = 8LL * n + 4096;
    while (cells > cap) {
        int kmax = 0;
        for (int k = 1; k < 3; ++k)
            if (g.nb[k] > g.nb[kmax]) kmax = k;
        cells /= g.nb[kmax];
        g.nb[kmax] = (g.nb[kmax] + 1) / 2;
        cells *= g.nb[kmax];
    }
    long long ecells = 1, slots = 1;
    for (int k = 0; k < 3; ++k) {
        const double w = fd[k] / g.nb[k];
        int sr = (int)ceil(rpad / w);
        if (sr < 1) sr = 1;
        if (!g.pbc[k] && g.nb[k] == 1) sr = 0;
        g.sr[k] = sr;
        g.g[k] = g.pbc[k] ? sr : 0;
        g.ne[k] = g.nb[k] + 2 * g.g[k];
        g.tl[k] = (g.nb[k] + B - 1) / B;
        ecells *= g.ne[k];
        slots *= (long long)g.tl[k] * B;
        if (sr > 500) {
            tab_set_error("cutoff %g spans %d images of the cell: unsupported", rc, sr);
            return TAB_EUNSUPPORTED;
        }
    }
    if (ecells > 0x3fffffffLL || slots > 0x1fffffffLL) {
        tab_set_error("cell table too large");
        return TAB_EUNSUPPORTED;
    }
    g.n_ecells = (int)ecells;
    g.n_slots = (int)slots;
    return TAB_OK;
}

extern "C" int tab_nbr_create(tab_nbr **out) {
    if (!out) return TAB_EINVAL;
    *out = new tab_nbr();
    return TAB_OK;
}

extern "C" int tab_nbr_free(tab_nbr *nbr) {
    if (!nbr) return TAB_OK;
    DevBuf *bufs[] = {&nbr->cell_of, &nbr->s0, &nbr->types_in, &nbr->perm, &nbr->counts,
                      &nbr->atoms, &nbr->types_ext, &nbr->ghost_owner, &nbr->ghost_S,
                      &nbr->cell_count, &nbr->cell_start, &nbr->cell_fill,
                      &nbr->ext_tab, &nbr->gcount, &nbr->gstart,
                      &nbr->slice_w, &nbr->slice_ptr, &nbr->col, &nbr->scan_tmp,
                      &nbr->stats, &nbr->row_ptr, &nbr->rho, &nbr->partial, &nbr->adp,
                      &nbr->tcounts, &nbr->rev, &nbr->pcache, &nbr->rows_tmp,
                      &nbr->pos_ref, &nbr->disp, &nbr->ls_ptr, &nbr->ls_col, &nbr->ls_w,
                      &nbr->rec16};
    for (DevBuf *b : bufs) b->release();
    if (nbr->h_pin) cudaFreeHost(nbr->h_pin);
    delete nbr;
    return TAB_OK;
}

// page-locked target of the build's small read-backs (a copy into pageable memory is staged by
// the driver and blocks the calling thread longer)
static int nbr_pinned(tab_nbr *nbr, unsigned long long **out) {
    if (!nbr->h_pin) {
        cudaError_t e = cudaHostAlloc((void **)&nbr->h_pin, 8 * sizeof(unsigned long long),
                                      cudaHostAllocDefault);
        if (e != cudaSuccess) {
            nbr->h_pin = nullptr;
            tab_set_error("cudaHostAlloc -> %s", cudaGetErrorString(e));
            return TAB_ENOMEM;
        }
    }
    *out = nbr->h_pin;
    return TAB_OK;
}

static inline int nblocks(long long n, int t) { return (int)((n + t - 1) / t); }

static QFrame qframe_of(const tab_nbr *nbr) {
    QFrame q;
    q.ox = nbr->q_origin[0];
    q.oy = nbr->q_origin[1];
    q.oz = nbr->q_origin[2];
    q.inv_delta = nbr->q_inv_delta;
    return q;
}

// Fixed-point frame of the float32 records: Cartesian bounding box of the binning frame
// plus the reach of the ghost images and some room for drift; delta = power of two.
static void setup_qframe(tab_nbr *nbr) {
    const Grid &g = nbr->grid;
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int c = 0; c < 8; ++c)
        for (int k = 0; k < 3; ++k) {
            const double v = g.origin[k] + ((c & 1) ? g.h[k] : 0.0) + ((c & 2) ? g.h[3 + k] : 0.0) +
                             ((c & 4) ? g.h[6 + k] : 0.0);
            lo[k] = fmin(lo[k], v);
            hi[k] = fmax(hi[k], v);
        }
    const double margin = 2.0 * g.rc + 8.0;
    double ext = 0.0;
    for (int k = 0; k < 3; ++k) {
        nbr->q_origin[k] = lo[k] - margin;
        ext = fmax(ext, hi[k] - lo[k] + 2.0 * margin);
    }
    int e;
    frexp(ext / 1073741824.0, &e);            // ext / 2^30 = m 2^e, m in [0.5, 1)
    nbr->q_delta = ldexp(1.0, e);             // >= ext / 2^30: coordinates fit 31 bits
    nbr->q_inv_delta = ldexp(1.0, -e);
}

// The sentinel record (extended index n_ext): a point far outside the frame, so that every
// pair with it fails the cutoff mask.  Fixed point: all real coordinates are >= 0, the
// sentinel sits at -2^30 per axis (differences stay inside int32).
__global__ void k_sentinels(int n_ext, double ox, double oy, double oz, Atom4 *atoms,
                            Rec16 *rec16) {
    Atom4 a;
    a.x = ox - 1.0e4;
    a.y = oy - 1.0e4;
    a.z = oz - 1.0e4;
    a.w = 0.0;
    atoms[n_ext] = a;
    if (rec16) {
        Rec16 r;
        r.qx = r.qy = r.qz = -(1 << 30);
        r.w = 0.f;
        rec16[n_ext] = r;
    }
}

static int write_sentinels(tab_nbr *nbr, bool with_rec16, cudaStream_t st) {
    const Grid &g = nbr->grid;
    k_sentinels<<<1, 1, 0, st>>>(nbr->n_ext, g.origin[0], g.origin[1], g.origin[2],
                                 nbr->atoms.as<Atom4>(),
                                 with_rec16 ? nbr->rec16.as<Rec16>() : nullptr);
    TAB_LAUNCH_CHECK();
    return TAB_OK;
}

__global__ void k_make_rec16(int n_ext, const Atom4 *__restrict__ atoms, QFrame qf,
                             Rec16 *__restrict__ rec16) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_ext) return;
    const Atom4 a = atoms[e];
    rec16[e] = make_rec16(qf, a.x, a.y, a.z);
}

// float32 records of the current positions (first use on a handle; afterwards the build and
// the refresh kernels keep them up to date)
int tab_nbr_ensure_rec16(tab_nbr *nbr, cudaStream_t st) {
    if (nbr->rec16_valid) return TAB_OK;
    nbr->want_rec16 = true;
    setup_qframe(nbr);
    TAB_TRY(nbr->rec16.ensure(sizeof(Rec16) * ((size_t)nbr->n_ext + 1)));
    TAB_TRY(write_sentinels(nbr, true, st));
    k_make_rec16<<<(nbr->n_ext + 255) / 256, 256, 0, st>>>(nbr->n_ext, nbr->atoms.as<Atom4>(),
                                                          qframe_of(nbr), nbr->rec16.as<Rec16>());
    TAB_LAUNCH_CHECK();
    nbr->rec16_valid = true;
    return TAB_OK;
}

static int refresh_positions(tab_nbr *nbr, const double *d_pos, cudaStream_t st) {
    const Grid &g = nbr->grid;
    nbr->pcache_valid = false;
    const bool track = nbr->skin_built > 0.0;
    Rec16 *rec = nbr->rec16_valid ? nbr->rec16.as<Rec16>() : nullptr;
    const QFrame qf = qframe_of(nbr);
    if (track) {
        TAB_CUDA(cudaMemsetAsync(nbr->disp.p, 0, 4 * sizeof(unsigned int), st));
        k_gather_owned<2><<<nblocks(nbr->n_loc, 256), 256, 0, st>>>(
            nbr->n_loc, d_pos, nullptr, g, nbr->perm.as<int>(), nbr->s0.as<int>(),
            nbr->atoms.as<Atom4>(), nullptr, nbr->pos_ref.as<double>(),
            nbr->disp.as<unsigned int>(), qf, rec);
    } else {
        k_gather_owned<0><<<nblocks(nbr->n_loc, 256), 256, 0, st>>>(
            nbr->n_loc, d_pos, nullptr, g, nbr->perm.as<int>(), nbr->s0.as<int>(),
            nbr->atoms.as<Atom4>(), nullptr, nullptr, nullptr, qf, rec);
    }
    TAB_LAUNCH_CHECK();
    if (nbr->n_ghost > 0) {
        k_refresh_ghosts<<<nblocks(nbr->n_ghost, 256), 256, 0, st>>>(
            nbr->n_ghost, nbr->n_loc, g, nbr->ghost_owner.as<int>(), nbr->ghost_S.as<int>(),
            nbr->atoms.as<Atom4>(), qf, rec);
        TAB_LAUNCH_CHECK();
    }
    return TAB_OK;
}

extern "C" int tab_nbr_build_dd(tab_nbr *nbr, int32_t n_owned, int32_t n_halo,
                                const double *d_pos, const int32_t *d_types,
                                const double *h_cell, const double *h_origin,
                                const int32_t *h_pbc, double rc, void *stream) {
    if (!nbr || n_owned <= 0 || n_halo < 0 || !d_pos || !h_cell || !h_pbc || !(rc > 0)) {
        tab_set_error("tab_nbr_build: bad argument");
        return TAB_EINVAL;
    }
    const long long n_loc_ll = (long long)n_owned + n_halo;
    if (n_loc_ll > (long long)(TAB_COL_IDX_MASK / 2)) {
        tab_set_error("tab_nbr_build: too many atoms for one device (%lld)", n_loc_ll);
        return TAB_EUNSUPPORTED;
    }
    cudaStream_t st = (cudaStream_t)stream;
    nbr->built = false;
    nbr->pcache_valid = false;
    nbr->n_struct = 0;          // single structure (a batch handle may be reused)
    nbr->has_row_ptr = false;
    nbr->ls_L = 0;
    nbr->col_padded = false;
    nbr->rec16_valid = false;
    // lists with a skin: radius rc + skin, the pair kernels mask r >= rc
    nbr->rc_model = rc;
    nbr->skin_built = nbr->skin;
    rc += nbr->skin;
    Grid &g = nbr->grid;
    const int n = n_owned, n_loc = (int)n_loc_ll;
    // large systems (the tile kernel): half-width cells.  TAB_NBR_SUBDIV=1|2 overrides (A/B)
    int subdiv = n_loc > 20000 ? 2 : 1;
    if (const char *env = getenv("TAB_NBR_SUBDIV")) subdiv = atoi(env) == 2 ? 2 : 1;
    TAB_TRY(setup_grid(g, n_loc, h_cell, h_origin, h_pbc, rc, subdiv));
    nbr->n = n;
    nbr->n_halo = n_halo;
    nbr->n_loc = n_loc;
    nbr->n_slices = (n + 31) / 32;
    const int slots2 = 2 * g.n_slots;

    TAB_TRY(nbr->cell_of.ensure(sizeof(int) * n_loc));
    TAB_TRY(nbr->s0.ensure(sizeof(int) * n_loc));
    TAB_TRY(nbr->perm.ensure(sizeof(int) * n_loc));
    TAB_TRY(nbr->counts.ensure(sizeof(int) * n));
    TAB_TRY(nbr->cell_count.ensure(sizeof(uint32_t) * slots2));
    TAB_TRY(nbr->cell_start.ensure(sizeof(uint32_t) * slots2));
    TAB_TRY(nbr->cell_fill.ensure(sizeof(uint32_t) * slots2));
    TAB_TRY(nbr->ext_tab.ensure(sizeof(uint4) * g.n_ecells));
    TAB_TRY(nbr->gcount.ensure(sizeof(uint32_t) * g.n_ecells));
    TAB_TRY(nbr->gstart.ensure(sizeof(uint32_t) * g.n_ecells));
    TAB_TRY(nbr->slice_w.ensure(sizeof(uint32_t) * nbr->n_slices));
    TAB_TRY(nbr->slice_ptr.ensure(sizeof(uint32_t) * nbr->n_slices));
    TAB_TRY(nbr->stats.ensure(5 * sizeof(unsigned long long)));
    unsigned long long *d_stats = nbr->stats.as<unsigned long long>();

    TAB_CUDA(cudaMemsetAsync(nbr->cell_count.p, 0, sizeof(uint32_t) * slots2, st));
    TAB_CUDA(cudaMemsetAsync(nbr->cell_fill.p, 0, sizeof(uint32_t) * slots2, st));
    TAB_CUDA(cudaMemsetAsync(d_stats, 0, 5 * sizeof(unsigned long long), st));

    k_bin<<<nblocks(n_loc, 256), 256, 0, st>>>(n_loc, n, d_pos, d_types, g,
                                               nbr->cell_of.as<int>(), nbr->s0.as<int>(),
                                               nbr->cell_count.as<uint32_t>(), d_stats);
    TAB_LAUNCH_CHECK();
    TAB_TRY(tab_scan_exclusive_u32(nbr->cell_count.as<uint32_t>(),
                                   nbr->cell_start.as<uint32_t>(), slots2, nullptr,
                                   nbr->scan_tmp, st));
    TAB_TRY(nbr->row_ptr.ensure(sizeof(int) * (size_t)n_loc));      // scratch: raw order
    k_scatter<<<nblocks(n_loc, 256), 256, 0, st>>>(n_loc, nbr->cell_of.as<int>(),
                                                   nbr->cell_start.as<uint32_t>(),
                                                   nbr->cell_fill.as<uint32_t>(),
                                                   nbr->row_ptr.as<int>());
    TAB_LAUNCH_CHECK();
    k_rank_in_cell<<<nblocks(n_loc, 256), 256, 0, st>>>(
        n_loc, nbr->cell_of.as<int>(), nbr->cell_start.as<uint32_t>(),
        nbr->cell_count.as<uint32_t>(), nbr->row_ptr.as<int>(), nbr->perm.as<int>());
    TAB_LAUNCH_CHECK();

    // ghost bookkeeping -> n_ghost (one small read-back)
    k_ext_cells<<<nblocks(g.n_ecells, 256), 256, 0, st>>>(
        g, nbr->cell_start.as<uint32_t>(), nbr->cell_count.as<uint32_t>(),
        nbr->ext_tab.as<uint4>(), nbr->gcount.as<uint32_t>());
    TAB_LAUNCH_CHECK();
    TAB_TRY(tab_scan_exclusive_u32(nbr->gcount.as<uint32_t>(),
                                   nbr->gstart.as<uint32_t>(), g.n_ecells, d_stats + 2,
                                   nbr->scan_tmp, st));
    unsigned long long *h_pin;
    TAB_TRY(nbr_pinned(nbr, &h_pin));
    unsigned long long *two = h_pin;         // [n_ghost, max type, atoms far outside the frame]
    TAB_CUDA(cudaMemcpyAsync(two, d_stats + 2, 3 * sizeof(unsigned long long),
                             cudaMemcpyDeviceToHost, st));
    TAB_CUDA(cudaStreamSynchronize(st));
    const unsigned long long n_ghost = two[0];
    nbr->n_types = (int)two[1] + 1;
    nbr->has_rev = false;
    if (nbr->n_types > TAB_MAX_ELEMENTS) {
        tab_set_error("element index %d exceeds the supported maximum", nbr->n_types - 1);
        return TAB_EINVAL;
    }
    if ((unsigned long long)n_loc + n_ghost > TAB_COL_IDX_MASK) {
        tab_set_error("owned + ghost atoms exceed the 28-bit index space");
        return TAB_EUNSUPPORTED;
    }
    nbr->n_ghost = (int)n_ghost;
    nbr->n_ext = n_loc + nbr->n_ghost;
    // one record past the end: the SENTINEL that pads the lane-split rows (eam_fast.cuh)
    TAB_TRY(nbr->atoms.ensure(sizeof(Atom4) * ((size_t)nbr->n_ext + 1)));
    TAB_TRY(nbr->types_ext.ensure((size_t)nbr->n_ext + 16));
    TAB_TRY(nbr->ghost_owner.ensure(sizeof(int) * (size_t)(nbr->n_ghost + 1)));
    TAB_TRY(nbr->ghost_S.ensure(sizeof(int) * (size_t)(nbr->n_ghost + 1)));

    Rec16 *rec = nullptr;
    if (nbr->want_rec16) {
        setup_qframe(nbr);
        TAB_TRY(nbr->rec16.ensure(sizeof(Rec16) * ((size_t)nbr->n_ext + 1)));
        rec = nbr->rec16.as<Rec16>();
    }
    TAB_TRY(write_sentinels(nbr, rec != nullptr, st));
    const QFrame qf = qframe_of(nbr);
    if (nbr->skin_built > 0.0) {
        TAB_TRY(nbr->pos_ref.ensure(sizeof(double) * 3 * (size_t)n_loc));
        TAB_TRY(nbr->disp.ensure(4 * sizeof(unsigned int)));
        TAB_CUDA(cudaMemsetAsync(nbr->disp.p, 0, 4 * sizeof(unsigned int), st));
    }
    k_gather_owned<1><<<nblocks(n_loc, 256), 256, 0, st>>>(
        n_loc, d_pos, d_types, g, nbr->perm.as<int>(), nbr->s0.as<int>(),
        nbr->atoms.as<Atom4>(), nbr->types_ext.as<uint8_t>(),
        nbr->skin_built > 0.0 ? nbr->pos_ref.as<double>() : nullptr, nullptr, qf, rec);
    TAB_LAUNCH_CHECK();
    if (nbr->n_ghost > 0) {
        k_fill_ghosts<<<nblocks((long long)g.n_ecells * (g.tb == 4 ? 8 : 32), 256), 256, 0, st>>>(
            g, n_loc, nbr->cell_start.as<uint32_t>(), nbr->cell_count.as<uint32_t>(),
            nbr->gstart.as<uint32_t>(), nbr->gcount.as<uint32_t>(),
            nbr->ext_tab.as<uint4>(), nbr->atoms.as<Atom4>(),
            nbr->types_ext.as<uint8_t>(), nbr->ghost_owner.as<int>(),
            nbr->ghost_S.as<int>(), 0, qf, rec);
        TAB_LAUNCH_CHECK();
    }
    nbr->rec16_valid = rec != nullptr;

    ExactCtx x;
    x.pos = d_pos;
    x.perm = nbr->perm.as<int>();
    x.s0 = nbr->s0.as<int>();
    x.ghost_owner = nbr->ghost_owner.as<int>();
    x.ghost_S = nbr->ghost_S.as<int>();
    x.n_loc = n_loc;
    const int nthreads = nbr->n_slices * 32;
    // small systems: one warp per atom (parallelism); large: one block per tile of
    // cells, single pass (thread per atom with shared-memory candidate staging).
    // (TAB_NBR_MODE=thread|tile|warp overrides, for A/B measurements)
    bool warp_mode = n_loc <= 20000;
    // tile kernel: box indices fit 8 bits; float32 pre-filter band from the largest
    // staged coordinate R (half extent of a tile's candidate box + one bin):
    //   |d2_f32 - d2| <= 2 sqrt(3) rc (2 R + rc) eps + 4 rc^2 eps,  eps = 2^-24
    // (doubled for safety).  A band wider than 1e-3 rc^2 would re-scan too often.
    bool tile_ok = g.sr[0] <= 100 && g.sr[1] <= 100 && g.sr[2] <= 100 && two[2] == 0;
    float band = 0.f;
    {
        double R = 0.0;
        for (int k = 0; k < 3; ++k) {
            const double len = sqrt(g.h[3 * k] * g.h[3 * k] + g.h[3 * k + 1] * g.h[3 * k + 1] +
                                    g.h[3 * k + 2] * g.h[3 * k + 2]);
            R += 0.5 * len * (double)(g.tb + 2 * g.sr[k] + 2) / (double)g.nb[k];
        }
        const double eps = 5.97e-8;
        const double bw = 2.0 * (2.0 * 1.7321 * rc * (2.0 * R + rc) * eps + 4.0 * g.rc2 * eps);
        if (bw > 1e-3 * g.rc2) tile_ok = false;
        band = (float)bw;
    }
    bool tile_mode = !warp_mode && tile_ok;
    if (const char *env = getenv("TAB_NBR_MODE")) {
        warp_mode = !strcmp(env, "warp");
        tile_mode = !strcmp(env, "tile") && tile_ok;
    }
    const int n_tiles = g.n_slots / (g.tb * g.tb * g.tb);
    TAB_TRY(nbr->tcounts.ensure(sizeof(int) * (size_t)n * nbr->n_types));
    unsigned long long *h_stats = h_pin + 4;

    if (tile_mode) {
        // row capacity: last build of this handle, else the mean density
        const bool multi = nbr->n_types > 1;
        uint32_t wcap;
        // (valid while the atom count stays within 5 %: a rank of a spatial decomposition gains
        // and loses a few atoms at every rebuild)
        if (nbr->wcap_hint > 0 && abs(nbr->wcap_hint_n - n_loc) * 20 <= n_loc) wcap = nbr->wcap_hint;
        else {
            const double vol = fabs(g.h[0] * (g.h[4] * g.h[8] - g.h[5] * g.h[7]) -
                                    g.h[1] * (g.h[3] * g.h[8] - g.h[5] * g.h[6]) +
                                    g.h[2] * (g.h[3] * g.h[7] - g.h[4] * g.h[6]));
            const double expect = 4.18879 * rc * rc * rc * (double)n_loc / vol;
            wcap = (uint32_t)fmin(expect * 1.25 + 16.0, 4096.0);
        }
        wcap = (wcap + 7u) & ~7u;
        for (int attempt = 0;; ++attempt) {
            const size_t row_words = ((size_t)nbr->n_slices * wcap + TAB_SPARE_ROWS) * 32u + 32u;
            DevBuf &rows = multi ? nbr->rows_tmp : nbr->col;
            TAB_TRY(rows.ensure(sizeof(uint32_t) * row_words));
            k_nbr_tile<<<n_tiles, NBT_THREADS, 0, st>>>(
                n, g, x, nbr->atoms.as<Atom4>(), nbr->types_ext.as<uint8_t>(),
                nbr->cell_start.as<uint32_t>(), nbr->cell_count.as<uint32_t>(),
                nbr->ext_tab.as<uint4>(), wcap, band, nbr->counts.as<int>(),
                rows.as<uint32_t>());
            TAB_LAUNCH_CHECK();
            k_slice_stats<<<min(nblocks(nthreads, 256), 1184), 256, 0, st>>>(
                n, nbr->counts.as<int>(), nbr->slice_w.as<uint32_t>(), d_stats);
            TAB_LAUNCH_CHECK();
            if (multi)
                TAB_TRY(tab_scan_exclusive_u32(nbr->slice_w.as<uint32_t>(),
                                               nbr->slice_ptr.as<uint32_t>(), nbr->n_slices,
                                               d_stats + 2, nbr->scan_tmp, st));
            TAB_CUDA(cudaMemcpyAsync(h_stats, d_stats, 3 * sizeof(unsigned long long),
                                     cudaMemcpyDeviceToHost, st));
            TAB_CUDA(cudaStreamSynchronize(st));
            if (h_stats[1] <= wcap) break;
            if (attempt > 0) {
                tab_set_error("tab_nbr_build: row capacity retry failed (%llu > %u)",
                              h_stats[1], wcap);
                return TAB_ESTATE;
            }
            wcap = ((uint32_t)h_stats[1] + 7u) & ~7u;      // exact width: counting never stops
            TAB_CUDA(cudaMemsetAsync(d_stats, 0, 3 * sizeof(unsigned long long), st));
        }
        nbr->nij = (long long)h_stats[0];
        nbr->nnl_max = (int)h_stats[1];
        nbr->ell_rows = multi ? (long long)h_stats[2] : (long long)nbr->n_slices * wcap;
        if (nbr->ell_rows > 0xffffffffLL) {
            tab_set_error("neighbour table too large (%lld rows)", nbr->ell_rows);
            return TAB_EUNSUPPORTED;
        }
        if (multi) {
            TAB_TRY(nbr->col.ensure(sizeof(uint32_t) * 32 * (size_t)(nbr->ell_rows + 1)));
            k_rows_by_species<<<nblocks(n, 128), 128, 0, st>>>(
                n, nbr->n_types, wcap, nbr->counts.as<int>(), nbr->rows_tmp.as<uint32_t>(),
                nbr->slice_ptr.as<uint32_t>(), nbr->col.as<uint32_t>(),
                nbr->tcounts.as<int>());
        } else {
            k_fixed_slices<<<nblocks(max(n, nbr->n_slices), 256), 256, 0, st>>>(
                n, nbr->n_slices, wcap, nbr->counts.as<int>(), nbr->slice_ptr.as<uint32_t>(),
                nbr->tcounts.as<int>());
            TAB_LAUNCH_CHECK();
            // fixed-stride rows: sentinel entries past the counts (and the spare rows, which for
            // the last slice lie beyond the stride: the buffer holds TAB_SPARE_ROWS more)
            k_pad_sentinel<<<nblocks(nthreads, 128), 128, 0, st>>>(
                n, nbr->n_slices, (uint32_t)nbr->n_ext, TAB_SPARE_ROWS, wcap,
                nbr->counts.as<int>(), nbr->slice_w.as<uint32_t>(),
                nbr->slice_ptr.as<uint32_t>(), nbr->col.as<uint32_t>());
            nbr->col_padded = true;
        }
        TAB_LAUNCH_CHECK();
        nbr->wcap_hint = (uint32_t)(nbr->nnl_max + nbr->nnl_max / 8 + 24);
        nbr->wcap_hint_n = n_loc;
        nbr->built = true;
        return TAB_OK;
    }

    if (warp_mode) {
        k_nbr_warp<false><<<nblocks((long long)n * 32, 128), 128, 0, st>>>(
            n, g, x, nbr->cell_of.as<int>(), nbr->atoms.as<Atom4>(),
            nbr->types_ext.as<uint8_t>(), nbr->ext_tab.as<uint4>(), nbr->n_types,
            nbr->counts.as<int>(), nbr->tcounts.as<int>(), nullptr, nullptr, nullptr, 0u);
    } else {
        k_count<<<nblocks(n, 128), 128, 0, st>>>(
            n, g, x, nbr->cell_of.as<int>(), nbr->atoms.as<Atom4>(),
            nbr->ext_tab.as<uint4>(), nbr->types_ext.as<uint8_t>(), nbr->n_types,
            nbr->counts.as<int>(), nbr->tcounts.as<int>());
    }
    TAB_LAUNCH_CHECK();
    k_slice_stats<<<min(nblocks(nthreads, 256), 1184), 256, 0, st>>>(
        n, nbr->counts.as<int>(), nbr->slice_w.as<uint32_t>(), d_stats);
    TAB_LAUNCH_CHECK();
    TAB_TRY(tab_scan_exclusive_u32(nbr->slice_w.as<uint32_t>(),
                                   nbr->slice_ptr.as<uint32_t>(), nbr->n_slices,
                                   d_stats + 2, nbr->scan_tmp, st));
    TAB_CUDA(cudaMemcpyAsync(h_stats, d_stats, 3 * sizeof(unsigned long long),
                             cudaMemcpyDeviceToHost, st));
    TAB_CUDA(cudaStreamSynchronize(st));
    nbr->nij = (long long)h_stats[0];
    nbr->nnl_max = (int)h_stats[1];
    nbr->ell_rows = (long long)h_stats[2];
    if (nbr->ell_rows > 0xffffffffLL) {
        tab_set_error("neighbour table too large (%lld rows)", nbr->ell_rows);
        return TAB_EUNSUPPORTED;
    }
    TAB_TRY(nbr->col.ensure(sizeof(uint32_t) * 32 * (size_t)(nbr->ell_rows + 1 + TAB_SPARE_ROWS)));
    if (warp_mode) {
        k_nbr_warp<true><<<nblocks((long long)n * 32, 128), 128, 0, st>>>(
            n, g, x, nbr->cell_of.as<int>(), nbr->atoms.as<Atom4>(),
            nbr->types_ext.as<uint8_t>(), nbr->ext_tab.as<uint4>(), nbr->n_types,
            nbr->counts.as<int>(), nbr->tcounts.as<int>(), nbr->slice_w.as<uint32_t>(),
            nbr->slice_ptr.as<uint32_t>(), nbr->col.as<uint32_t>(), (uint32_t)nbr->n_ext);
        TAB_LAUNCH_CHECK();
        if (n & 31) {
            k_pad_tail<<<1, 32, 0, st>>>(n, nbr->slice_w.as<uint32_t>(),
                                         nbr->slice_ptr.as<uint32_t>(),
                                         nbr->col.as<uint32_t>(), (uint32_t)nbr->n_ext);
            TAB_LAUNCH_CHECK();
        }
    } else {
        k_fill<<<nblocks(nthreads, 128), 128, 0, st>>>(
            n, g, x, nbr->cell_of.as<int>(), nbr->atoms.as<Atom4>(),
            nbr->types_ext.as<uint8_t>(), nbr->ext_tab.as<uint4>(), nbr->n_types,
            nbr->tcounts.as<int>(), nbr->slice_w.as<uint32_t>(),
            nbr->slice_ptr.as<uint32_t>(), nbr->col.as<uint32_t>(), (uint32_t)nbr->n_ext);
        TAB_LAUNCH_CHECK();
    }
    // the spare rows after the last slice
    k_pad_sentinel<<<nblocks(nthreads, 128), 128, 0, st>>>(
        n, nbr->n_slices, (uint32_t)nbr->n_ext, TAB_SPARE_ROWS, 0u, nbr->counts.as<int>(),
        nbr->slice_w.as<uint32_t>(), nbr->slice_ptr.as<uint32_t>(), nbr->col.as<uint32_t>());
    TAB_LAUNCH_CHECK();
    nbr->col_padded = true;
    nbr->built = true;
    return TAB_OK;
}

extern "C" int tab_nbr_build(tab_nbr *nbr, int32_t n, const double *d_pos,
                             const int32_t *d_types, const double *h_cell,
                             const int32_t *h_pbc, double rc, void *stream) {
    return tab_nbr_build_dd(nbr, n, 0, d_pos, d_types, h_cell, nullptr, h_pbc, rc, stream);
}

extern "C" int tab_nbr_update(tab_nbr *nbr, const double *d_pos, const double *h_cell,
                              void *stream) {
    if (!nbr || !d_pos) return TAB_EINVAL;
    if (!nbr->built) {
        tab_set_error("tab_nbr_update before tab_nbr_build");
        return TAB_ESTATE;
    }
    if (nbr->n_struct > 0) {
        tab_set_error("tab_nbr_update: batch handles are rebuilt, not refreshed");
        return TAB_EUNSUPPORTED;
    }
    if (h_cell) {
        double det;
        memcpy(nbr->grid.h, h_cell, 9 * sizeof(double));
        invert3(nbr->grid.h, nbr->grid.hinv, &det);
    }
    return refresh_positions(nbr, d_pos, (cudaStream_t)stream);
}

extern "C" int tab_nbr_set_skin(tab_nbr *nbr, double skin) {
    if (!nbr || !(skin >= 0.0)) {
        tab_set_error("tab_nbr_set_skin: bad argument");
        return TAB_EINVAL;
    }
    nbr->skin = skin;
    return TAB_OK;
}

extern "C" int tab_nbr_max_displacement(tab_nbr *nbr, double *h_disp, double *h_skin,
                                        void *stream) {
    if (!nbr || !h_disp) return TAB_EINVAL;
    if (!nbr->built) {
        tab_set_error("tab_nbr_max_displacement before tab_nbr_build");
        return TAB_ESTATE;
    }
    if (h_skin) *h_skin = nbr->skin_built;
    if (!(nbr->skin_built > 0.0)) {
        *h_disp = 0.0;          // lists without a skin: every update needs a rebuild anyway
        return TAB_OK;
    }
    unsigned int bits = 0;
    cudaStream_t st = (cudaStream_t)stream;
    TAB_CUDA(cudaMemcpyAsync(&bits, nbr->disp.p, sizeof(bits), cudaMemcpyDeviceToHost, st));
    TAB_CUDA(cudaStreamSynchronize(st));
    float d2;
    memcpy(&d2, &bits, sizeof(d2));
    *h_disp = d2 == d2 ? sqrt((double)d2) : (double)INFINITY;     // NaN positions: rebuild
    return TAB_OK;
}

extern "C" int tab_nbr_sizes(const tab_nbr *nbr, int64_t *nij, int32_t *nnl_max,
                             int32_t *n_ext) {
    if (!nbr) return TAB_EINVAL;
    if (!nbr->built) {
        tab_set_error("tab_nbr_sizes before tab_nbr_build");
        return TAB_ESTATE;
    }
    if (nij) *nij = nbr->nij;
    if (nnl_max) *nnl_max = nbr->nnl_max;
    if (n_ext) *n_ext = nbr->n_ext;
    return TAB_OK;
}

extern "C" int tab_nbr_counts(const tab_nbr *nbr, int32_t *d_counts, void *stream) {
    if (!nbr || !d_counts) return TAB_EINVAL;
    if (!nbr->built) return TAB_ESTATE;
    k_scatter_counts<<<nblocks(nbr->n, 256), 256, 0, (cudaStream_t)stream>>>(
        nbr->n, nbr->perm.as<int>(), nbr->counts.as<int>(), d_counts);
    TAB_LAUNCH_CHECK();
    return TAB_OK;
}

extern "C" int tab_nbr_export(const tab_nbr *cnbr, int32_t *d_i, int32_t *d_j,
                              int32_t *d_S, void *stream) {
    tab_nbr *nbr = const_cast<tab_nbr *>(cnbr);
    if (!nbr || !d_i || !d_j || !d_S) return TAB_EINVAL;
    if (!nbr->built) return TAB_ESTATE;
    if (nbr->skin_built > 0.0) {
        tab_set_error("tab_nbr_export: the lists carry a skin of %g A (entries beyond rc); "
                      "build with skin = 0 for the reference's (ilist, jlist, n1)",
                      nbr->skin_built);
        return TAB_ESTATE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int n = nbr->n;
    TAB_TRY(nbr->row_ptr.ensure(sizeof(uint32_t) * (size_t)n));
    k_scatter_counts<<<nblocks(n, 256), 256, 0, st>>>(
        n, nbr->perm.as<int>(), nbr->counts.as<int>(), nbr->row_ptr.as<int>());
    TAB_LAUNCH_CHECK();
    TAB_TRY(tab_scan_exclusive_u32(nbr->row_ptr.as<uint32_t>(),
                                   nbr->row_ptr.as<uint32_t>(), n, nullptr,
                                   nbr->scan_tmp, st));
    k_export<<<nblocks(n, 128), 128, 0, st>>>(
        n, nbr->n_loc, nbr->perm.as<int>(), nbr->s0.as<int>(), nbr->counts.as<int>(),
        nbr->slice_ptr.as<uint32_t>(), nbr->col.as<uint32_t>(),
        nbr->ghost_owner.as<int>(), nbr->ghost_S.as<int>(),
        nbr->row_ptr.as<uint32_t>(), d_i, d_j, d_S);
    TAB_LAUNCH_CHECK();
    return TAB_OK;
}

// internal: make sure the reverse-pair index exists (lazy, once per build)
int tab_nbr_ensure_reverse(tab_nbr *nbr, cudaStream_t st) {
    if (nbr->has_rev) return TAB_OK;
    TAB_TRY(nbr->rev.ensure(sizeof(uint32_t) * 32 * (size_t)(nbr->ell_rows + 1)));
    k_build_reverse<<<nbr->n, 128, 0, st>>>(
        nbr->n, nbr->n_loc, nbr->counts.as<int>(), nbr->slice_ptr.as<uint32_t>(),
        nbr->col.as<uint32_t>(), nbr->ghost_owner.as<int>(), nbr->ghost_S.as<int>(),
        nbr->rev.as<uint32_t>());
    TAB_LAUNCH_CHECK();
    nbr->has_rev = true;
    return TAB_OK;
}

// ---------------------------------------------------------------------------
// halo packing (domain decomposition): dst[k, :] = src[idx[k], :] (+ shift)
// ---------------------------------------------------------------------------
__global__ void k_pack_rows(int m, int ncol, const double *__restrict__ src,
                            const long long *__restrict__ idx, double sx, double sy,
                            double sz, int shifted, double *__restrict__ dst) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m * ncol) return;
    const int k = t / ncol, c = t - k * ncol;
    double v = src[(size_t)idx[k] * ncol + c];
    if (shifted) v += c == 0 ? sx : (c == 1 ? sy : sz);
    dst[t] = v;
}

extern "C" int tab_pack_rows(const double *d_src, const int64_t *d_idx, int32_t m,
                             int32_t ncol, const double *h_shift, double *d_dst,
                             void *stream) {
    if (m <= 0) return TAB_OK;
    if (!d_src || !d_idx || !d_dst || ncol < 1 || (h_shift && ncol != 3)) {
        tab_set_error("tab_pack_rows: bad argument");
        return TAB_EINVAL;
    }
    k_pack_rows<<<nblocks((long long)m * ncol, 256), 256, 0, (cudaStream_t)stream>>>(
        m, ncol, d_src, (const long long *)d_idx, h_shift ? h_shift[0] : 0.0,
        h_shift ? h_shift[1] : 0.0, h_shift ? h_shift[2] : 0.0, h_shift ? 1 : 0, d_dst);
    TAB_LAUNCH_CHECK();
    return TAB_OK;
}

// ---------------------------------------------------------------------------
// small all-reduce over peer memory (see include/tab200.h)
// ---------------------------------------------------------------------------
__global__ void k_peer_put(int n, int n_peers, int slot, const double *__restrict__ src,
                           const unsigned long long *__restrict__ peer_ptrs) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * n_peers) return;
    const int p = t / n, q = t - p * n;
    double *dst = reinterpret_cast<double *>(peer_ptrs[p]);
    dst[(size_t)slot * n + q] = src[q];
}

// entries [0, n_sum) are summed in rank order, entries [n_sum, n) reduced with max
__global__ void k_sum_slots(int n_slots, int n, int n_sum, const double *__restrict__ slots,
                            double *__restrict__ out) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    double v = 0.0;
    if (q < n_sum) {
        for (int s = 0; s < n_slots; ++s) v += slots[(size_t)s * n + q];   // fixed order
    } else {
        v = slots[q];
        for (int s = 1; s < n_slots; ++s) v = fmax(v, slots[(size_t)s * n + q]);
    }
    out[q] = v;
}

__global__ void k_disp_to(const unsigned int *__restrict__ disp, double *__restrict__ out) {
    const float d2 = __uint_as_float(disp[0]);
    out[0] = d2 == d2 ? sqrt((double)d2) : (double)INFINITY;
}

extern "C" int tab_peer_put(const double *d_src, int32_t n, const uint64_t *d_peer_ptrs,
                            int32_t n_peers, int32_t slot, void *stream) {
    if (!d_src || !d_peer_ptrs || n <= 0 || n_peers <= 0 || slot < 0) {
        tab_set_error("tab_peer_put: bad argument");
        return TAB_EINVAL;
    }
    k_peer_put<<<nblocks((long long)n * n_peers, 128), 128, 0, (cudaStream_t)stream>>>(
        n, n_peers, slot, d_src, (const unsigned long long *)d_peer_ptrs);
    TAB_LAUNCH_CHECK();
    return TAB_OK;
}

extern "C" int tab_sum_slots(const double *d_slots, int32_t n_slots, int32_t n,
                             double *d_out, void *stream) {
    if (!d_slots || !d_out || n <= 0 || n_slots <= 0) {
        tab_set_error("tab_sum_slots: bad argument");
        return TAB_EINVAL;
    }
    k_sum_slots<<<nblocks(n, 128), 128, 0, (cudaStream_t)stream>>>(n_slots, n, n, d_slots, d_out);
    TAB_LAUNCH_CHECK();
    return TAB_OK;
}

extern "C" int tab_reduce_slots(const double *d_slots, int32_t n_slots, int32_t n,
                                int32_t n_sum, double *d_out, void *stream) {
    if (!d_slots || !d_out || n <= 0 || n_slots <= 0 || n_sum < 0 || n_sum > n) {
        tab_set_error("tab_reduce_slots: bad argument");
        return TAB_EINVAL;
    }
    k_sum_slots<<<nblocks(n, 128), 128, 0, (cudaStream_t)stream>>>(n_slots, n, n_sum, d_slots,
                                                                  d_out);
    TAB_LAUNCH_CHECK();
    return TAB_OK;
}

extern "C" int tab_nbr_displacement_device(tab_nbr *nbr, double *d_out, void *stream) {
    if (!nbr || !d_out) return TAB_EINVAL;
    if (!nbr->built) {
        tab_set_error("tab_nbr_displacement_device before tab_nbr_build");
        return TAB_ESTATE;
    }
    if (!(nbr->skin_built > 0.0)) {
        tab_set_error("tab_nbr_displacement_device: the lists carry no skin");
        return TAB_ESTATE;
    }
    k_disp_to<<<1, 1, 0, (cudaStream_t)stream>>>(nbr->disp.as<unsigned int>(), d_out);
    TAB_LAUNCH_CHECK();
    return TAB_OK;
}

#include "nbr_batch.cuh"
