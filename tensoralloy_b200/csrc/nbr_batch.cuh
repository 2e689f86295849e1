// nbr_batch.cuh -- neighbour lists of MANY small structures in one pass (included at
// the end of nbr.cu: it reuses k_slice_stats / k_pad_tail / the scan).
//
// The reference batches structures by padding every per-structure tensor to the batch
// maxima (BatchUniversalTransformer, transformer/universal.py:921-1388: `[B, N+1, 3]`
// positions, `[B, nij_max, ...]` maps) and still builds each neighbour list with ASE in
// Python.  Here a batch is ONE extended atom array and ONE sliced-ELL table: structure b
// owns the atom range [off_b, off_b + n_b) (caller order = sorted order) and a range of
// ghost records (its periodic images); every kernel downstream (EAM / ADP passes,
// symmetry functions, MLPs, force assembly, JVP) runs once over the whole batch.
//
// Structures of a batch are small (~100 atoms: configs 2 and 4), so no cell list: the
// candidates of an atom are all atoms and all ghost images of its own structure, tested
// by one warp per atom with ballot compaction (the order inside a row is the candidate
// order: owned atoms by index, then ghosts by (atom, shift)) -- deterministic.
// Membership is ASE's, decided like in the single-structure builders: fast test on the
// pre-shifted records, candidates within 1e-9 rc^2 of the cutoff re-decided with ASE's
// exact expression on the caller's positions.
#pragma once

struct BStruct {
    double h[9];        // lattice, rows = vectors
    double hinv[9];
    double lo[3], hi[3];   // fractional window that can hold images within rc of the cell
    int off, n;         // owned atoms [off, off + n)
    int nimg[3];        // image layers per direction (0 = not periodic)
    int nshift;         // (2 nimg0+1)(2 nimg1+1)(2 nimg2+1)
    int goff, ng;       // ghost records: extended indices [N + goff, N + goff + ng)
};

// wrapped position records of the owned atoms (periodic directions only)
__global__ void kb_wrap(int N, const double *__restrict__ pos, const int *__restrict__ types,
                        const BStruct *__restrict__ S, const int *__restrict__ struct_of,
                        Atom4 *__restrict__ atoms, uint8_t *__restrict__ types_ext,
                        int *__restrict__ s0, int *__restrict__ perm,
                        unsigned long long *__restrict__ stats) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const BStruct &b = S[struct_of[i]];
    double x = pos[3 * i], y = pos[3 * i + 1], z = pos[3 * i + 2];
    int sh[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        sh[k] = 0;
        if (b.nimg[k] > 0) {
            const double s = x * b.hinv[k] + y * b.hinv[3 + k] + z * b.hinv[6 + k];
            sh[k] = (int)floor(s);
        }
    }
    if (sh[0] | sh[1] | sh[2]) {
        x -= sh[0] * b.h[0] + sh[1] * b.h[3] + sh[2] * b.h[6];
        y -= sh[0] * b.h[1] + sh[1] * b.h[4] + sh[2] * b.h[7];
        z -= sh[0] * b.h[2] + sh[1] * b.h[5] + sh[2] * b.h[8];
    }
    Atom4 r;
    r.x = x;
    r.y = y;
    r.z = z;
    r.w = 0.0;
    atoms[i] = r;
    const int t = types ? types[i] : 0;
    types_ext[i] = (uint8_t)t;
    if (t > 0) atomicMax(&stats[3], (unsigned long long)t);
    s0[i] = tab_pack_shift(sh[0], sh[1], sh[2]);
    perm[i] = i;
}

// One block per structure.  Candidate c = (atom j, shift index q), q != 0 shift; a ghost
// record is kept when its fractional coordinates fall into the window [lo, hi] (a
// superset of the images within rc of any point of the cell).  FILL = false counts,
// FILL = true writes the records in candidate order (block-wide exclusive scan).
template <bool FILL>
__global__ void __launch_bounds__(256)
kb_ghosts(int N, BStruct *__restrict__ S, Atom4 *__restrict__ atoms,
          uint8_t *__restrict__ types_ext, int *__restrict__ ghost_owner,
          int *__restrict__ ghost_S, uint32_t *__restrict__ gcount) {
    __shared__ uint32_t warp_tot[8];
    __shared__ uint32_t base_s;
    const BStruct b = S[blockIdx.x];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) base_s = 0;
    __syncthreads();
    const long long ncand = (long long)b.n * b.nshift;
    const int w0 = 2 * b.nimg[0] + 1, w1 = 2 * b.nimg[1] + 1;
    for (long long c0 = 0; c0 < ncand; c0 += 256) {
        const long long c = c0 + threadIdx.x;
        bool keep = false;
        int j = 0, sa = 0, sb = 0, sc = 0;
        Atom4 a;
        if (c < ncand) {
            j = b.off + (int)(c / b.nshift);
            const int q = (int)(c % b.nshift);
            sa = q % w0 - b.nimg[0];
            sb = (q / w0) % w1 - b.nimg[1];
            sc = q / (w0 * w1) - b.nimg[2];
            if (sa | sb | sc) {
                a = atoms[j];
                keep = true;
                const int sh[3] = {sa, sb, sc};
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const double f = a.x * b.hinv[k] + a.y * b.hinv[3 + k] +
                                     a.z * b.hinv[6 + k] + (double)sh[k];
                    keep = keep && f >= b.lo[k] && f <= b.hi[k];
                }
            }
        }
        const uint32_t m = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) warp_tot[warp] = __popc(m);
        __syncthreads();
        uint32_t before = base_s;
        for (int w = 0; w < warp; ++w) before += warp_tot[w];
        if (FILL && keep) {
            const uint32_t g = (uint32_t)b.goff + before + __popc(m & ((1u << lane) - 1u));
            a.x += sa * b.h[0] + sb * b.h[3] + sc * b.h[6];
            a.y += sa * b.h[1] + sb * b.h[4] + sc * b.h[7];
            a.z += sa * b.h[2] + sb * b.h[5] + sc * b.h[8];
            a.w = 0.0;
            atoms[(size_t)N + g] = a;
            types_ext[(size_t)N + g] = types_ext[j];
            ghost_owner[g] = j;
            ghost_S[g] = tab_pack_shift(sa, sb, sc);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t t = 0;
            for (int w = 0; w < 8; ++w) t += warp_tot[w];
            base_s += t;
        }
        __syncthreads();
    }
    if (!FILL && threadIdx.x == 0) gcount[blockIdx.x] = base_s;
}

__global__ void kb_set_goff(int n_struct, BStruct *__restrict__ S,
                            const uint32_t *__restrict__ gstart,
                            const uint32_t *__restrict__ gcount) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_struct) return;
    S[s].goff = (int)gstart[s];
    S[s].ng = (int)gcount[s];
}

// ASE's membership expression on the caller's positions (no FMA contraction); i owned,
// j any extended index of the same structure.
__device__ __noinline__ bool exact_inside_b(const BStruct &b, double rc, int N,
                                            const double *__restrict__ pos,
                                            const int *__restrict__ s0,
                                            const int *__restrict__ ghost_owner,
                                            const int *__restrict__ ghost_S, int i, int j) {
    int owner = j, Sa = 0, Sb = 0, Sc = 0;
    if (j >= N) {
        owner = ghost_owner[j - N];
        tab_unpack_shift(ghost_S[j - N], Sa, Sb, Sc);
    }
    int ia, ib, ic, ja, jb, jc;
    tab_unpack_shift(s0[i], ia, ib, ic);
    tab_unpack_shift(s0[owner], ja, jb, jc);
    const double S0 = (double)(Sa - ja + ia), S1 = (double)(Sb - jb + ib),
                 S2 = (double)(Sc - jc + ic);
    double D[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const double sh = __dadd_rn(__dadd_rn(__dmul_rn(S0, b.h[k]), __dmul_rn(S1, b.h[3 + k])),
                                    __dmul_rn(S2, b.h[6 + k]));
        D[k] = __dadd_rn(__dsub_rn(pos[3 * owner + k], pos[3 * i + k]), sh);
    }
    const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(D[0], D[0]), __dmul_rn(D[1], D[1])),
                                __dmul_rn(D[2], D[2]));
    return __dsqrt_rn(d2) < rc;
}

// One warp per owned atom; the 32 lanes test 32 candidates at a time.  Same row layout
// and species ordering as k_nbr_warp (nbr.cu).
template <bool FILL>
__global__ void __launch_bounds__(128)
kb_nbr(int N, double rc, const BStruct *__restrict__ S, const int *__restrict__ struct_of,
       const double *__restrict__ pos, const int *__restrict__ s0,
       const int *__restrict__ ghost_owner, const int *__restrict__ ghost_S,
       const Atom4 *__restrict__ atoms, const uint8_t *__restrict__ types_ext, int n_types,
       int *__restrict__ counts, int *__restrict__ tcounts,
       const uint32_t *__restrict__ slice_w, const uint32_t *__restrict__ slice_ptr,
       uint32_t *__restrict__ col) {
    const int idx = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (idx >= N) return;
    const uint32_t lt = (1u << lane) - 1u;
    const BStruct &b = S[struct_of[idx]];
    const Atom4 me = atoms[idx];
    const double rc2 = rc * rc, tol = 1e-9 * rc2;
    uint32_t off[TAB_MAX_ELEMENTS];
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
#pragma unroll 1
    for (int part = 0; part < 2; ++part) {
        const int start = part ? N + b.goff : b.off;
        const int cnt = part ? b.ng : b.n;
        for (int k0 = 0; k0 < cnt; k0 += 32) {
            const int k = k0 + lane;
            bool in = false;
            int j = 0;
            if (k < cnt) {
                j = start + k;
                if (j != idx) {
                    const Atom4 a = atoms[j];
                    const double ddx = a.x - me.x, ddy = a.y - me.y, ddz = a.z - me.z;
                    const double d2 = ddx * ddx + ddy * ddy + ddz * ddz;
                    in = d2 < rc2;
                    if (fabs(d2 - rc2) <= tol)
                        in = exact_inside_b(b, rc, N, pos, s0, ghost_owner, ghost_S, idx, j);
                }
            }
            const uint32_t t = in ? types_ext[j] : 0u;
            if (n_types == 1) {
                const uint32_t m = __ballot_sync(0xffffffffu, in);
                if (FILL && in) base[(size_t)(off[0] + __popc(m & lt)) * 32u] = (uint32_t)j;
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
    if (!FILL) {
        uint32_t total = 0;
        for (int t = 0; t < n_types; ++t) total += off[t];
        if (lane == 0) {
            counts[idx] = (int)total;
            for (int t = 0; t < n_types; ++t) tcounts[(size_t)idx * n_types + t] = (int)off[t];
        }
    } else {
        const uint32_t w = slice_w[idx >> 5];
        const uint32_t have = (uint32_t)counts[idx];
        for (uint32_t k = have + lane; k < w; k += 32) base[(size_t)k * 32u] = TAB_COL_PAD;   // batch rows are read through the counts only
    }
}

extern "C" int tab_nbr_build_batch(tab_nbr *nbr, int32_t n_struct, const int32_t *h_offsets,
                                   const double *d_pos, const int32_t *d_types,
                                   const double *h_cells, const int32_t *h_pbc, double rc,
                                   void *stream) {
    if (!nbr || n_struct <= 0 || !h_offsets || !d_pos || !h_cells || !h_pbc || !(rc > 0)) {
        tab_set_error("tab_nbr_build_batch: bad argument");
        return TAB_EINVAL;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int N = h_offsets[n_struct];
    if (h_offsets[0] != 0 || N <= 0) {
        tab_set_error("tab_nbr_build_batch: offsets must start at 0 and end at the atom count");
        return TAB_EINVAL;
    }
    nbr->built = false;
    nbr->skin_built = 0.0;      // batch handles are rebuilt per batch: no skin
    nbr->rc_model = rc;
    nbr->ls_L = 0;
    nbr->col_padded = false;
    nbr->rec16_valid = false;
    nbr->pcache_valid = false;
    nbr->has_rev = false;
    nbr->has_row_ptr = false;
    nbr->wcap_hint = 0;
    std::vector<BStruct> hs(n_struct);
    std::vector<int> h_struct_of((size_t)N);
    const double rpad = rc * (1.0 + 1e-6);
    for (int s = 0; s < n_struct; ++s) {
        BStruct &b = hs[s];
        memcpy(b.h, h_cells + 9 * (size_t)s, 9 * sizeof(double));
        double det;
        invert3(b.h, b.hinv, &det);
        if (!(fabs(det) > 1e-12)) {
            tab_set_error("structure %d: cell is singular (det=%g); non-periodic structures "
                          "must be given a bounding cell by the caller", s, det);
            return TAB_EINVAL;
        }
        b.off = h_offsets[s];
        b.n = h_offsets[s + 1] - h_offsets[s];
        if (b.n <= 0) {
            tab_set_error("structure %d is empty", s);
            return TAB_EINVAL;
        }
        long long nshift = 1;
        for (int k = 0; k < 3; ++k) {
            const double nrm = sqrt(b.hinv[k] * b.hinv[k] + b.hinv[3 + k] * b.hinv[3 + k] +
                                    b.hinv[6 + k] * b.hinv[6 + k]);
            const double fd = 1.0 / nrm;          // face distance along direction k
            b.nimg[k] = h_pbc[3 * s + k] ? (int)ceil(rpad / fd) : 0;
            // (non-periodic directions carry no images and no constraint: atoms may lie
            // anywhere relative to the caller's bounding cell)
            b.lo[k] = b.nimg[k] ? -rpad / fd - 1e-9 : -1e300;
            b.hi[k] = b.nimg[k] ? 1.0 + rpad / fd + 1e-9 : 1e300;
            nshift *= 2 * b.nimg[k] + 1;
            if (b.nimg[k] > 500) {
                tab_set_error("structure %d: cutoff spans %d images of the cell", s, b.nimg[k]);
                return TAB_EUNSUPPORTED;
            }
        }
        if (nshift * b.n > 0x3fffffffLL) {
            tab_set_error("structure %d: too many image candidates", s);
            return TAB_EUNSUPPORTED;
        }
        b.nshift = (int)nshift;
        b.goff = 0;
        b.ng = 0;
        for (int i = b.off; i < b.off + b.n; ++i) h_struct_of[i] = s;
    }
    nbr->n = N;
    nbr->n_halo = 0;
    nbr->n_loc = N;
    nbr->n_slices = (N + 31) / 32;
    nbr->n_struct = n_struct;
    nbr->h_struct_off.assign(h_offsets, h_offsets + n_struct + 1);
    nbr->blk_T = 0;
    Grid &g = nbr->grid;
    memset(&g, 0, sizeof(g));
    memcpy(g.h, hs[0].h, sizeof(g.h));
    memcpy(g.hinv, hs[0].hinv, sizeof(g.hinv));
    g.rc = rc;
    g.rc2 = rc * rc;

    TAB_TRY(nbr->bstructs.ensure(sizeof(BStruct) * (size_t)n_struct));
    TAB_TRY(nbr->struct_off.ensure(sizeof(int) * (size_t)(n_struct + 1)));
    TAB_TRY(nbr->struct_of.ensure(sizeof(int) * (size_t)N));
    TAB_TRY(nbr->s0.ensure(sizeof(int) * (size_t)N));
    TAB_TRY(nbr->perm.ensure(sizeof(int) * (size_t)N));
    TAB_TRY(nbr->counts.ensure(sizeof(int) * (size_t)N));
    TAB_TRY(nbr->gcount.ensure(sizeof(uint32_t) * (size_t)n_struct));
    TAB_TRY(nbr->gstart.ensure(sizeof(uint32_t) * (size_t)n_struct));
    TAB_TRY(nbr->slice_w.ensure(sizeof(uint32_t) * nbr->n_slices));
    TAB_TRY(nbr->slice_ptr.ensure(sizeof(uint32_t) * nbr->n_slices));
    TAB_TRY(nbr->stats.ensure(5 * sizeof(unsigned long long)));
    // owned records first; the ghost part is sized after the count (room for 9 images per
    // atom up front so that the usual batch does not have to regrow and copy)
    TAB_TRY(nbr->atoms.ensure(sizeof(Atom4) * (size_t)N * 10));
    TAB_TRY(nbr->types_ext.ensure((size_t)N * 10 + 16));
    unsigned long long *d_stats = nbr->stats.as<unsigned long long>();
    BStruct *d_S = nbr->bstructs.as<BStruct>();
    TAB_CUDA(cudaMemcpyAsync(d_S, hs.data(), sizeof(BStruct) * (size_t)n_struct,
                             cudaMemcpyHostToDevice, st));
    TAB_CUDA(cudaMemcpyAsync(nbr->struct_off.p, h_offsets, sizeof(int) * (size_t)(n_struct + 1),
                             cudaMemcpyHostToDevice, st));
    TAB_CUDA(cudaMemcpyAsync(nbr->struct_of.p, h_struct_of.data(), sizeof(int) * (size_t)N,
                             cudaMemcpyHostToDevice, st));
    TAB_CUDA(cudaMemsetAsync(d_stats, 0, 5 * sizeof(unsigned long long), st));
    kb_wrap<<<nblocks(N, 256), 256, 0, st>>>(N, d_pos, d_types, d_S, nbr->struct_of.as<int>(),
                                             nbr->atoms.as<Atom4>(),
                                             nbr->types_ext.as<uint8_t>(), nbr->s0.as<int>(),
                                             nbr->perm.as<int>(), d_stats);
    TAB_LAUNCH_CHECK();
    kb_ghosts<false><<<n_struct, 256, 0, st>>>(N, d_S, nbr->atoms.as<Atom4>(), nullptr, nullptr,
                                               nullptr, nbr->gcount.as<uint32_t>());
    TAB_LAUNCH_CHECK();
    TAB_TRY(tab_scan_exclusive_u32(nbr->gcount.as<uint32_t>(), nbr->gstart.as<uint32_t>(),
                                   n_struct, d_stats + 2, nbr->scan_tmp, st));
    kb_set_goff<<<nblocks(n_struct, 128), 128, 0, st>>>(n_struct, d_S,
                                                        nbr->gstart.as<uint32_t>(),
                                                        nbr->gcount.as<uint32_t>());
    TAB_LAUNCH_CHECK();
    unsigned long long two[2] = {0, 0};      // [n_ghost, max type]
    TAB_CUDA(cudaMemcpyAsync(two, d_stats + 2, sizeof(two), cudaMemcpyDeviceToHost, st));
    // the pageable host vectors above must outlive their copies: this sync covers them
    TAB_CUDA(cudaStreamSynchronize(st));
    nbr->n_types = (int)two[1] + 1;
    if (nbr->n_types > TAB_MAX_ELEMENTS) {
        tab_set_error("element index %d exceeds the supported maximum", nbr->n_types - 1);
        return TAB_EINVAL;
    }
    if ((unsigned long long)N + two[0] > TAB_COL_IDX_MASK) {
        tab_set_error("owned + ghost atoms exceed the 28-bit index space");
        return TAB_EUNSUPPORTED;
    }
    nbr->n_ghost = (int)two[0];
    nbr->n_ext = N + nbr->n_ghost;
    {   // grow the record arrays, keeping the owned part
        DevBuf old_atoms = nbr->atoms, old_types = nbr->types_ext;
        if (old_atoms.cap < sizeof(Atom4) * (size_t)nbr->n_ext ||
            old_types.cap < (size_t)nbr->n_ext + 16) {
            nbr->atoms = DevBuf();
            nbr->types_ext = DevBuf();
            TAB_TRY(nbr->atoms.ensure(sizeof(Atom4) * (size_t)nbr->n_ext));
            TAB_TRY(nbr->types_ext.ensure((size_t)nbr->n_ext + 16));
            TAB_CUDA(cudaMemcpyAsync(nbr->atoms.p, old_atoms.p, sizeof(Atom4) * (size_t)N,
                                     cudaMemcpyDeviceToDevice, st));
            TAB_CUDA(cudaMemcpyAsync(nbr->types_ext.p, old_types.p, (size_t)N,
                                     cudaMemcpyDeviceToDevice, st));
            TAB_CUDA(cudaStreamSynchronize(st));
            old_atoms.release();
            old_types.release();
        }
    }
    TAB_TRY(nbr->ghost_owner.ensure(sizeof(int) * (size_t)(nbr->n_ghost + 1)));
    TAB_TRY(nbr->ghost_S.ensure(sizeof(int) * (size_t)(nbr->n_ghost + 1)));
    if (nbr->n_ghost > 0) {
        kb_ghosts<true><<<n_struct, 256, 0, st>>>(
            N, d_S, nbr->atoms.as<Atom4>(), nbr->types_ext.as<uint8_t>(),
            nbr->ghost_owner.as<int>(), nbr->ghost_S.as<int>(), nbr->gcount.as<uint32_t>());
        TAB_LAUNCH_CHECK();
    }
    TAB_TRY(nbr->tcounts.ensure(sizeof(int) * (size_t)N * nbr->n_types));
    kb_nbr<false><<<nblocks((long long)N * 32, 128), 128, 0, st>>>(
        N, rc, d_S, nbr->struct_of.as<int>(), d_pos, nbr->s0.as<int>(),
        nbr->ghost_owner.as<int>(), nbr->ghost_S.as<int>(), nbr->atoms.as<Atom4>(),
        nbr->types_ext.as<uint8_t>(), nbr->n_types, nbr->counts.as<int>(),
        nbr->tcounts.as<int>(), nullptr, nullptr, nullptr);
    TAB_LAUNCH_CHECK();
    TAB_CUDA(cudaMemsetAsync(d_stats, 0, 3 * sizeof(unsigned long long), st));
    k_slice_stats<<<min(nblocks(nbr->n_slices * 32, 256), 1184), 256, 0, st>>>(
        N, nbr->counts.as<int>(), nbr->slice_w.as<uint32_t>(), d_stats);
    TAB_LAUNCH_CHECK();
    TAB_TRY(tab_scan_exclusive_u32(nbr->slice_w.as<uint32_t>(), nbr->slice_ptr.as<uint32_t>(),
                                   nbr->n_slices, d_stats + 2, nbr->scan_tmp, st));
    unsigned long long h_stats[3];
    TAB_CUDA(cudaMemcpyAsync(h_stats, d_stats, sizeof(h_stats), cudaMemcpyDeviceToHost, st));
    TAB_CUDA(cudaStreamSynchronize(st));
    nbr->nij = (long long)h_stats[0];
    nbr->nnl_max = (int)h_stats[1];
    nbr->ell_rows = (long long)h_stats[2];
    if (nbr->ell_rows > 0xffffffffLL) {
        tab_set_error("neighbour table too large (%lld rows)", nbr->ell_rows);
        return TAB_EUNSUPPORTED;
    }
    TAB_TRY(nbr->col.ensure(sizeof(uint32_t) * 32 * (size_t)(nbr->ell_rows + 1)));
    kb_nbr<true><<<nblocks((long long)N * 32, 128), 128, 0, st>>>(
        N, rc, d_S, nbr->struct_of.as<int>(), d_pos, nbr->s0.as<int>(),
        nbr->ghost_owner.as<int>(), nbr->ghost_S.as<int>(), nbr->atoms.as<Atom4>(),
        nbr->types_ext.as<uint8_t>(), nbr->n_types, nbr->counts.as<int>(),
        nbr->tcounts.as<int>(), nbr->slice_w.as<uint32_t>(), nbr->slice_ptr.as<uint32_t>(),
        nbr->col.as<uint32_t>());
    TAB_LAUNCH_CHECK();
    if (N & 31) {
        k_pad_tail<<<1, 32, 0, st>>>(N, nbr->slice_w.as<uint32_t>(),
                                     nbr->slice_ptr.as<uint32_t>(), nbr->col.as<uint32_t>(),
                                     TAB_COL_PAD);
        TAB_LAUNCH_CHECK();
    }
    nbr->built = true;
    return TAB_OK;
}

// per-structure atom counts / offsets of a batch handle (host copies)
extern "C" int tab_nbr_batch_size(const tab_nbr *nbr) { return nbr ? nbr->n_struct : 0; }
