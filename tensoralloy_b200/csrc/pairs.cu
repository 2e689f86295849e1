// pairs.cu -- pair-level training operators of the EAM / ADP family.
//
// The reference trains EAM / ADP models (empirical parameters and 'nn' functions) with
// TF second-order autograd over the padded pair tensors: loss(E, F, stress) ->
// d loss / d parameters (nn/basic.py:446-631, nn/opt.py:132-157, nn/eam/eam.py:495-570).
// Here the energy of a batch is a function of the directed pair vectors D_p = R_j - R_i
// (one variable per list entry); what touches the neighbour lists is LINEAR in the
// per-pair gradient g_p = dE/dD_p and lives in three kernels:
//   tab_pairs_export   (i, j, D_p) of every list entry, rows sorted by i
//   tab_pair_forces    F_i = sum_{p in row i} g_p - sum_{p -> i} g_p ,
//                      W   = sum_p sym(g_p (x) D_p)          (per structure)
//   tab_pair_jvp       its transpose:  t_p = u_i - u_j + sym(A) D_p
// ("PyTorch custom ops where tensors cross into training"): torch evaluates the scalar
// functions rho(r), phi(r), F(rho), u(r), w(r) of the pair table and differentiates them;
// the force / virial assembly and its adjoint never leave libtab200.  No atomics: the
// "-> i" sum runs over the reverse-pair index of row i, fixed order.
#include "tab_internal.h"

static inline int nblocks_p(long long n, int t) { return (int)((n + t - 1) / t); }

__global__ void k_counts_caller(int n, const int *__restrict__ perm,
                                const int *__restrict__ counts, uint32_t *__restrict__ out) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < n) out[perm[idx]] = (uint32_t)counts[idx];
}

// export position of the first entry of every row (caller order), cached per build
static int ensure_row_ptr(tab_nbr *nbr, cudaStream_t st) {
    if (nbr->has_row_ptr) return TAB_OK;
    const int n = nbr->n;
    TAB_TRY(nbr->pair_row_ptr.ensure(sizeof(uint32_t) * (size_t)(n + 1)));
    k_counts_caller<<<nblocks_p(n, 256), 256, 0, st>>>(n, nbr->perm.as<int>(),
                                                       nbr->counts.as<int>(),
                                                       nbr->pair_row_ptr.as<uint32_t>());
    TAB_LAUNCH_CHECK();
    TAB_TRY(tab_scan_exclusive_u32(nbr->pair_row_ptr.as<uint32_t>(),
                                   nbr->pair_row_ptr.as<uint32_t>(), n, nullptr,
                                   nbr->scan_tmp, st));
    nbr->has_row_ptr = true;
    return TAB_OK;
}

// one warp per atom row
__global__ void __launch_bounds__(128)
k_pairs_export(int n, int n_loc, const int *__restrict__ perm, const int *__restrict__ counts,
               const uint32_t *__restrict__ slice_ptr, const uint32_t *__restrict__ col,
               const int *__restrict__ ghost_owner, const Atom4 *__restrict__ atoms,
               const uint32_t *__restrict__ row_ptr, int *__restrict__ out_i,
               int *__restrict__ out_j, double *__restrict__ out_D) {
    const int idx = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (idx >= n) return;
    const int oi = perm[idx];
    const Atom4 me = atoms[idx];
    const uint32_t *base = col + ((size_t)slice_ptr[idx >> 5] * 32u + (idx & 31));
    const size_t o0 = row_ptr[oi];
    const int cnt = counts[idx];
    for (int k = lane; k < cnt; k += 32) {
        const int j = (int)(base[(size_t)k * 32u] & TAB_COL_IDX_MASK);
        const int owner = j >= n_loc ? ghost_owner[j - n_loc] : j;
        const Atom4 a = atoms[j];
        const size_t o = o0 + k;
        if (out_i) out_i[o] = oi;
        if (out_j) out_j[o] = perm[owner];
        out_D[3 * o + 0] = a.x - me.x;
        out_D[3 * o + 1] = a.y - me.y;
        out_D[3 * o + 2] = a.z - me.z;
    }
}

__device__ __forceinline__ double warp_sum_p(double v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

// one warp per atom row: forces and the atom's virial row (partial[idx*8 + 1..6])
__global__ void __launch_bounds__(128)
k_pair_forces(int n, int n_loc, const int *__restrict__ perm, const int *__restrict__ counts,
              const uint32_t *__restrict__ slice_ptr, const uint32_t *__restrict__ col,
              const uint32_t *__restrict__ rev, const int *__restrict__ ghost_owner,
              const Atom4 *__restrict__ atoms, const uint32_t *__restrict__ row_ptr,
              const double *__restrict__ g, double *__restrict__ forces,
              double *__restrict__ partial) {
    const int idx = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (idx >= n) return;
    const int oi = perm[idx];
    const Atom4 me = atoms[idx];
    const size_t ebase = (size_t)slice_ptr[idx >> 5] * 32u + (idx & 31);
    const size_t o0 = row_ptr[oi];
    const int cnt = counts[idx];
    double fx = 0, fy = 0, fz = 0, v[6] = {0, 0, 0, 0, 0, 0};
    for (int k = lane; k < cnt; k += 32) {
        const size_t ent = ebase + (size_t)k * 32u;
        const int j = (int)(col[ent] & TAB_COL_IDX_MASK);
        const Atom4 a = atoms[j];
        const double dx = a.x - me.x, dy = a.y - me.y, dz = a.z - me.z;
        const double *gp = g + 3 * (o0 + k);
        const double gx = gp[0], gy = gp[1], gz = gp[2];
        fx += gx;
        fy += gy;
        fz += gz;
        v[0] += gx * dx;
        v[1] += gy * dy;
        v[2] += gz * dz;
        v[3] += 0.5 * (gy * dz + gz * dy);
        v[4] += 0.5 * (gx * dz + gz * dx);
        v[5] += 0.5 * (gx * dy + gy * dx);
        const uint32_t q = rev[ent];
        if (q != 0xFFFFFFFFu) {
            const int owner = j >= n_loc ? ghost_owner[j - n_loc] : j;
            const double *gr = g + 3 * ((size_t)row_ptr[perm[owner]] + q);
            fx -= gr[0];
            fy -= gr[1];
            fz -= gr[2];
        }
    }
    fx = warp_sum_p(fx);
    fy = warp_sum_p(fy);
    fz = warp_sum_p(fz);
#pragma unroll
    for (int q = 0; q < 6; ++q) v[q] = warp_sum_p(v[q]);
    if (lane == 0) {
        if (forces) {
            forces[3 * (size_t)oi + 0] = fx;
            forces[3 * (size_t)oi + 1] = fy;
            forces[3 * (size_t)oi + 2] = fz;
        }
#pragma unroll
        for (int q = 0; q < 6; ++q) partial[(size_t)idx * 8 + 1 + q] = v[q];
    }
}

// virial of structure s (or of the whole single structure) from the per-atom rows
__global__ void __launch_bounds__(256)
k_pair_virial(int n, const int *__restrict__ struct_off, const double *__restrict__ partial,
              double *__restrict__ virial) {
    __shared__ double sm[256][6];
    const int s = blockIdx.x;
    const int lo = struct_off ? struct_off[s] : 0, hi = struct_off ? struct_off[s + 1] : n;
    double a[6] = {0, 0, 0, 0, 0, 0};
    for (int i = lo + threadIdx.x; i < hi; i += 256)
#pragma unroll
        for (int q = 0; q < 6; ++q) a[q] += partial[(size_t)i * 8 + 1 + q];
#pragma unroll
    for (int q = 0; q < 6; ++q) sm[threadIdx.x][q] = a[q];
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if (threadIdx.x < w)
#pragma unroll
            for (int q = 0; q < 6; ++q) sm[threadIdx.x][q] += sm[threadIdx.x + w][q];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        double *v = virial + 9 * (size_t)s;
        const double xx = sm[0][0], yy = sm[0][1], zz = sm[0][2], yz = sm[0][3],
                     xz = sm[0][4], xy = sm[0][5];
        v[0] = xx; v[1] = xy; v[2] = xz;
        v[3] = xy; v[4] = yy; v[5] = yz;
        v[6] = xz; v[7] = yz; v[8] = zz;
    }
}

// t_p = u_i - u_j + sym(A) D_p   (A of the atom's structure)
__global__ void __launch_bounds__(128)
k_pair_jvp(int n, int n_loc, const int *__restrict__ perm, const int *__restrict__ counts,
           const uint32_t *__restrict__ slice_ptr, const uint32_t *__restrict__ col,
           const int *__restrict__ ghost_owner, const Atom4 *__restrict__ atoms,
           const uint32_t *__restrict__ row_ptr, const int *__restrict__ struct_of,
           const double *__restrict__ u, const double *__restrict__ A,
           double *__restrict__ t) {
    const int idx = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (idx >= n) return;
    const int oi = perm[idx];
    const Atom4 me = atoms[idx];
    const double *As = A + (struct_of ? 9 * (size_t)struct_of[idx] : 0);
    const double axx = As[0], ayy = As[4], azz = As[8];
    const double axy = 0.5 * (As[1] + As[3]), axz = 0.5 * (As[2] + As[6]),
                 ayz = 0.5 * (As[5] + As[7]);
    const double ux = u[3 * (size_t)oi], uy = u[3 * (size_t)oi + 1], uz = u[3 * (size_t)oi + 2];
    const uint32_t *base = col + ((size_t)slice_ptr[idx >> 5] * 32u + (idx & 31));
    const size_t o0 = row_ptr[oi];
    const int cnt = counts[idx];
    for (int k = lane; k < cnt; k += 32) {
        const int j = (int)(base[(size_t)k * 32u] & TAB_COL_IDX_MASK);
        const int owner = j >= n_loc ? ghost_owner[j - n_loc] : j;
        const int oj = perm[owner];
        const Atom4 a = atoms[j];
        const double dx = a.x - me.x, dy = a.y - me.y, dz = a.z - me.z;
        double *tp = t + 3 * (o0 + k);
        tp[0] = ux - u[3 * (size_t)oj + 0] + axx * dx + axy * dy + axz * dz;
        tp[1] = uy - u[3 * (size_t)oj + 1] + axy * dx + ayy * dy + ayz * dz;
        tp[2] = uz - u[3 * (size_t)oj + 2] + axz * dx + ayz * dy + azz * dz;
    }
}

static int pair_checks(tab_nbr *nbr, const char *who) {
    if (!nbr) {
        tab_set_error("%s: null handle", who);
        return TAB_EINVAL;
    }
    if (!nbr->built) {
        tab_set_error("%s before tab_nbr_build", who);
        return TAB_ESTATE;
    }
    if (nbr->n_halo > 0) {
        tab_set_error("%s: halo atoms (domain decomposition) are not supported", who);
        return TAB_EUNSUPPORTED;
    }
    if (nbr->skin_built > 0.0) {
        tab_set_error("%s: the lists carry a skin (entries beyond rc); build with skin = 0", who);
        return TAB_ESTATE;
    }
    return TAB_OK;
}

extern "C" int tab_pairs_export(tab_nbr *nbr, int32_t *d_i, int32_t *d_j, double *d_D,
                                void *stream) {
    TAB_TRY(pair_checks(nbr, "tab_pairs_export"));
    if (!d_D) return TAB_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    TAB_TRY(ensure_row_ptr(nbr, st));
    k_pairs_export<<<nblocks_p((long long)nbr->n * 32, 128), 128, 0, st>>>(
        nbr->n, nbr->n_loc, nbr->perm.as<int>(), nbr->counts.as<int>(),
        nbr->slice_ptr.as<uint32_t>(), nbr->col.as<uint32_t>(), nbr->ghost_owner.as<int>(),
        nbr->atoms.as<Atom4>(), nbr->pair_row_ptr.as<uint32_t>(), d_i, d_j, d_D);
    TAB_LAUNCH_CHECK();
    return TAB_OK;
}

extern "C" int tab_pair_forces(tab_nbr *nbr, const double *d_g, double *d_forces,
                               double *d_virial, void *stream) {
    TAB_TRY(pair_checks(nbr, "tab_pair_forces"));
    if (!d_g) return TAB_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    TAB_TRY(ensure_row_ptr(nbr, st));
    TAB_TRY(tab_nbr_ensure_reverse(nbr, st));
    TAB_TRY(nbr->partial.ensure(sizeof(double) * 8 * (size_t)nbr->n));
    k_pair_forces<<<nblocks_p((long long)nbr->n * 32, 128), 128, 0, st>>>(
        nbr->n, nbr->n_loc, nbr->perm.as<int>(), nbr->counts.as<int>(),
        nbr->slice_ptr.as<uint32_t>(), nbr->col.as<uint32_t>(), nbr->rev.as<uint32_t>(),
        nbr->ghost_owner.as<int>(), nbr->atoms.as<Atom4>(), nbr->pair_row_ptr.as<uint32_t>(),
        d_g, d_forces, nbr->partial.as<double>());
    TAB_LAUNCH_CHECK();
    if (d_virial) {
        const bool batch = nbr->n_struct > 0;
        k_pair_virial<<<batch ? nbr->n_struct : 1, 256, 0, st>>>(
            nbr->n, batch ? nbr->struct_off.as<int>() : nullptr, nbr->partial.as<double>(),
            d_virial);
        TAB_LAUNCH_CHECK();
    }
    return TAB_OK;
}

extern "C" int tab_pair_jvp(tab_nbr *nbr, const double *d_u, const double *d_A, double *d_t,
                            void *stream) {
    TAB_TRY(pair_checks(nbr, "tab_pair_jvp"));
    if (!d_u || !d_A || !d_t) return TAB_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    TAB_TRY(ensure_row_ptr(nbr, st));
    k_pair_jvp<<<nblocks_p((long long)nbr->n * 32, 128), 128, 0, st>>>(
        nbr->n, nbr->n_loc, nbr->perm.as<int>(), nbr->counts.as<int>(),
        nbr->slice_ptr.as<uint32_t>(), nbr->col.as<uint32_t>(), nbr->ghost_owner.as<int>(),
        nbr->atoms.as<Atom4>(), nbr->pair_row_ptr.as<uint32_t>(),
        nbr->n_struct > 0 ? nbr->struct_of.as<int>() : nullptr, d_u, d_A, d_t);
    TAB_LAUNCH_CHECK();
    return TAB_OK;
}
