// eam.cu -- EAM / Finnis-Sinclair energy, forces and virial on the ELL lists.
//
// Replaces, for the reference:
//   transformer/universal.py:448-474,583-620  gather, D = Rj-Ri+S.h, r = sqrt(D.D+eps)
//   nn/eam/alloy.py:128-196, fs.py:146-203    rho_i = sum_j rho(r_ij)
//   nn/eam/eam.py:401-449                     F(rho_i)
//   nn/eam/eam.py:300-362                     0.5 * sum_j phi(r_ij)
//   nn/eam/eam.py:265-298                     E = sum_i E_i
//   nn/basic.py:276-331                       F = -dE/dR, virial = sum_p dE/dD_p (x) D_p
// TF autograd is replaced by the analytic derivative.  Lists are FULL (directed),
// so the force on atom i is assembled from row i alone:
//   F_i = sum_j [F'_i rho'_{ij}(r) + F'_j rho'_{ji}(r) + phi'_{ij}(r)] D_ij / r
// -- no scatter to j, no atomics, deterministic summation order.
//
// Two passes, thread per owned atom, lanes of a warp = the 32 atoms of one ELL
// slice (coalesced index loads), one 32-byte Atom4 gather per pair:
//   pass 1: rho_i, F(rho_i), F'(rho_i)
//   spread: Atom4.w <- F' for owned atoms and their ghost images
//   pass 2: forces, per-atom energy, block-reduced energy and virial
#include <stdlib.h>

#include "potentials.cuh"

struct EamDev {
    int kind;
    int n_el;
    const tab_fn *rho;     // [n_el*n_el] centre a, neighbour b
    const tab_fn *phi;     // [n_el*n_el]
    const tab_fn *embed;   // [n_el]
    const double *pool;    // spline coefficient pool (TAB_FN_SPLINE) or NULL
};

// single-element zjw04: rho and the B term of phi share one exponential
struct Zhou1 {
    double fe, beta, lamda, re, A, alpha, kappa, B;   // re holds 1/r_eq
    // float64 fast path: folded terms (potentials.cuh, ZTerm)
    ZTerm t_rho;        // fe exp(-beta (x-1)) / (1 + (x-lamda)^20)
    ZTerm t_a;          // A  exp(-alpha (x-1)) / (1 + (x-kappa)^20)
    ZTerm t_b;          // B  exp(-beta (x-1)) / (1 + (x-lamda)^20)
    double fe_over_B;
};

static void zterm_fold(ZTerm &t, double a, double b, double c, double re) {
    t.nb = -b;
    t.c = b + log(a);
    t.kappa = c;
    t.c20 = 20.0 / (re * re);
    t.k20 = -20.0 * c / re;
    t.nb_re = -b / re;
}

struct ZPair;
static ZPair *eamz_new_pair(const Zhou1 &z);
static void eamz_free_pair(ZPair *p);

struct tab_model {
    int family = 0;        // 0 = EAM
    int kind = 0;
    int n_el = 0;
    bool zhou1 = false;    // single element, all-zjw04: shared-exponential fast path
    bool has_mlp_fn = false;   // some function is an 'nn' MLP (no analytic Hessian)
    bool no_hessian = false;   // some function kind has no second-derivative evaluator
    Zhou1 z1;              // fe, beta, lamda, 1/re, A, alpha, kappa, B + folded terms
    bool z1_folded = false;   // prefactors positive: the folded float64 terms are usable
    struct ZPair *zp = nullptr;   // the same, in the form of the lane-split kernels (eam_fast.cuh)
    DevBuf tables;         // tab_fn [2*n_el*n_el + n_el (+ 2*n_el*n_el)]
    DevBuf pool;           // spline coefficients
    tab_fn embed0;         // host copy (fast path epilogue parameters)
};

#ifndef EAM_T
#define EAM_T 128        // threads per block of the pair kernels (sweep: profiles/r01d)
#endif
#ifndef EAM_MINB
#define EAM_MINB 4       // __launch_bounds__ min blocks per SM (register cap)
#endif
#ifndef EAM_FORCE_UNROLL
#define EAM_FORCE_UNROLL 1   // pairs per iteration of the force loop (ILP vs registers)
#endif
constexpr int kForceUnroll = EAM_FORCE_UNROLL;
// exp2 table size (log2) of the folded float64 zjw04 terms, per pass (potentials.cuh zexp)
#ifndef ZEXP_RHO
#define ZEXP_RHO 0
#endif
#ifndef ZEXP_FORCE
#define ZEXP_FORCE 6
#endif
#define ADP_MAX_EL 3     // ADP keeps n_el x 9 moment accumulators in registers

// Block -> atoms of the pair kernels.  Single structure: block b owns atoms
// [b T, (b+1) T).  Batch handles (tab_nbr_build_batch): blocks are cut at structure
// boundaries (blk_first[b] .. blk_first[b+1]) so that a block's partial sums belong to
// one structure.
__device__ __forceinline__ bool block_atom(int n, const int *__restrict__ blk_first,
                                           int &idx) {
    if (blk_first) {
        idx = blk_first[blockIdx.x] + (int)threadIdx.x;
        return idx < blk_first[blockIdx.x + 1];
    }
    idx = blockIdx.x * blockDim.x + threadIdx.x;
    return idx < n;
}

template <typename Real>
__device__ __forceinline__ void pair_r(const Atom4 &me, const Atom4 &a, Real &dx,
                                       Real &dy, Real &dz, Real &r, Real &rinv) {
    // geometry always in float64 (positions are float64), then the working type
    const double ddx = a.x - me.x, ddy = a.y - me.y, ddz = a.z - me.z;
    dx = (Real)ddx;
    dy = (Real)ddy;
    dz = (Real)ddz;
    // r = sqrt(D.D + eps)  (universal.py:470-473), with 1/r for the force
    const Real r2 = fma(dx, dx, fma(dy, dy, fma(dz, dz, Math<Real>::eps())));
    rinv = Math<Real>::rsqrt_(r2);
    r = r2 * rinv;
}

// Software-pipelined traversal of one ELL row: the index two entries ahead and
// the Atom4 record one entry ahead are in flight while entry k is evaluated
// (the gather latency was the top stall of the first version, profiles/r01a).
struct RowIter {
    const uint32_t *cp;
    const Atom4 *atoms;
    int cnt, k;
    uint32_t c_cur, c_nx, c_nx2;
    Atom4 a_nx;
    __device__ __forceinline__ RowIter(const uint32_t *cp_, const Atom4 *atoms_, int cnt_)
        : cp(cp_), atoms(atoms_), cnt(cnt_), k(0) {
        c_nx = cnt > 0 ? cp[0] : 0u;
        c_nx2 = cnt > 1 ? cp[32] : 0u;
        a_nx = atoms[c_nx & TAB_COL_IDX_MASK];
    }
    __device__ __forceinline__ bool next(Atom4 &a, uint32_t &c) {
        if (k >= cnt) return false;
        a = a_nx;
        c = c_nx;
        c_nx = c_nx2;
        ++k;
        if (k < cnt) a_nx = atoms[c_nx & TAB_COL_IDX_MASK];
        if (k + 1 < cnt) c_nx2 = cp[(size_t)(k + 1) * 32u];
        return true;
    }
    // entry index of the element returned by the last next()
    __device__ __forceinline__ int pos() const { return k - 1; }
};

// Per-pair cache between the two passes of the single-element zjw04 fast path:
// pass 1 has to evaluate g(r) = exp(-beta (x-1)) / (1 + (x-lamda)^20) for rho anyway;
// it also stores (g, dg/dr) per directed pair (ELL layout, 16 B, coalesced,
// streaming) and pass 2 reads them back instead of repeating one of its two
// exponentials.  The kernels are FP64-pipe bound and HBM has slack (profiles/r01d),
// so 32 B of traffic per pair buys ~38 of ~113 FP64 instructions of pass 2.
__device__ __forceinline__ void cache_store(double2 *p, double a, double b) {
    __stcs(p, make_double2(a, b));
}
__device__ __forceinline__ double2 cache_load(const double2 *p) { return __ldcs(p); }

__device__ __forceinline__ void load_tables(tab_fn *s, const tab_fn *g, int count) {
    const int words = count * (int)(sizeof(tab_fn) / 8);
    const double *src = reinterpret_cast<const double *>(g);
    double *dst = reinterpret_cast<double *>(s);
    for (int k = threadIdx.x; k < words; k += blockDim.x) dst[k] = src[k];
    __syncthreads();
}

// ---------------------------------------------------------------------------
// pass 1
// ---------------------------------------------------------------------------
template <typename Real, bool FAST, bool CACHE, bool NN>
__global__ void __launch_bounds__(EAM_T, EAM_MINB)
k_eam_rho(int n, const Atom4 *__restrict__ atoms,
          const uint8_t *__restrict__ types_ext, const int *__restrict__ counts,
          const uint32_t *__restrict__ slice_ptr, const uint32_t *__restrict__ col,
          const int *__restrict__ perm, EamDev m, Zhou1 z, tab_fn embed0,
          double *__restrict__ fprime, double *__restrict__ fembed,
          double *__restrict__ fprime_caller, double2 *__restrict__ pcache,
          const int *__restrict__ blk_first, double rcm) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    tab_fn *tabs = reinterpret_cast<tab_fn *>(smem_raw);
    __shared__ double s_etab[TAB_EXP_TAB];
    const int nn = m.n_el * m.n_el;
    if (!FAST) load_tables(tabs, m.rho, 2 * nn + m.n_el);
    if (FAST && sizeof(Real) == 8 && !CACHE) load_exp2_tab<ZEXP_RHO>(s_etab);
    int idx;
    if (!block_atom(n, blk_first, idx)) return;
    const Atom4 me = atoms[idx];
    const int ti = FAST ? 0 : (int)types_ext[idx];
    const int cnt = counts[idx];
    const size_t row0 = (size_t)slice_ptr[idx >> 5] * 32u + (idx & 31);
    const uint32_t *cp = col + row0;
    Real rho = Real(0);
    RowIter it(cp, atoms, cnt);
    Atom4 a;
    uint32_t c;
    while (it.next(a, c)) {
        Real dx, dy, dz, r, rinv, f, df;
        pair_r<Real>(me, a, dx, dy, dz, r, rinv);
        if (FAST) {
            if (sizeof(Real) == 8 && !CACHE) {
                double f8, d8;      // folded float64 term, fe inside the exponent
                zterm_eval<false, ZEXP_RHO>((double)r, (double)r * z.re, z.t_rho, s_etab, f8, d8);
                f = (Real)f8;
            } else {
                // g with unit prefactor; rho = fe * sum g
                zhou_exp<Real>(r, Real(1), (Real)z.beta, (Real)z.lamda, (Real)z.re, f, df);
                if (CACHE)
                    cache_store(pcache + row0 + (size_t)it.pos() * 32u, (double)f, (double)df);
            }
        } else {
            const int tj = (int)(c >> TAB_COL_TYPE_SHIFT);
            eval_pair_fn<Real, NN>(tabs[ti * m.n_el + tj], r, f, df, m.pool);
        }
        // lists with a skin: entries beyond the model's cutoff contribute exactly 0
        rho += (double)r < rcm ? f : Real(0);
    }
    if (FAST && !(sizeof(Real) == 8 && !CACHE)) rho *= (Real)z.fe;
    Real F, dF;
    // the fast path is all-zjw04 (tab_eam_create): call the embedding directly, the generic
    // switch would drag the 'nn' function's local arrays into this kernel
    if (FAST) zhou_embed<Real>(embed0.p, embed0.kind == TAB_FN_ZHOU_EMBED_XC, rho, F, dF);
    else eval_embed_fn<Real, NN>(tabs[2 * nn + ti], rho, F, dF, m.pool);
    fprime[idx] = (double)dF;
    fembed[idx] = (double)F;
    if (fprime_caller) fprime_caller[perm[idx]] = (double)dF;
}

// Atom4.w <- F'(rho) for every extended atom: owned atoms from pass 1, halo
// atoms from the values received from their owner ranks (caller order), periodic
// images from their source atom.
__global__ void k_spread_w(int n_owned, int n_loc, int n_ext,
                           const double *__restrict__ v,
                           const double *__restrict__ halo_v,
                           const int *__restrict__ perm,
                           const int *__restrict__ ghost_owner,
                           Atom4 *__restrict__ atoms, double scale = 1.0, double shift = 0.0) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_ext) return;
    const int o = e < n_loc ? e : ghost_owner[e - n_loc];
    // halo_v == NULL (recompute mode): F' of the outer halo is never needed by an own atom
    const double fp = o < n_owned ? v[o] : (halo_v ? halo_v[perm[o] - n_owned] : 0.0);
    atoms[e].w = fma(fp, scale, shift);      // (1, 0): F' itself; the lane-split kernels fold
}

// ---------------------------------------------------------------------------
// pass 2
// ---------------------------------------------------------------------------
template <typename Real, bool FAST, bool CACHE, bool NN>
__global__ void __launch_bounds__(EAM_T, EAM_MINB)
k_eam_force(int n, const Atom4 *__restrict__ atoms,
            const uint8_t *__restrict__ types_ext, const int *__restrict__ counts,
            const uint32_t *__restrict__ slice_ptr, const uint32_t *__restrict__ col,
            const int *__restrict__ perm, EamDev m, Zhou1 z,
            const double *__restrict__ fembed, double *__restrict__ eatom,
            double *__restrict__ forces, double *__restrict__ partial,
            const double2 *__restrict__ pcache, const int *__restrict__ blk_first,
            const int *__restrict__ own_mask, double rcm) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    tab_fn *tabs = reinterpret_cast<tab_fn *>(smem_raw);
    __shared__ double red[EAM_T / 32][7];
    __shared__ double s_etab[TAB_EXP_TAB];
    const int nn = m.n_el * m.n_el;
    if (!FAST) load_tables(tabs, m.rho, 2 * nn + m.n_el);
    if (FAST && sizeof(Real) == 8 && !CACHE) load_exp2_tab<ZEXP_FORCE>(s_etab);
    int idx;
    const bool active = block_atom(n, blk_first, idx);
    double acc[7] = {0, 0, 0, 0, 0, 0, 0};   // E, vxx, vyy, vzz, vyz, vxz, vxy
    if (active) {
        const Atom4 me = atoms[idx];
        const int ti = FAST ? 0 : (int)types_ext[idx];
        const int cnt = counts[idx];
        const size_t row0 = (size_t)slice_ptr[idx >> 5] * 32u + (idx & 31);
        const uint32_t *cp = col + row0;
        const Real fpi = (Real)me.w;
        Real fx = 0, fy = 0, fz = 0, ep = 0;
        Real vxx = 0, vyy = 0, vzz = 0, vyz = 0, vxz = 0, vxy = 0;
        RowIter it(cp, atoms, cnt);
        Atom4 a;
        uint32_t c;
        double2 pc_nx = make_double2(0.0, 0.0);
        if (CACHE && cnt > 0) pc_nx = cache_load(pcache + row0);
#pragma unroll kForceUnroll
        while (it.next(a, c)) {
            Real dx, dy, dz, r, rinv;
            pair_r<Real>(me, a, dx, dy, dz, r, rinv);
            const Real fpj = (Real)a.w;
            Real phi, dphi, der;   // der = dE/dr of the undirected pair seen from i
            if (FAST && sizeof(Real) == 8 && !CACHE) {
                // folded float64 terms: A and B inside the exponents
                const double x = (double)r * z.re;
                double ga, dga, gb, dgb;
                zterm_eval<true, ZEXP_FORCE>((double)r, x, z.t_a, s_etab, ga, dga);
                zterm_eval<true, ZEXP_FORCE>((double)r, x, z.t_b, s_etab, gb, dgb);
                phi = (Real)(ga - gb);
                dphi = (Real)(dga - dgb);
                der = (Real)fma(((double)fpi + (double)fpj) * z.fe_over_B, dgb, (double)dphi);
            } else if (FAST) {
                Real ga, dga, gb, dgb;
                zhou_exp<Real>(r, Real(1), (Real)z.alpha, (Real)z.kappa, (Real)z.re, ga, dga);
                if (CACHE) {
                    gb = (Real)pc_nx.x;
                    dgb = (Real)pc_nx.y;
                    if (it.pos() + 1 < cnt)
                        pc_nx = cache_load(pcache + row0 + (size_t)(it.pos() + 1) * 32u);
                } else {
                    zhou_exp<Real>(r, Real(1), (Real)z.beta, (Real)z.lamda, (Real)z.re, gb, dgb);
                }
                phi = (Real)z.A * ga - (Real)z.B * gb;
                dphi = (Real)z.A * dga - (Real)z.B * dgb;
                der = (fpi + fpj) * ((Real)z.fe * dgb) + dphi;
            } else {
                const int tj = (int)(c >> TAB_COL_TYPE_SHIFT);
                Real rij, drij, rji, drji;
                eval_pair_fn<Real, NN>(tabs[ti * m.n_el + tj], r, rij, drij, m.pool);
                if (ti == tj) drji = drij;
                else eval_pair_fn<Real, NN>(tabs[tj * m.n_el + ti], r, rji, drji, m.pool);
                eval_pair_fn<Real, NN>(tabs[nn + ti * m.n_el + tj], r, phi, dphi, m.pool);
                der = fpi * drij + fpj * drji + dphi;
            }
            if (!((double)r < rcm)) {      // beyond the model's cutoff (skin)
                der = Real(0);
                phi = Real(0);
            }
            const Real s = der * rinv;
            const Real gx = s * dx, gy = s * dy, gz = s * dz;
            fx += gx;
            fy += gy;
            fz += gz;
            ep += phi;
            vxx += gx * dx;
            vyy += gy * dy;
            vzz += gz * dz;
            vyz += gy * dz;
            vxz += gx * dz;
            vxy += gx * dy;
        }
        const double ei = fembed[idx] + 0.5 * (double)ep;
        const int o = perm[idx];
        if (eatom) eatom[o] = ei;
        if (forces) {
            forces[3 * o + 0] = (double)fx;
            forces[3 * o + 1] = (double)fy;
            forces[3 * o + 2] = (double)fz;
        }
        acc[0] = ei;
        // every undirected pair is visited from both ends: half of g (x) D each
        acc[1] = 0.5 * (double)vxx;
        acc[2] = 0.5 * (double)vyy;
        acc[3] = 0.5 * (double)vzz;
        acc[4] = 0.5 * (double)vyz;
        acc[5] = 0.5 * (double)vxz;
        acc[6] = 0.5 * (double)vxy;
        // spatial decomposition with recomputed inner-halo rows: only own atoms count
        if (own_mask && !own_mask[o])
#pragma unroll
            for (int q = 0; q < 7; ++q) acc[q] = 0.0;
    }
#pragma unroll
    for (int q = 0; q < 7; ++q) {
        double v = acc[q];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][q] = v;
    }
    __syncthreads();
    if (threadIdx.x < 7) {
        double v = 0;
        for (int w = 0; w < EAM_T / 32; ++w) v += red[w][threadIdx.x];
        partial[(size_t)blockIdx.x * 8 + threadIdx.x] = v;
    }
}


// ---------------------------------------------------------------------------
// ADP (nn/eam/adp.py:315-586): per centre atom and per NEIGHBOUR SPECIES T
//   mu_T = sum_{p in T} u(r_p) D_p ,  lam_T = sum_{p in T} w(r_p) D_p (x) D_p
//   E_i += sum_T [ 1/2 |mu_T|^2 + 1/2 sum_ab lam_T,ab^2 - 1/6 tr(lam_T)^2 ]
// (the per-term squaring is the reference's behaviour: adp.py:370-389,455-494).
// Moments are stored per extended atom as [n_el][9] = mux muy muz lxx lyy lzz
// lyz lxz lxy; pass 2 needs the neighbour's moments w.r.t. the centre's species.
// ---------------------------------------------------------------------------
template <typename Real, bool NN>
__global__ void __launch_bounds__(EAM_T)
k_adp_rho(int n, const Atom4 *__restrict__ atoms,
          const uint8_t *__restrict__ types_ext, const int *__restrict__ counts,
          const uint32_t *__restrict__ slice_ptr, const uint32_t *__restrict__ col,
          EamDev m, double *__restrict__ fprime, double *__restrict__ fembed,
          double *__restrict__ moments, const int *__restrict__ blk_first, double rcm) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    tab_fn *tabs = reinterpret_cast<tab_fn *>(smem_raw);
    const int nn = m.n_el * m.n_el;
    load_tables(tabs, m.rho, 4 * nn + m.n_el);
    const tab_fn *t_dip = tabs + 2 * nn + m.n_el, *t_quad = t_dip + nn;
    int idx;
    if (!block_atom(n, blk_first, idx)) return;
    const Atom4 me = atoms[idx];
    const int ti = (int)types_ext[idx];
    const int cnt = counts[idx];
    const uint32_t *cp = col + ((size_t)slice_ptr[idx >> 5] * 32u + (idx & 31));
    Real rho = Real(0);
    Real mom[ADP_MAX_EL][9];
#pragma unroll
    for (int t = 0; t < ADP_MAX_EL; ++t)
#pragma unroll
        for (int q = 0; q < 9; ++q) mom[t][q] = Real(0);
    RowIter it(cp, atoms, cnt);
    Atom4 a;
    uint32_t c;
    while (it.next(a, c)) {
        Real dx, dy, dz, r, rinv, f, df, u, du, w, dw;
        pair_r<Real>(me, a, dx, dy, dz, r, rinv);
        const int tj = (int)(c >> TAB_COL_TYPE_SHIFT);
        eval_pair_fn<Real, NN>(tabs[ti * m.n_el + tj], r, f, df, m.pool);
        eval_pair_fn<Real, NN>(t_dip[ti * m.n_el + tj], r, u, du, m.pool);
        eval_pair_fn<Real, NN>(t_quad[ti * m.n_el + tj], r, w, dw, m.pool);
        if (!((double)r < rcm)) f = u = w = Real(0);
        rho += f;
#pragma unroll
        for (int t = 0; t < ADP_MAX_EL; ++t) {
            const Real sel = (t == tj) ? Real(1) : Real(0);
            const Real us = u * sel, ws = w * sel;
            mom[t][0] += us * dx;
            mom[t][1] += us * dy;
            mom[t][2] += us * dz;
            mom[t][3] += ws * dx * dx;
            mom[t][4] += ws * dy * dy;
            mom[t][5] += ws * dz * dz;
            mom[t][6] += ws * dy * dz;
            mom[t][7] += ws * dx * dz;
            mom[t][8] += ws * dx * dy;
        }
    }
    Real F, dF;
    eval_embed_fn<Real, NN>(tabs[2 * nn + ti], rho, F, dF, m.pool);
    Real eadp = Real(0);
#pragma unroll
    for (int t = 0; t < ADP_MAX_EL; ++t) {
        if (t >= m.n_el) break;
        const Real *q = mom[t];
        const Real tr = q[3] + q[4] + q[5];
        eadp += Real(0.5) * (q[0] * q[0] + q[1] * q[1] + q[2] * q[2]) +
                Real(0.5) * (q[3] * q[3] + q[4] * q[4] + q[5] * q[5] +
                             Real(2) * (q[6] * q[6] + q[7] * q[7] + q[8] * q[8])) -
                tr * tr / Real(6);
#pragma unroll
        for (int k = 0; k < 9; ++k)
            moments[((size_t)idx * m.n_el + t) * 9 + k] = (double)q[k];
    }
    fprime[idx] = (double)dF;
    fembed[idx] = (double)(F + eadp);
}

// moments of the library's periodic images (records [n_loc, n_ext)) = their source atom's;
// halo atoms [n, n_loc) keep zeros (decomposition: tab_eam_eval_dd recomputes the inner halo
// as part of the row-owning group, the outer halo's moments are never read by an own atom)
__global__ void k_adp_spread(int n_loc, int n_ext, int stride,
                             const int *__restrict__ ghost_owner,
                             double *__restrict__ moments) {
    const int e = n_loc + blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_ext) return;
    const int o = ghost_owner[e - n_loc];
    for (int k = 0; k < stride; ++k)
        moments[(size_t)e * stride + k] = moments[(size_t)o * stride + k];
}

template <typename Real, bool NN>
__global__ void __launch_bounds__(EAM_T)
k_adp_force(int n, const Atom4 *__restrict__ atoms,
            const uint8_t *__restrict__ types_ext, const int *__restrict__ counts,
            const uint32_t *__restrict__ slice_ptr, const uint32_t *__restrict__ col,
            const int *__restrict__ perm, EamDev m,
            const double *__restrict__ moments, const double *__restrict__ fembed,
            double *__restrict__ eatom, double *__restrict__ forces,
            double *__restrict__ partial, const int *__restrict__ blk_first,
            const int *__restrict__ own_mask, double rcm) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    tab_fn *tabs = reinterpret_cast<tab_fn *>(smem_raw);
    __shared__ double red[EAM_T / 32][7];
    const int nn = m.n_el * m.n_el;
    load_tables(tabs, m.rho, 4 * nn + m.n_el);
    const tab_fn *t_dip = tabs + 2 * nn + m.n_el, *t_quad = t_dip + nn;
    int idx;
    const bool active = block_atom(n, blk_first, idx);
    double acc[7] = {0, 0, 0, 0, 0, 0, 0};
    if (active) {
        const Atom4 me = atoms[idx];
        const int ti = (int)types_ext[idx];
        const int cnt = counts[idx];
        const uint32_t *cp = col + ((size_t)slice_ptr[idx >> 5] * 32u + (idx & 31));
        const Real fpi = (Real)me.w;
        const int stride = m.n_el * 9;
        const double *mine = moments + (size_t)idx * stride;
        Real fx = 0, fy = 0, fz = 0, ep = 0;
        Real vxx = 0, vyy = 0, vzz = 0, vyz = 0, vxz = 0, vxy = 0;
        RowIter it(cp, atoms, cnt);
        Atom4 a;
        uint32_t c;
        while (it.next(a, c)) {
            Real dx, dy, dz, r, rinv;
            pair_r<Real>(me, a, dx, dy, dz, r, rinv);
            const int j = (int)(c & TAB_COL_IDX_MASK);
            const int tj = (int)(c >> TAB_COL_TYPE_SHIFT);
            const Real fpj = (Real)a.w;
            Real rij, drij, rji, drji, phi, dphi, u, du, w, dw;
            eval_pair_fn<Real, NN>(tabs[ti * m.n_el + tj], r, rij, drij, m.pool);
            if (ti == tj) drji = drij;
            else eval_pair_fn<Real, NN>(tabs[tj * m.n_el + ti], r, rji, drji, m.pool);
            eval_pair_fn<Real, NN>(tabs[nn + ti * m.n_el + tj], r, phi, dphi, m.pool);
            eval_pair_fn<Real, NN>(t_dip[ti * m.n_el + tj], r, u, du, m.pool);
            eval_pair_fn<Real, NN>(t_quad[ti * m.n_el + tj], r, w, dw, m.pool);
            if (!((double)r < rcm))
                drij = drji = phi = dphi = u = du = w = dw = Real(0);
            // EAM part, symmetric in the pair
            const Real s = (fpi * drij + fpj * drji + dphi) * rinv;
            Real gx = s * dx, gy = s * dy, gz = s * dz;
            vxx += Real(0.5) * gx * dx;
            vyy += Real(0.5) * gy * dy;
            vzz += Real(0.5) * gz * dz;
            vyz += Real(0.5) * gy * dz;
            vxz += Real(0.5) * gx * dz;
            vxy += Real(0.5) * gx * dy;
            ep += phi;
            // ADP part: own moments w.r.t. species tj, neighbour's w.r.t. species ti
            const double *qi = mine + tj * 9;
            const double *qj = moments + (size_t)j * stride + ti * 9;
            Real mi[9], mj[9];
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                mi[k] = (Real)qi[k];
                mj[k] = (Real)qj[k];
            }
            const Real tri = (mi[3] + mi[4] + mi[5]) / Real(3);
            const Real trj = (mj[3] + mj[4] + mj[5]) / Real(3);
            // M = lam - tr/3 I  (xx yy zz yz xz xy)
            const Real Mi[6] = {mi[3] - tri, mi[4] - tri, mi[5] - tri, mi[6], mi[7], mi[8]};
            const Real Mj[6] = {mj[3] - trj, mj[4] - trj, mj[5] - trj, mj[6], mj[7], mj[8]};
            auto Mv = [&](const Real *M, Real &x, Real &y, Real &z) {
                x = M[0] * dx + M[5] * dy + M[4] * dz;
                y = M[5] * dx + M[1] * dy + M[3] * dz;
                z = M[4] * dx + M[3] * dy + M[2] * dz;
            };
            Real ax, ay, az, bx, by, bz;
            Mv(Mi, ax, ay, az);
            Mv(Mj, bx, by, bz);
            const Real mudi = mi[0] * dx + mi[1] * dy + mi[2] * dz;
            const Real mudj = mj[0] * dx + mj[1] * dy + mj[2] * dz;
            const Real dMdi = ax * dx + ay * dy + az * dz;
            const Real dMdj = bx * dx + by * dy + bz * dz;
            // directed gradient g_{i->j} (own moments): virial
            const Real ci = (du * mudi + dw * dMdi) * rinv;
            const Real gix = ci * dx + u * mi[0] + Real(2) * w * ax;
            const Real giy = ci * dy + u * mi[1] + Real(2) * w * ay;
            const Real giz = ci * dz + u * mi[2] + Real(2) * w * az;
            vxx += gix * dx;
            vyy += giy * dy;
            vzz += giz * dz;
            vyz += Real(0.5) * (giy * dz + giz * dy);
            vxz += Real(0.5) * (gix * dz + giz * dx);
            vxy += Real(0.5) * (gix * dy + giy * dx);
            // force: g_{i->j}(D) - g_{j->i}(-D)
            const Real cj = (du * mudj - dw * dMdj) * rinv;
            gx += gix - (cj * dx + u * mj[0] - Real(2) * w * bx);
            gy += giy - (cj * dy + u * mj[1] - Real(2) * w * by);
            gz += giz - (cj * dz + u * mj[2] - Real(2) * w * bz);
            fx += gx;
            fy += gy;
            fz += gz;
        }
        const double ei = fembed[idx] + 0.5 * (double)ep;
        const int o = perm[idx];
        if (eatom) eatom[o] = ei;
        if (forces) {
            forces[3 * o + 0] = (double)fx;
            forces[3 * o + 1] = (double)fy;
            forces[3 * o + 2] = (double)fz;
        }
        acc[0] = ei;
        acc[1] = (double)vxx;
        acc[2] = (double)vyy;
        acc[3] = (double)vzz;
        acc[4] = (double)vyz;
        acc[5] = (double)vxz;
        acc[6] = (double)vxy;
        if (own_mask && !own_mask[o])
#pragma unroll
            for (int q = 0; q < 7; ++q) acc[q] = 0.0;
    }
#pragma unroll
    for (int q = 0; q < 7; ++q) {
        double v = acc[q];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][q] = v;
    }
    __syncthreads();
    if (threadIdx.x < 7) {
        double v = 0;
        for (int w = 0; w < EAM_T / 32; ++w) v += red[w][threadIdx.x];
        partial[(size_t)blockIdx.x * 8 + threadIdx.x] = v;
    }
}

// fixed-order final reduction: energy[0], virial[9].  Batch handles: block s of the grid
// reduces the partial rows [struct_blk[s], struct_blk[s+1]) into energy[s], virial[9 s..].
__global__ void __launch_bounds__(256)
k_reduce_partials(int nblk, const double *__restrict__ partial,
                  double *__restrict__ energy, double *__restrict__ virial,
                  const int *__restrict__ struct_blk) {
    __shared__ double sm[256][7];
    double a[7] = {0, 0, 0, 0, 0, 0, 0};
    int b_lo = 0;
    if (struct_blk) {
        b_lo = struct_blk[blockIdx.x];
        nblk = struct_blk[blockIdx.x + 1];
        if (energy) energy += blockIdx.x;
        if (virial) virial += 9 * (size_t)blockIdx.x;
    }
    for (int b = b_lo + threadIdx.x; b < nblk; b += 256)
#pragma unroll
        for (int q = 0; q < 7; ++q) a[q] += partial[(size_t)b * 8 + q];
#pragma unroll
    for (int q = 0; q < 7; ++q) sm[threadIdx.x][q] = a[q];
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s)
#pragma unroll
            for (int q = 0; q < 7; ++q) sm[threadIdx.x][q] += sm[threadIdx.x + s][q];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        if (energy) energy[0] = sm[0][0];
        if (virial) {
            const double xx = sm[0][1], yy = sm[0][2], zz = sm[0][3], yz = sm[0][4],
                         xz = sm[0][5], xy = sm[0][6];
            virial[0] = xx; virial[1] = xy; virial[2] = xz;
            virial[3] = xy; virial[4] = yy; virial[5] = yz;
            virial[6] = xz; virial[7] = yz; virial[8] = zz;
        }
    }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
extern "C" int tab_eam_create(tab_model **out, int32_t kind, int32_t n_el,
                              const tab_fn *rho, const tab_fn *phi,
                              const tab_fn *embed, const tab_fn *dipole,
                              const tab_fn *quadrupole) {
    if (!out || !rho || !phi || !embed || n_el < 1 || n_el > TAB_MAX_ELEMENTS) {
        tab_set_error("tab_eam_create: bad argument");
        return TAB_EINVAL;
    }
    if (kind != TAB_EAM_ALLOY && kind != TAB_EAM_FS && kind != TAB_EAM_ADP) {
        tab_set_error("tab_eam_create: unknown kind %d", kind);
        return TAB_EINVAL;
    }
    if (kind == TAB_EAM_ADP && (!dipole || !quadrupole || n_el > ADP_MAX_EL)) {
        tab_set_error("tab_eam_create: ADP needs dipole+quadrupole tables and "
                      "at most %d elements", ADP_MAX_EL);
        return TAB_EUNSUPPORTED;
    }
    if (n_el > 8) {
        tab_set_error("tab_eam_create: more than 8 elements not supported");
        return TAB_EUNSUPPORTED;
    }
    tab_model *m = new tab_model();
    m->kind = kind;
    m->n_el = n_el;
    const int nn = n_el * n_el;
    const bool adp = kind == TAB_EAM_ADP;
    const size_t count = (size_t)2 * nn + n_el + (adp ? 2 * nn : 0);
    int rc = m->tables.ensure(count * sizeof(tab_fn));
    if (rc != TAB_OK) {
        delete m;
        return rc;
    }
    tab_fn *host = new tab_fn[count];
    memcpy(host, rho, nn * sizeof(tab_fn));
    memcpy(host + nn, phi, nn * sizeof(tab_fn));
    memcpy(host + 2 * nn, embed, n_el * sizeof(tab_fn));
    if (adp) {
        memcpy(host + 2 * nn + n_el, dipole, nn * sizeof(tab_fn));
        memcpy(host + 3 * nn + n_el, quadrupole, nn * sizeof(tab_fn));
    }
    // device representation: the r_eq slots hold 1/r_eq (potentials.cuh)
    for (size_t k = 0; k < count; ++k) {
        tab_fn &f = host[k];
        if (f.kind == TAB_FN_MLP) {
            m->has_mlp_fn = true;
            const int nh = (int)f.p[0];
            bool bad = nh < 1 || nh > TAB_MLP_FN_MAXL;
            for (int l = 0; !bad && l < nh; ++l)
                bad = f.p[2 + l] < 1 || f.p[2 + l] > TAB_MLP_FN_MAXW;
            if (bad) {
                tab_set_error("tab_eam_create: 'nn' function with %d hidden layers / a layer "
                              "wider than %d", nh, TAB_MLP_FN_MAXW);
                delete[] host;
                m->tables.release();
                delete m;
                return TAB_EUNSUPPORTED;
            }
        }
        if (f.kind == TAB_FN_ZHOU_RHO) f.p[3] = 1.0 / f.p[3];
        else if (f.kind == TAB_FN_ZHOU_PHI) f.p[6] = 1.0 / f.p[6];
        else if (f.kind == TAB_FN_ZHOU_PHI_MIX) {
            f.p[6] = 1.0 / f.p[6];
            f.p[10] = 1.0 / f.p[10];
            f.p[17] = 1.0 / f.p[17];
            f.p[21] = 1.0 / f.p[21];
        }
    }
    cudaError_t e = cudaMemcpy(m->tables.p, host, count * sizeof(tab_fn),
                               cudaMemcpyHostToDevice);
    delete[] host;
    if (e != cudaSuccess) {
        tab_set_error("tab_eam_create: cudaMemcpy -> %s", cudaGetErrorString(e));
        m->tables.release();
        delete m;
        return TAB_ECUDA;
    }
    m->embed0 = embed[0];
    // fast path: one element, zjw04 rho/phi/embed whose B-term parameters agree
    if (n_el == 1 && rho[0].kind == TAB_FN_ZHOU_RHO && phi[0].kind == TAB_FN_ZHOU_PHI &&
        (embed[0].kind == TAB_FN_ZHOU_EMBED || embed[0].kind == TAB_FN_ZHOU_EMBED_XC) &&
        rho[0].p[1] == phi[0].p[4] && rho[0].p[2] == phi[0].p[5] &&
        rho[0].p[3] == phi[0].p[6]) {
        m->zhou1 = true;
        Zhou1 &z = m->z1;
        const double re = rho[0].p[3];
        z.fe = rho[0].p[0];       // f_eq
        z.beta = rho[0].p[1];
        z.lamda = rho[0].p[2];
        z.re = 1.0 / re;          // 1 / r_eq
        z.A = phi[0].p[0];
        z.alpha = phi[0].p[1];
        z.kappa = phi[0].p[2];
        z.B = phi[0].p[3];
        m->z1_folded = z.fe > 0.0 && z.A > 0.0 && z.B > 0.0 && re > 0.0;
        if (m->z1_folded) {
            zterm_fold(z.t_rho, z.fe, z.beta, z.lamda, re);
            zterm_fold(z.t_a, z.A, z.alpha, z.kappa, re);
            zterm_fold(z.t_b, z.B, z.beta, z.lamda, re);
            z.fe_over_B = z.fe / z.B;
            m->zp = eamz_new_pair(z);
        }
    }
    *out = m;
    return TAB_OK;
}

// accessor for hessian.cu (tab_model is private to this file)
int tab_eam_tables(tab_model *m, const tab_fn **rho, const tab_fn **phi,
                   const tab_fn **embed, int *n_el, int *kind) {
    if (m->has_mlp_fn) {
        tab_set_error("analytic Hessian: 'nn' (MLP) functions carry no second derivative");
        return TAB_EUNSUPPORTED;
    }
    if (m->no_hessian) {
        tab_set_error("analytic Hessian: a function kind of this model carries no second "
                      "derivative evaluator");
        return TAB_EUNSUPPORTED;
    }
    const int nn = m->n_el * m->n_el;
    *rho = m->tables.as<tab_fn>();
    *phi = *rho + nn;
    *embed = *rho + 2 * nn;
    *n_el = m->n_el;
    *kind = m->kind;
    return TAB_OK;
}

extern "C" int tab_eam_set_splines(tab_model *m, const double *h_coeffs,
                                   int64_t n_doubles) {
    if (!m || !h_coeffs || n_doubles <= 0) {
        tab_set_error("tab_eam_set_splines: bad argument");
        return TAB_EINVAL;
    }
    TAB_TRY(m->pool.ensure(sizeof(double) * (size_t)n_doubles));
    TAB_CUDA(cudaMemcpy(m->pool.p, h_coeffs, sizeof(double) * (size_t)n_doubles,
                        cudaMemcpyHostToDevice));
    return TAB_OK;
}

const double *tab_eam_pool(tab_model *m) { return m->pool.as<double>(); }

extern "C" int tab_model_free(tab_model *m) {
    if (!m) return TAB_OK;
    m->tables.release();
    m->pool.release();
    eamz_free_pair(m->zp);
    delete m;
    return TAB_OK;
}

// ---------------------------------------------------------------------------
// optional per-kernel timing with CUDA events on the launching stream
// (bench.py's roofline numbers); off by default.
// ---------------------------------------------------------------------------
#define PROF_MAX_CALLS 512
#define PROF_MARKS 5
static bool g_prof_on = false;
static int g_prof_calls = 0;
static cudaEvent_t g_prof_ev[PROF_MAX_CALLS][PROF_MARKS];
static int g_prof_created = 0;

static inline void prof_mark(int k, cudaStream_t st) {
    if (!g_prof_on || g_prof_calls >= PROF_MAX_CALLS) return;
    while (g_prof_created <= g_prof_calls) {
        for (int q = 0; q < PROF_MARKS; ++q) cudaEventCreate(&g_prof_ev[g_prof_created][q]);
        ++g_prof_created;
    }
    cudaEventRecord(g_prof_ev[g_prof_calls][k], st);
}

extern "C" int tab_profile_enable(int32_t on) {
    g_prof_on = on != 0;
    g_prof_calls = 0;
    return TAB_OK;
}

// ms[0..3] = mean duration of k_eam_rho, k_spread_w, k_eam_force, k_reduce_partials
extern "C" int tab_profile_read(double *ms, int32_t *calls) {
    if (!ms || !calls) return TAB_EINVAL;
    TAB_CUDA(cudaDeviceSynchronize());
    for (int q = 0; q < PROF_MARKS - 1; ++q) ms[q] = 0.0;
    for (int c = 0; c < g_prof_calls; ++c)
        for (int q = 0; q < PROF_MARKS - 1; ++q) {
            float t = 0.f;
            TAB_CUDA(cudaEventElapsedTime(&t, g_prof_ev[c][q], g_prof_ev[c][q + 1]));
            ms[q] += t;
        }
    if (g_prof_calls > 0)
        for (int q = 0; q < PROF_MARKS - 1; ++q) ms[q] /= g_prof_calls;
    *calls = g_prof_calls;
    return TAB_OK;
}

struct EamLaunch {
    EamDev dev;
    Zhou1 z;
    size_t smem;
    int nblk;
    double *fprime, *fembed;
    const int *blk_first;    // batch handles: blocks cut at structure boundaries, else NULL
    const int *struct_blk;
    int n_red;               // grid of k_reduce_partials (1, or the number of structures)
    double rcm;              // mask radius: +inf, or the model's cutoff for lists with a skin
};

// batch handles: the pair-kernel block table for blocks of EAM_T atoms (cached in nbr)
static int ensure_batch_blocks(tab_nbr *nbr, cudaStream_t st) {
    if (nbr->n_struct <= 0 || nbr->blk_T == EAM_T) return TAB_OK;
    std::vector<int> first, sblk(nbr->n_struct + 1);
    for (int s = 0; s < nbr->n_struct; ++s) {
        sblk[s] = (int)first.size();
        for (int a = nbr->h_struct_off[s]; a < nbr->h_struct_off[s + 1]; a += EAM_T)
            first.push_back(a);
    }
    sblk[nbr->n_struct] = (int)first.size();
    nbr->n_blk = (int)first.size();
    first.push_back(nbr->h_struct_off[nbr->n_struct]);
    TAB_TRY(nbr->blk_first.ensure(sizeof(int) * first.size()));
    TAB_TRY(nbr->struct_blk.ensure(sizeof(int) * sblk.size()));
    TAB_CUDA(cudaMemcpyAsync(nbr->blk_first.p, first.data(), sizeof(int) * first.size(),
                             cudaMemcpyHostToDevice, st));
    TAB_CUDA(cudaMemcpyAsync(nbr->struct_blk.p, sblk.data(), sizeof(int) * sblk.size(),
                             cudaMemcpyHostToDevice, st));
    TAB_CUDA(cudaStreamSynchronize(st));     // pageable host vectors
    nbr->blk_T = EAM_T;
    return TAB_OK;
}

static int eam_prepare(tab_model *m, tab_nbr *nbr, bool fast, EamLaunch &L,
                       cudaStream_t st) {
    const int n = nbr->n;
    L.nblk = (n + EAM_T - 1) / EAM_T;
    L.blk_first = nullptr;
    L.struct_blk = nullptr;
    L.n_red = 1;
    if (nbr->n_struct > 0) {
        TAB_TRY(ensure_batch_blocks(nbr, st));
        L.nblk = nbr->n_blk;
        L.blk_first = nbr->blk_first.as<int>();
        L.struct_blk = nbr->struct_blk.as<int>();
        L.n_red = nbr->n_struct;
    }
    TAB_TRY(nbr->rho.ensure(sizeof(double) * 2 * (size_t)n));
    TAB_TRY(nbr->partial.ensure(sizeof(double) * 8 * (size_t)L.nblk));
    L.fprime = nbr->rho.as<double>();
    L.fembed = L.fprime + n;
    L.dev.kind = m->kind;
    L.dev.n_el = m->n_el;
    const int nn = m->n_el * m->n_el;
    L.dev.rho = m->tables.as<tab_fn>();
    L.dev.phi = L.dev.rho + nn;
    L.dev.embed = L.dev.rho + 2 * nn;
    L.dev.pool = m->pool.as<double>();
    L.z = m->z1;
    L.rcm = nbr->skin_built > 0.0 ? nbr->rc_model : (double)INFINITY;
    L.smem = fast ? 0 : (size_t)(2 * nn + m->n_el) * sizeof(tab_fn);
    return TAB_OK;
}

// The per-pair cache is OFF by default: measured on B200 (1 M atoms) it makes pass 1
// slower (0.38 -> 0.56 ms, the 1.46 GB stream) by more than pass 2 gains (0.73 -> 0.71 ms;
// that pass is co-limited by L1 gather wavefronts, not by FP64 alone).  TAB_EAM_PAIR_CACHE=1
// enables it for experiments (profiles/README.md, round 1).
static bool pair_cache_enabled() {
    static int on = -1;
    if (on < 0) {
        const char *e = getenv("TAB_EAM_PAIR_CACHE");
        on = (e && e[0] == '1') ? 1 : 0;
    }
    return on == 1;
}

// the float64 fast path folds prefactors into exponents and drops the underflow guard
static bool zhou1_f64_ok(const tab_model *m, const tab_nbr *nbr) {
    if (!m->zhou1 || !m->z1_folded) return false;
    const double xmax = nbr->grid.rc * m->z1.re;
    const double bmax = m->z1.alpha > m->z1.beta ? m->z1.alpha : m->z1.beta;
    return bmax * (xmax + 1.0) < 600.0;
}

// ---------------------------------------------------------------------------
// lane-split kernels of the single-element zjw04 model (eam_fast.cuh)
// ---------------------------------------------------------------------------
#include "eam_fast.cuh"

static ZPair *eamz_new_pair(const Zhou1 &z) {
    ZPair *p = new ZPair();
    zpair_fold(*p, z);
    return p;
}
static void eamz_free_pair(ZPair *p) { delete p; }

static int env_int(const char *name, int dflt) {
    const char *e = getenv(name);
    return e && *e ? atoi(e) : dflt;
}

// lanes per atom of the fast kernels (TAB_EAMZ_L = 1 | 2 | 4 | 8 for A/B runs; 0 = the
// thread-per-atom kernels of round 1)
// (read on every call: the tests and the A/B tools switch them inside one process)
static int eamz_lanes() {
    const int L = env_int("TAB_EAMZ_L", EAMZ_L);
    return (L == 0 || L == 1 || L == 2 || L == 4 || L == 8) ? L : EAMZ_L;
}


static bool eamz_usable(const tab_model *m, const tab_nbr *nbr, int precision) {
    if (eamz_lanes() == 0 || !m->zhou1 || !m->z1_folded || !m->zp || nbr->n_struct > 0)
        return false;
    if (precision == TAB_PRECISION_HIGH) return zhou1_f64_ok(m, nbr);
    // float32: exponents of the three terms must stay inside ex2's range
    return zhou1_f64_ok(m, nbr) && (m->z1.alpha > m->z1.beta ? m->z1.alpha : m->z1.beta) *
                                           (nbr->grid.rc * m->z1.re + 1.0) < 80.0;
}

static QScale eamz_qscale(const tab_model *m, const tab_nbr *nbr) {
    QScale q;
    const double d = nbr->q_delta;
    q.x_per_q = (float)(d * m->z1.re);
    q.eps_q = (float)(1e-8 / (d * d));
    q.rc2_q = (float)(tab_mask_rc2(nbr) / (d * d));
    q.pad_ = 0.f;
    q.ox = nbr->q_origin[0];
    q.oy = nbr->q_origin[1];
    q.oz = nbr->q_origin[2];
    q.ddelta = d;
    return q;
}

#define EAMZ_FOR_L(L_, CALL)                        \
    switch (L_) {                                   \
    case 1: { constexpr int LL = 1; CALL; } break;  \
    case 2: { constexpr int LL = 2; CALL; } break;  \
    case 4: { constexpr int LL = 4; CALL; } break;  \
    default: { constexpr int LL = 8; CALL; } break; \
    }

template <int L>
static int eamz_pass1_L(tab_model *m, tab_nbr *nbr, int precision, double *d_fprime_caller,
                        double *fprime, double *fembed, cudaStream_t st) {
    LsView v;
    TAB_TRY(ensure_lanesplit<L>(nbr, st, v));
    const int n = nbr->n;
    constexpr int G = 32 / L;
    const long long threads = (long long)((n + G - 1) / G) * 32;
    const int nblk = (int)((threads + EAMZ_T - 1) / EAMZ_T);
    if (precision == TAB_PRECISION_HIGH) {
        k_eamz_rho<L><<<nblk, EAMZ_T, 0, st>>>(
            n, nbr->atoms.as<Atom4>(), v.ptr, v.w, v.col, nbr->perm.as<int>(), *m->zp,
            tab_mask_rc2(nbr), m->embed0, fprime, fembed, d_fprime_caller);
    } else {
        TAB_TRY(tab_nbr_ensure_rec16(nbr, st));
        k_eamz_rho_f32<L><<<nblk, EAMZ_T, 0, st>>>(
            n, nbr->rec16.as<Rec16>(), v.ptr, v.w, v.col, nbr->perm.as<int>(), *m->zp,
            eamz_qscale(m, nbr), m->embed0, fprime, fembed, d_fprime_caller);
    }
    TAB_LAUNCH_CHECK();
    return TAB_OK;
}

template <int L>
static int eamz_pass2_L(tab_model *m, tab_nbr *nbr, int precision, const double *fembed,
                        double *d_eatom, double *d_forces, const int *d_own_mask, int *nblk_out,
                        cudaStream_t st) {
    LsView v;
    TAB_TRY(ensure_lanesplit<L>(nbr, st, v));
    const int n = nbr->n;
    constexpr int G = 32 / L;
    const long long threads = (long long)((n + G - 1) / G) * 32;
    const int nblk = (int)((threads + EAMZ_T - 1) / EAMZ_T);
    *nblk_out = nblk;
    TAB_TRY(nbr->partial.ensure(sizeof(double) * 8 * (size_t)nblk));
    if (precision == TAB_PRECISION_HIGH) {
        k_eamz_force<L><<<nblk, EAMZ_T, 0, st>>>(
            n, nbr->atoms.as<Atom4>(), v.ptr, v.w, v.col, nbr->perm.as<int>(), *m->zp,
            tab_mask_rc2(nbr), fembed, d_eatom, d_forces, nbr->partial.as<double>(), d_own_mask);
    } else {
        k_eamz_force_f32<L><<<nblk, EAMZ_T, 0, st>>>(
            n, nbr->rec16.as<Rec16>(), v.ptr, v.w, v.col, nbr->perm.as<int>(), *m->zp,
            eamz_qscale(m, nbr), fembed, d_eatom, d_forces, nbr->partial.as<double>(),
            d_own_mask);
    }
    TAB_LAUNCH_CHECK();
    return TAB_OK;
}

template <typename Real, bool FAST>
static int eam_pass1(tab_model *m, tab_nbr *nbr, double *d_fprime_caller,
                     cudaStream_t st) {
    EamLaunch L;
    TAB_TRY(eam_prepare(m, nbr, FAST, L, st));
    const int prec = sizeof(Real) == 8 ? TAB_PRECISION_HIGH : TAB_PRECISION_MEDIUM;
    if (FAST && eamz_usable(m, nbr, prec)) {
        nbr->pcache_valid = false;
        prof_mark(0, st);
        EAMZ_FOR_L(eamz_lanes(), TAB_TRY(eamz_pass1_L<LL>(m, nbr, prec, d_fprime_caller,
                                                          L.fprime, L.fembed, st)));
        prof_mark(1, st);
        return TAB_OK;
    }
    const bool cache = FAST && sizeof(Real) == 8 && pair_cache_enabled();
    if (cache) TAB_TRY(nbr->pcache.ensure(sizeof(double2) * 32 * (size_t)(nbr->ell_rows + 1)));
    nbr->pcache_valid = false;
    prof_mark(0, st);
    // 'nn' functions (NN) exist only on the generic path; the pair cache only on the fast one
    if (cache)
        k_eam_rho<Real, FAST, FAST, false><<<L.nblk, EAM_T, L.smem, st>>>(
            nbr->n, nbr->atoms.as<Atom4>(), nbr->types_ext.as<uint8_t>(),
            nbr->counts.as<int>(), nbr->slice_ptr.as<uint32_t>(), nbr->col.as<uint32_t>(),
            nbr->perm.as<int>(), L.dev, L.z, m->embed0, L.fprime, L.fembed,
            d_fprime_caller, nbr->pcache.as<double2>(), L.blk_first, L.rcm);
    else if (!FAST && m->has_mlp_fn)
        k_eam_rho<Real, false, false, true><<<L.nblk, EAM_T, L.smem, st>>>(
            nbr->n, nbr->atoms.as<Atom4>(), nbr->types_ext.as<uint8_t>(),
            nbr->counts.as<int>(), nbr->slice_ptr.as<uint32_t>(), nbr->col.as<uint32_t>(),
            nbr->perm.as<int>(), L.dev, L.z, m->embed0, L.fprime, L.fembed,
            d_fprime_caller, nullptr, L.blk_first, L.rcm);
    else
        k_eam_rho<Real, FAST, false, false><<<L.nblk, EAM_T, L.smem, st>>>(
            nbr->n, nbr->atoms.as<Atom4>(), nbr->types_ext.as<uint8_t>(),
            nbr->counts.as<int>(), nbr->slice_ptr.as<uint32_t>(), nbr->col.as<uint32_t>(),
            nbr->perm.as<int>(), L.dev, L.z, m->embed0, L.fprime, L.fembed,
            d_fprime_caller, nullptr, L.blk_first, L.rcm);
    TAB_LAUNCH_CHECK();
    nbr->pcache_valid = cache;
    prof_mark(1, st);
    return TAB_OK;
}

template <typename Real, bool FAST>
static int eam_pass2(tab_model *m, tab_nbr *nbr, const double *d_fprime_halo,
                     double *d_energy, double *d_eatom, double *d_forces,
                     double *d_virial, cudaStream_t st, const int *d_own_mask = nullptr) {
    EamLaunch L;
    TAB_TRY(eam_prepare(m, nbr, FAST, L, st));
    if (nbr->n_halo > 0 && !d_fprime_halo && !d_own_mask) {
        tab_set_error("tab_eam_pass2: halo atoms present but no halo F' given");
        return TAB_EINVAL;
    }
    const int prec = sizeof(Real) == 8 ? TAB_PRECISION_HIGH : TAB_PRECISION_MEDIUM;
    if (FAST && eamz_usable(m, nbr, prec)) {
        // w = F' fe / B - 1/2 (see k_eamz_force)
        if (prec == TAB_PRECISION_HIGH) {
            k_spread_w<<<(nbr->n_ext + 255) / 256, 256, 0, st>>>(
                nbr->n, nbr->n_loc, nbr->n_ext, L.fprime, d_fprime_halo, nbr->perm.as<int>(),
                nbr->ghost_owner.as<int>(), nbr->atoms.as<Atom4>(), m->zp->fe_over_B, -0.5);
        } else {
            TAB_TRY(tab_nbr_ensure_rec16(nbr, st));
            k_spread_w_rec<<<(nbr->n_ext + 255) / 256, 256, 0, st>>>(
                nbr->n, nbr->n_loc, nbr->n_ext, L.fprime, d_fprime_halo, nbr->perm.as<int>(),
                nbr->ghost_owner.as<int>(), m->zp->fe_over_B, -0.5, nbr->rec16.as<Rec16>());
        }
        TAB_LAUNCH_CHECK();
        prof_mark(2, st);
        int nblk = 0;
        EAMZ_FOR_L(eamz_lanes(), TAB_TRY(eamz_pass2_L<LL>(m, nbr, prec, L.fembed, d_eatom,
                                                          d_forces, d_own_mask, &nblk, st)));
        prof_mark(3, st);
        if (d_energy || d_virial) {
            k_reduce_partials<<<1, 256, 0, st>>>(nblk, nbr->partial.as<double>(), d_energy,
                                                 d_virial, nullptr);
            TAB_LAUNCH_CHECK();
        }
        prof_mark(4, st);
        if (g_prof_on && g_prof_calls < PROF_MAX_CALLS) ++g_prof_calls;
        return TAB_OK;
    }
    k_spread_w<<<(nbr->n_ext + 255) / 256, 256, 0, st>>>(
        nbr->n, nbr->n_loc, nbr->n_ext, L.fprime, d_fprime_halo, nbr->perm.as<int>(),
        nbr->ghost_owner.as<int>(), nbr->atoms.as<Atom4>());
    TAB_LAUNCH_CHECK();
    prof_mark(2, st);
    // pass 1 of the same evaluation left (g, g') per pair behind (positions unchanged
    // in between: tab_nbr_update / tab_nbr_build invalidate)
    if (FAST && sizeof(Real) == 8 && nbr->pcache_valid)
        k_eam_force<Real, FAST, FAST, false><<<L.nblk, EAM_T, L.smem, st>>>(
            nbr->n, nbr->atoms.as<Atom4>(), nbr->types_ext.as<uint8_t>(),
            nbr->counts.as<int>(), nbr->slice_ptr.as<uint32_t>(), nbr->col.as<uint32_t>(),
            nbr->perm.as<int>(), L.dev, L.z, L.fembed, d_eatom, d_forces,
            nbr->partial.as<double>(), nbr->pcache.as<double2>(), L.blk_first, d_own_mask, L.rcm);
    else if (!FAST && m->has_mlp_fn)
        k_eam_force<Real, false, false, true><<<L.nblk, EAM_T, L.smem, st>>>(
            nbr->n, nbr->atoms.as<Atom4>(), nbr->types_ext.as<uint8_t>(),
            nbr->counts.as<int>(), nbr->slice_ptr.as<uint32_t>(), nbr->col.as<uint32_t>(),
            nbr->perm.as<int>(), L.dev, L.z, L.fembed, d_eatom, d_forces,
            nbr->partial.as<double>(), nullptr, L.blk_first, d_own_mask, L.rcm);
    else
        k_eam_force<Real, FAST, false, false><<<L.nblk, EAM_T, L.smem, st>>>(
            nbr->n, nbr->atoms.as<Atom4>(), nbr->types_ext.as<uint8_t>(),
            nbr->counts.as<int>(), nbr->slice_ptr.as<uint32_t>(), nbr->col.as<uint32_t>(),
            nbr->perm.as<int>(), L.dev, L.z, L.fembed, d_eatom, d_forces,
            nbr->partial.as<double>(), nullptr, L.blk_first, d_own_mask, L.rcm);
    TAB_LAUNCH_CHECK();
    prof_mark(3, st);
    if (d_energy || d_virial) {
        k_reduce_partials<<<L.n_red, 256, 0, st>>>(L.nblk, nbr->partial.as<double>(),
                                                   d_energy, d_virial, L.struct_blk);
        TAB_LAUNCH_CHECK();
    }
    prof_mark(4, st);
    if (g_prof_on && g_prof_calls < PROF_MAX_CALLS) ++g_prof_calls;
    return TAB_OK;
}

template <typename Real>
static int adp_pass1(tab_model *m, tab_nbr *nbr, cudaStream_t st, bool recompute = false) {
    EamLaunch L;
    TAB_TRY(eam_prepare(m, nbr, false, L, st));
    if (nbr->n_halo > 0 && !recompute) {
        tab_set_error("ADP on lists with halo atoms: use tab_eam_eval_dd (inner-halo rows "
                      "recomputed, no moment exchange)");
        return TAB_EUNSUPPORTED;
    }
    const int nn = m->n_el * m->n_el, stride = m->n_el * 9;
    L.smem = (size_t)(4 * nn + m->n_el) * sizeof(tab_fn);
    TAB_TRY(nbr->adp.ensure(sizeof(double) * (size_t)nbr->n_ext * stride));
    if (nbr->n_halo > 0)     // outer-halo moments are never computed: keep them finite
        TAB_CUDA(cudaMemsetAsync(nbr->adp.p, 0, sizeof(double) * (size_t)nbr->n_ext * stride, st));
    prof_mark(0, st);
    auto kr = m->has_mlp_fn ? k_adp_rho<Real, true> : k_adp_rho<Real, false>;
    kr<<<L.nblk, EAM_T, L.smem, st>>>(
        nbr->n, nbr->atoms.as<Atom4>(), nbr->types_ext.as<uint8_t>(),
        nbr->counts.as<int>(), nbr->slice_ptr.as<uint32_t>(), nbr->col.as<uint32_t>(),
        L.dev, L.fprime, L.fembed, nbr->adp.as<double>(), L.blk_first, L.rcm);
    TAB_LAUNCH_CHECK();
    if (nbr->n_ext > nbr->n_loc) {
        k_adp_spread<<<(nbr->n_ext - nbr->n_loc + 255) / 256, 256, 0, st>>>(
            nbr->n_loc, nbr->n_ext, stride, nbr->ghost_owner.as<int>(), nbr->adp.as<double>());
        TAB_LAUNCH_CHECK();
    }
    prof_mark(1, st);
    return TAB_OK;
}

template <typename Real>
static int adp_pass2(tab_model *m, tab_nbr *nbr, double *d_energy, double *d_eatom,
                     double *d_forces, double *d_virial, cudaStream_t st,
                     const int *d_own_mask = nullptr) {
    EamLaunch L;
    TAB_TRY(eam_prepare(m, nbr, false, L, st));
    const int nn = m->n_el * m->n_el;
    L.smem = (size_t)(4 * nn + m->n_el) * sizeof(tab_fn);
    k_spread_w<<<(nbr->n_ext + 255) / 256, 256, 0, st>>>(
        nbr->n, nbr->n_loc, nbr->n_ext, L.fprime, nullptr, nbr->perm.as<int>(),
        nbr->ghost_owner.as<int>(), nbr->atoms.as<Atom4>());
    TAB_LAUNCH_CHECK();
    prof_mark(2, st);
    auto kf = m->has_mlp_fn ? k_adp_force<Real, true> : k_adp_force<Real, false>;
    kf<<<L.nblk, EAM_T, L.smem, st>>>(
        nbr->n, nbr->atoms.as<Atom4>(), nbr->types_ext.as<uint8_t>(),
        nbr->counts.as<int>(), nbr->slice_ptr.as<uint32_t>(), nbr->col.as<uint32_t>(),
        nbr->perm.as<int>(), L.dev, nbr->adp.as<double>(), L.fembed, d_eatom, d_forces,
        nbr->partial.as<double>(), L.blk_first, d_own_mask, L.rcm);
    TAB_LAUNCH_CHECK();
    prof_mark(3, st);
    if (d_energy || d_virial) {
        k_reduce_partials<<<L.n_red, 256, 0, st>>>(L.nblk, nbr->partial.as<double>(),
                                                   d_energy, d_virial, L.struct_blk);
        TAB_LAUNCH_CHECK();
    }
    prof_mark(4, st);
    if (g_prof_on && g_prof_calls < PROF_MAX_CALLS) ++g_prof_calls;
    return TAB_OK;
}

#define EAM_DISPATCH(FN, ...)                                                        \
    do {                                                                             \
        if (precision == TAB_PRECISION_HIGH)                                         \
            return zhou1_f64_ok(m, nbr) ? FN<double, true>(__VA_ARGS__)              \
                                        : FN<double, false>(__VA_ARGS__);            \
        if (precision == TAB_PRECISION_MEDIUM)                                       \
            return m->zhou1 ? FN<float, true>(__VA_ARGS__) : FN<float, false>(__VA_ARGS__);   \
        tab_set_error("unknown precision %d", precision);                            \
        return TAB_EINVAL;                                                           \
    } while (0)

static int check_handles(tab_model *m, tab_nbr *nbr, const char *who) {
    if (!m || !nbr) {
        tab_set_error("%s: null handle", who);
        return TAB_EINVAL;
    }
    if (!nbr->built) {
        tab_set_error("%s before tab_nbr_build", who);
        return TAB_ESTATE;
    }
    return TAB_OK;
}

extern "C" int tab_eam_pass1(tab_model *m, tab_nbr *nbr, int32_t precision,
                             double *d_fprime, void *stream) {
    TAB_TRY(check_handles(m, nbr, "tab_eam_pass1"));
    cudaStream_t st = (cudaStream_t)stream;
    if (m->kind == TAB_EAM_ADP) {
        if (precision == TAB_PRECISION_HIGH) return adp_pass1<double>(m, nbr, st);
        return adp_pass1<float>(m, nbr, st);
    }
    EAM_DISPATCH(eam_pass1, m, nbr, d_fprime, st);
}

extern "C" int tab_eam_pass2(tab_model *m, tab_nbr *nbr, int32_t precision,
                             const double *d_fprime_halo, double *d_energy,
                             double *d_eatom, double *d_forces, double *d_virial,
                             void *stream) {
    TAB_TRY(check_handles(m, nbr, "tab_eam_pass2"));
    cudaStream_t st = (cudaStream_t)stream;
    if (m->kind == TAB_EAM_ADP) {
        if (precision == TAB_PRECISION_HIGH)
            return adp_pass2<double>(m, nbr, d_energy, d_eatom, d_forces, d_virial, st);
        return adp_pass2<float>(m, nbr, d_energy, d_eatom, d_forces, d_virial, st);
    }
    EAM_DISPATCH(eam_pass2, m, nbr, d_fprime_halo, d_energy, d_eatom, d_forces, d_virial, st);
}

extern "C" int tab_eam_eval(tab_model *m, tab_nbr *nbr, int32_t precision,
                            double *d_energy, double *d_eatom, double *d_forces,
                            double *d_virial, void *stream) {
    TAB_TRY(check_handles(m, nbr, "tab_eam_eval"));
    if (nbr->n_halo > 0) {
        tab_set_error("tab_eam_eval: lists hold halo atoms; use tab_eam_pass1/pass2");
        return TAB_ESTATE;
    }
    TAB_TRY(tab_eam_pass1(m, nbr, precision, nullptr, stream));
    return tab_eam_pass2(m, nbr, precision, nullptr, d_energy, d_eatom, d_forces,
                         d_virial, stream);
}

// Spatial decomposition WITHOUT a second exchange (see include/tab200.h): lists over
// [own | inner halo] + outer halo, inner-halo rows recomputed, own_mask keeps them out of
// the rank's energy / virial sums.  Serves ADP (whose F' + moment exchange is not built)
// and is an alternative to tab_eam_pass1/pass2 for EAM / FS.
extern "C" int tab_eam_eval_dd(tab_model *m, tab_nbr *nbr, int32_t precision,
                               const int32_t *d_mask, double *d_energy, double *d_eatom,
                               double *d_forces, double *d_virial, void *stream) {
    TAB_TRY(check_handles(m, nbr, "tab_eam_eval_dd"));
    if (!d_mask) {
        tab_set_error("tab_eam_eval_dd: null mask");
        return TAB_EINVAL;
    }
    if (nbr->n_struct > 0) {
        tab_set_error("tab_eam_eval_dd: batch handles are not decomposed");
        return TAB_EUNSUPPORTED;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (m->kind == TAB_EAM_ADP) {
        if (precision == TAB_PRECISION_HIGH) {
            TAB_TRY(adp_pass1<double>(m, nbr, st, true));
            return adp_pass2<double>(m, nbr, d_energy, d_eatom, d_forces, d_virial, st, d_mask);
        }
        TAB_TRY(adp_pass1<float>(m, nbr, st, true));
        return adp_pass2<float>(m, nbr, d_energy, d_eatom, d_forces, d_virial, st, d_mask);
    }
    TAB_TRY(tab_eam_pass1(m, nbr, precision, nullptr, stream));
    EAM_DISPATCH(eam_pass2, m, nbr, nullptr, d_energy, d_eatom, d_forces, d_virial, st, d_mask);
}

// Tabulation of one function of the model on a caller-supplied grid (export_to_setfl):
// which = 0 rho, 1 phi, 2 embed, 3 dipole, 4 quadrupole; index = a * n_el + b (centre a,
// neighbour b) for the pair functions, the element for the embedding.  float64.
template <bool NN>
__global__ void k_eam_tabulate(EamDev m, const tab_fn *__restrict__ fnp, int is_embed, int n,
                               const double *__restrict__ x, double *__restrict__ y,
                               double *__restrict__ dy) {
    __shared__ tab_fn fn;
    if (threadIdx.x == 0) fn = *fnp;
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double f, df;
    if (is_embed) eval_embed_fn<double, NN>(fn, x[i], f, df, m.pool);
    else eval_pair_fn<double, NN>(fn, x[i], f, df, m.pool);
    y[i] = f;
    if (dy) dy[i] = df;
}

extern "C" int tab_eam_tabulate(tab_model *m, int32_t which, int32_t index, int32_t n,
                                const double *d_x, double *d_y, double *d_dy, void *stream) {
    if (!m || n <= 0 || !d_x || !d_y) {
        tab_set_error("tab_eam_tabulate: bad argument");
        return TAB_EINVAL;
    }
    const int nn = m->n_el * m->n_el;
    const bool adp = m->kind == TAB_EAM_ADP;
    int off, count;
    switch (which) {
    case 0: off = 0; count = nn; break;
    case 1: off = nn; count = nn; break;
    case 2: off = 2 * nn; count = m->n_el; break;
    case 3: off = 2 * nn + m->n_el; count = adp ? nn : 0; break;
    case 4: off = 3 * nn + m->n_el; count = adp ? nn : 0; break;
    default: off = 0; count = 0;
    }
    if (index < 0 || index >= count) {
        tab_set_error("tab_eam_tabulate: no function %d of kind %d in this model", index, which);
        return TAB_EINVAL;
    }
    EamDev dev;
    dev.kind = m->kind;
    dev.n_el = m->n_el;
    dev.rho = m->tables.as<tab_fn>();
    dev.phi = dev.rho + nn;
    dev.embed = dev.rho + 2 * nn;
    dev.pool = m->pool.as<double>();
    cudaStream_t st = (cudaStream_t)stream;
    const tab_fn *fnp = m->tables.as<tab_fn>() + off + index;
    if (m->has_mlp_fn)
        k_eam_tabulate<true><<<(n + 127) / 128, 128, 0, st>>>(dev, fnp, which == 2, n, d_x, d_y,
                                                              d_dy);
    else
        k_eam_tabulate<false><<<(n + 127) / 128, 128, 0, st>>>(dev, fnp, which == 2, n, d_x,
                                                               d_y, d_dy);
    TAB_LAUNCH_CHECK();
    return TAB_OK;
}

// Host-buffer convenience: see include/tab200.h.
struct HostStage {
    DevBuf pos, types, out;
};
static HostStage g_stage;

extern "C" int tab_eam_compute_host(tab_model *m, tab_nbr *nbr, int32_t precision,
                                    int32_t n, const double *h_pos,
                                    const int32_t *h_types, const double *h_cell,
                                    const int32_t *h_pbc, double rc, int32_t rebuild,
                                    double *h_energy, double *h_eatom,
                                    double *h_forces, double *h_virial, void *stream) {
    if (!m || !nbr || n <= 0 || !h_pos || !h_cell || !h_pbc) {
        tab_set_error("tab_eam_compute_host: bad argument");
        return TAB_EINVAL;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (nbr->n_struct > 0 && !rebuild) {
        tab_set_error("tab_eam_compute_host: the handle holds a batch; pass rebuild != 0");
        return TAB_ESTATE;
    }
    TAB_TRY(g_stage.pos.ensure(sizeof(double) * 3 * (size_t)n));
    TAB_TRY(g_stage.types.ensure(sizeof(int32_t) * (size_t)n));
    // out: energy[1] virial[9] (pad to 16) | forces[3n] | eatom[n]
    TAB_TRY(g_stage.out.ensure(sizeof(double) * (16 + 4 * (size_t)n)));
    double *d_pos = g_stage.pos.as<double>();
    int32_t *d_types = h_types ? g_stage.types.as<int32_t>() : nullptr;
    double *d_out = g_stage.out.as<double>();
    TAB_CUDA(cudaMemcpyAsync(d_pos, h_pos, sizeof(double) * 3 * (size_t)n,
                             cudaMemcpyHostToDevice, st));
    if (rebuild || !nbr->built || nbr->n != n) {
        if (h_types)
            TAB_CUDA(cudaMemcpyAsync(d_types, h_types, sizeof(int32_t) * (size_t)n,
                                     cudaMemcpyHostToDevice, st));
        TAB_TRY(tab_nbr_build(nbr, n, d_pos, d_types, h_cell, h_pbc, rc, stream));
    } else {
        TAB_TRY(tab_nbr_update(nbr, d_pos, h_cell, stream));
    }
    TAB_TRY(tab_eam_eval(m, nbr, precision, d_out, h_eatom ? d_out + 16 + 3 * (size_t)n : nullptr,
                         h_forces ? d_out + 16 : nullptr, d_out + 1, stream));
    if (h_energy)
        TAB_CUDA(cudaMemcpyAsync(h_energy, d_out, sizeof(double), cudaMemcpyDeviceToHost, st));
    if (h_virial)
        TAB_CUDA(cudaMemcpyAsync(h_virial, d_out + 1, 9 * sizeof(double),
                                 cudaMemcpyDeviceToHost, st));
    if (h_forces)
        TAB_CUDA(cudaMemcpyAsync(h_forces, d_out + 16, sizeof(double) * 3 * (size_t)n,
                                 cudaMemcpyDeviceToHost, st));
    if (h_eatom)
        TAB_CUDA(cudaMemcpyAsync(h_eatom, d_out + 16 + 3 * (size_t)n,
                                 sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, st));
    TAB_CUDA(cudaStreamSynchronize(st));
    return TAB_OK;
}
