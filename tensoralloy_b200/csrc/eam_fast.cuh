// eam_fast.cuh -- the single-element zjw04 EAM passes (BASELINE config 3: 1 M-atom Ni) over
// LANE-SPLIT rows.  Included by eam.cu.
//
// Same arithmetic as k_eam_rho / k_eam_force (reference: zjw04.py:187-389, generic.py:102-117,
// eam.py:265-570, basic.py:276-331), different work layout:
//
//   * L lanes of a warp share ONE atom's row (entry k is handled by lane k % L); a warp
//     covers 32 / L consecutive atoms.  The L1 data stage serves a gather in one wavefront per
//     distinct 128-byte line: with one atom per lane the 32 neighbours of a step sit in
//     ~21.5 lines (measured 22.5 wavefronts per gather, profiles/r01k); the L consecutive entries
//     of one row are neighbours in memory, L = 4 brings a gather down to ~14.6 lines, L = 8 to
//     ~13.4 (tools/sim_gather_lines.py reproduces the measured figure and these).
//   * virial in the form  W = - sum_i F_i^real (x) R_i + 1/2 sum_{p: j image/halo} g_p (x) D_p
//     (identity for symmetric pair gradients, derivation in DESIGN.md): the six per-pair FMAs
//     of g (x) D are only spent on pairs whose neighbour is a periodic image or a halo atom.
//   * r < rc mask: lists built with a skin hold entries beyond the model's cutoff; they
//     contribute exactly 0, so a reused list equals a fresh one (transformer/universal.py:58
//     rebuilds per call).
//   * float32 ('medium'): 16-byte fixed-point records (tab_internal.h Rec16), geometry, functions
//     and per-atom sums in float32, block sums in float64.
#pragma once

#ifndef EAMZ_T
#define EAMZ_T 128
#endif
#ifndef EAMZ_L
#define EAMZ_L 4          // lanes per atom (template parameter of the kernels; A/B: profiles/r02*)
#endif
#ifndef EAMZ_MINB_RHO
#define EAMZ_MINB_RHO 6
#endif
#ifndef EAMZ_MINB_FORCE
#define EAMZ_MINB_FORCE 5
#endif
#ifndef EAMZ_VIR
#define EAMZ_VIR 1        // 1: F (x) R form of the virial, 0: six FMAs on every pair
#endif

// ---------------------------------------------------------------------------
// lane-split copy of the lists: group g = atoms [g G, (g + 1) G), G = 32 / L; entry k of the
// atom with local index al at  ls_col[(ls_ptr[g] + k / L) * 32 + al * L + k % L]
// ---------------------------------------------------------------------------
template <int L>
__global__ void k_ls_widths(int n, const int *__restrict__ counts,
                            uint32_t *__restrict__ widths) {
    constexpr int G = 32 / L;
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g * G >= n) return;
    int mx = 0;
    for (int a = g * G; a < min(n, (g + 1) * G); ++a) mx = max(mx, counts[a]);
    widths[g] = (uint32_t)((mx + L - 1) / L);
}

// Every lane of a group runs the group's width: the slots past an atom's count (and the
// columns of the atoms missing from the last group) hold the index of the SENTINEL record
// (extended index n_ext, a point far outside the frame), which the cutoff mask turns into
// exact zeros.  No per-lane trip counts, no predicated loads in the pair loops.
template <int L>
__global__ void __launch_bounds__(128)
k_ls_fill(int n, int n_pad, uint32_t sentinel, const int *__restrict__ counts,
          const uint32_t *__restrict__ slice_ptr, const uint32_t *__restrict__ col,
          const uint32_t *__restrict__ ls_ptr, uint32_t *__restrict__ ls_col,
          unsigned long long total_rows) {
    constexpr int G = 32 / L;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_pad) {
        // the two rows after the last group (the loops prefetch two rows ahead)
        const int k = idx - n_pad;
        if (k < 64) ls_col[total_rows * 32u + k] = sentinel;
        return;
    }
    const int g = idx / G;
    const uint32_t w = ls_ptr[g + 1] - ls_ptr[g];
    uint32_t *dst = ls_col + ((size_t)ls_ptr[g] * 32u + (idx % G) * L);
    int cnt = 0;
    if (idx < n) {
        const uint32_t *src = col + ((size_t)slice_ptr[idx >> 5] * 32u + (idx & 31));
        cnt = counts[idx];
        for (int k = 0; k < cnt; ++k)
            dst[(size_t)(k / L) * 32u + (k % L)] = src[(size_t)k * 32u];
    }
    for (int k = cnt; k < (int)(w * L); ++k) dst[(size_t)(k / L) * 32u + (k % L)] = sentinel;
}

template <int L>
static int ensure_lanesplit(tab_nbr *nbr, cudaStream_t st) {
    if (nbr->ls_L == L) return TAB_OK;
    constexpr int G = 32 / L;
    const int n = nbr->n, n_groups = (n + G - 1) / G;
    TAB_TRY(nbr->ls_ptr.ensure(sizeof(uint32_t) * (size_t)(n_groups + 2)));
    TAB_TRY(nbr->stats.ensure(5 * sizeof(unsigned long long)));
    // widths of n_groups groups + one zero: the scan then leaves ls_ptr[n_groups] = total
    TAB_CUDA(cudaMemsetAsync(nbr->ls_ptr.as<uint32_t>() + n_groups, 0, sizeof(uint32_t), st));
    k_ls_widths<L><<<(n_groups + 127) / 128, 128, 0, st>>>(n, nbr->counts.as<int>(),
                                                           nbr->ls_ptr.as<uint32_t>());
    TAB_LAUNCH_CHECK();
    unsigned long long *d_total = nbr->stats.as<unsigned long long>() + 2;
    TAB_TRY(tab_scan_exclusive_u32(nbr->ls_ptr.as<uint32_t>(), nbr->ls_ptr.as<uint32_t>(),
                                   n_groups + 1, d_total, nbr->scan_tmp, st));
    unsigned long long total = 0;
    TAB_CUDA(cudaMemcpyAsync(&total, d_total, sizeof(total), cudaMemcpyDeviceToHost, st));
    TAB_CUDA(cudaStreamSynchronize(st));
    nbr->ls_rows = (long long)total;
    TAB_TRY(nbr->ls_col.ensure(sizeof(uint32_t) * 32 * (size_t)(total + 2)));
    const int n_pad = n_groups * G;
    k_ls_fill<L><<<(n_pad + 64 + 127) / 128, 128, 0, st>>>(
        n, n_pad, (uint32_t)nbr->n_ext, nbr->counts.as<int>(), nbr->slice_ptr.as<uint32_t>(),
        nbr->col.as<uint32_t>(), nbr->ls_ptr.as<uint32_t>(), nbr->ls_col.as<uint32_t>(), total);
    TAB_LAUNCH_CHECK();
    nbr->ls_L = L;
    return TAB_OK;
}

// lane -> (atom, part of its row)
template <int L>
struct LaneRow {
    int idx, part, steps;      // steps = width of the group: the same for the whole warp
    bool active;
    const uint32_t *cp;
    __device__ __forceinline__ LaneRow(int n, const uint32_t *__restrict__ ls_ptr,
                                       const uint32_t *__restrict__ ls_col) {
        constexpr int G = 32 / L;
        const int lane = threadIdx.x & 31;
        const int g = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
        idx = g * G + lane / L;
        part = lane % L;
        active = idx < n;
        uint32_t p0 = 0, p1 = 0;
        if (g * G < n) {
            p0 = ls_ptr[g];
            p1 = ls_ptr[g + 1];
        }
        steps = (int)(p1 - p0);
        cp = ls_col + ((size_t)p0 * 32u + lane);
    }
};

// sum over the L lanes of an atom
template <int L, typename T>
__device__ __forceinline__ T lanes_sum(T v) {
#pragma unroll
    for (int d = 1; d < L; d <<= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

// block sums of (E, vxx, vyy, vzz, vyz, vxz, vxy) -> partial[blockIdx.x * 8 + q]
__device__ __forceinline__ void block_partials(double (&acc)[7], double *__restrict__ partial) {
    __shared__ double red[EAMZ_T / 32][7];
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
#pragma unroll
        for (int w = 0; w < EAMZ_T / 32; ++w) v += red[w][threadIdx.x];
        partial[(size_t)blockIdx.x * 8 + threadIdx.x] = v;
    }
}

// ---------------------------------------------------------------------------
// float64: one zhou term  f(r) = a exp(-b (r/re - 1)) / (1 + (r/re - kappa)^20) with the
// exponent in units of ln2 / 64:  y = 64 log2(e) (-b r/re + b + ln a)  (one FMA from r),
//   n = rint(y), exp = 2^(n >> 6) * 2^((n & 63) / 64) * P(y - n),  P = exp((ln2 / 64) f), |f| <= 1/2
// (degree 5: truncation 4e-17).  The rounding of y (|y| < 2^11) costs 1.2e-15 relative.
// 23 FP64 instructions for f and df/dr, 17 for f alone.
// ---------------------------------------------------------------------------
struct ZT2 {
    double yr, yc;        // y = yr * r + yc
    double ire, nkappa;   // u = r / re - kappa = ire * r + nkappa
    double c20;           // 20 / re
    double nb_re;         // -b / re
};

struct ZPair {
    ZT2 rho, a, b;        // fe-term (pass 1), A-term, B-term of phi = A-term - B-term
    double fe_over_B;
    // float32 mirror (exponents in units of ln2: ex2.approx); x = r / re
    float f_yx_rho, f_yc_rho, f_k_rho;
    float f_yx_a, f_yc_a, f_k_a, f_nbre_a;
    float f_yx_b, f_yc_b, f_k_b, f_nbre_b;
    float f_fe_over_B, f_20_re;
};

static void zt2_fold(ZT2 &t, double a, double b, double c, double re) {
    const double S = 64.0 * 1.4426950408889634074;
    t.yr = -b * S / re;
    t.yc = (b + log(a)) * S;
    t.ire = 1.0 / re;
    t.nkappa = -c;
    t.c20 = 20.0 / re;
    t.nb_re = -b / re;
}

static void zpair_fold(ZPair &z, const Zhou1 &p) {
    const double re = 1.0 / p.re;
    zt2_fold(z.rho, p.fe, p.beta, p.lamda, re);
    zt2_fold(z.a, p.A, p.alpha, p.kappa, re);
    zt2_fold(z.b, p.B, p.beta, p.lamda, re);
    z.fe_over_B = p.fe / p.B;
    const double l2e = 1.4426950408889634074;
    z.f_yx_rho = (float)(-p.beta * l2e);
    z.f_yc_rho = (float)((p.beta + log(p.fe)) * l2e);
    z.f_k_rho = (float)p.lamda;
    z.f_yx_a = (float)(-p.alpha * l2e);
    z.f_yc_a = (float)((p.alpha + log(p.A)) * l2e);
    z.f_k_a = (float)p.kappa;
    z.f_nbre_a = (float)(-p.alpha / re);
    z.f_yx_b = (float)(-p.beta * l2e);
    z.f_yc_b = (float)((p.beta + log(p.B)) * l2e);
    z.f_k_b = (float)p.lamda;
    z.f_nbre_b = (float)(-p.beta / re);
    z.f_fe_over_B = (float)(p.fe / p.B);
    z.f_20_re = (float)(20.0 / re);
}

__device__ __forceinline__ double zexp64(double y, const double *__restrict__ etab) {
    const double SHIFT = 6755399441055744.0;   // 1.5 * 2^52
    const double zz = y + SHIFT;
    const int n = __double2loint(zz);
    const double fr = y - (zz - SHIFT);        // exact, |fr| <= 1/2
    // exp(c fr), c = ln2 / 64
    double e = 1.2676118923200506e-12;               // c^5 / 120
    e = fma(e, fr, 5.8521082468723563e-10);          // c^4 / 24
    e = fma(e, fr, 2.1614469394785756e-07);          // c^3 / 6
    e = fma(e, fr, 5.8652907259685552e-05);          // c^2 / 2
    e = fma(e, fr, 1.0830424696249145e-02);          // c
    e = fma(e, fr, 1.0);
    e *= etab[n & 63];
    return __hiloint2double(__double2hiint(e) + ((n >> 6) << 20), __double2loint(e));
}

template <bool DERIV>
__device__ __forceinline__ void zt2_eval(double r, const ZT2 &p,
                                         const double *__restrict__ etab, double &f,
                                         double &df) {
    const double u = fma(r, p.ire, p.nkappa);
    const double u2 = u * u, u4 = u2 * u2, u8 = u4 * u4, u16 = u8 * u8;
    const double q = tab_rcp(fma(u16, u4, 1.0));
    const double e = zexp64(fma(r, p.yr, p.yc), etab);
    f = e * q;
    if (DERIV) {
        // df/dr = f (-(20 / re) u^19 q - b / re)
        const double u19p = (u16 * u2) * (u * p.c20);
        df = f * fma(-u19p, q, p.nb_re);
    }
}

__device__ __forceinline__ void load_exp2_tab64(double *s_tab) {
    if (threadIdx.x < 64) s_tab[threadIdx.x] = c_exp2_tab[threadIdx.x];
    __syncthreads();
}

// ---------------------------------------------------------------------------
// pass 1, float64
// ---------------------------------------------------------------------------
template <int L>
__global__ void __launch_bounds__(EAMZ_T, EAMZ_MINB_RHO)
k_eamz_rho(int n, const Atom4 *__restrict__ atoms, const uint32_t *__restrict__ ls_ptr,
           const uint32_t *__restrict__ ls_col, const int *__restrict__ perm, ZPair z,
           double rc2m, tab_fn embed0, double *__restrict__ fprime,
           double *__restrict__ fembed, double *__restrict__ fprime_caller) {
    __shared__ double s_etab[64];
    load_exp2_tab64(s_etab);
    const LaneRow<L> row(n, ls_ptr, ls_col);
    double rho = 0.0;
    if (row.steps > 0) {
        Atom4 me;
        me.x = me.y = me.z = 0.0;
        if (row.active) me = atoms[row.idx];
        uint32_t c1 = row.cp[32];
        Atom4 a = atoms[row.cp[0] & TAB_COL_IDX_MASK];
#pragma unroll 2
        for (int t = 0; t < row.steps; ++t) {
            // the records of step t + 1 and the indices of step t + 2 are in flight (the two
            // rows after a group are readable: next group or the sentinel rows)
            const Atom4 an = atoms[c1 & TAB_COL_IDX_MASK];
            c1 = row.cp[(size_t)(t + 2) * 32u];
            const double dx = a.x - me.x, dy = a.y - me.y, dz = a.z - me.z;
            const double s = fma(dx, dx, fma(dy, dy, fma(dz, dz, 1e-14)));
            const double r = s * tab_rsqrt(s);
            double f, df;
            zt2_eval<false>(r, z.rho, s_etab, f, df);
            rho += s < rc2m ? f : 0.0;
            a = an;
        }
    }
    rho = lanes_sum<L>(rho);
    if (row.active && row.part == 0) {
        double F, dF;
        zhou_embed<double>(embed0.p, embed0.kind == TAB_FN_ZHOU_EMBED_XC, rho, F, dF);
        fprime[row.idx] = dF;
        fembed[row.idx] = F;
        if (fprime_caller) fprime_caller[perm[row.idx]] = dF;
    }
}

// ---------------------------------------------------------------------------
// pass 2, float64.  Atom4.w holds  w = F'(rho) fe / B - 1/2  (k_spread_w with scale and shift),
// so that  dE/dr = (F'_i + F'_j) rho'(r) + phi'(r) = (w_i + w_j) B-term' + A-term'.
// n_real: list entries below it are own, non-image atoms.
// ---------------------------------------------------------------------------
template <int L, int VIR>
__global__ void __launch_bounds__(EAMZ_T, EAMZ_MINB_FORCE)
k_eamz_force(int n, int n_real, const Atom4 *__restrict__ atoms,
             const uint32_t *__restrict__ ls_ptr, const uint32_t *__restrict__ ls_col,
             const int *__restrict__ perm, ZPair z, double rc2m,
             const double *__restrict__ fembed, double *__restrict__ eatom,
             double *__restrict__ forces, double *__restrict__ partial,
             const int *__restrict__ own_mask) {
    __shared__ double s_etab[64];
    // VIR == 1: image / halo pairs are rare; their sums live in shared memory, not registers
    __shared__ double s_gh[VIR ? 9 : 1][EAMZ_T];
    load_exp2_tab64(s_etab);
    const LaneRow<L> row(n, ls_ptr, ls_col);
    double fx = 0, fy = 0, fz = 0, ep = 0;
    double vxx = 0, vyy = 0, vzz = 0, vyz = 0, vxz = 0, vxy = 0;
    if (VIR) {
#pragma unroll
        for (int q = 0; q < 9; ++q) s_gh[q][threadIdx.x] = 0.0;
    }
    Atom4 me;
    me.x = me.y = me.z = me.w = 0.0;
    if (row.active) me = atoms[row.idx];
    if (row.steps > 0) {
        uint32_t c0 = row.cp[0];
        uint32_t c1 = row.cp[32];
        Atom4 a = atoms[c0 & TAB_COL_IDX_MASK];
#pragma unroll 2
        for (int t = 0; t < row.steps; ++t) {
            const Atom4 an = atoms[c1 & TAB_COL_IDX_MASK];
            const uint32_t cc = c0;
            c0 = c1;
            c1 = row.cp[(size_t)(t + 2) * 32u];
            const double dx = a.x - me.x, dy = a.y - me.y, dz = a.z - me.z;
            const double s = fma(dx, dx, fma(dy, dy, fma(dz, dz, 1e-14)));
            const double rinv = tab_rsqrt(s);
            const double r = s * rinv;
            double ga, dga, gb, dgb;
            zt2_eval<true>(r, z.a, s_etab, ga, dga);
            zt2_eval<true>(r, z.b, s_etab, gb, dgb);
            const bool in = s < rc2m;
            const double der = fma(me.w + a.w, dgb, dga);
            const double sc = in ? der * rinv : 0.0;
            ep += in ? ga - gb : 0.0;
            if (VIR) {
                fx = fma(sc, dx, fx);
                fy = fma(sc, dy, fy);
                fz = fma(sc, dz, fz);
                if (in && (int)(cc & TAB_COL_IDX_MASK) >= n_real) {
                    const double gx = sc * dx, gy = sc * dy, gz = sc * dz;
                    volatile double *g = &s_gh[0][threadIdx.x];   // keep the sums OUT of registers
                    g[0 * EAMZ_T] += gx;
                    g[1 * EAMZ_T] += gy;
                    g[2 * EAMZ_T] += gz;
                    g[3 * EAMZ_T] = fma(gx, dx, g[3 * EAMZ_T]);
                    g[4 * EAMZ_T] = fma(gy, dy, g[4 * EAMZ_T]);
                    g[5 * EAMZ_T] = fma(gz, dz, g[5 * EAMZ_T]);
                    g[6 * EAMZ_T] = fma(gy, dz, g[6 * EAMZ_T]);
                    g[7 * EAMZ_T] = fma(gx, dz, g[7 * EAMZ_T]);
                    g[8 * EAMZ_T] = fma(gx, dy, g[8 * EAMZ_T]);
                }
            } else {
                const double gx = sc * dx, gy = sc * dy, gz = sc * dz;
                fx += gx;
                fy += gy;
                fz += gz;
                vxx = fma(gx, dx, vxx);
                vyy = fma(gy, dy, vyy);
                vzz = fma(gz, dz, vzz);
                vyz = fma(gy, dz, vyz);
                vxz = fma(gx, dz, vxz);
                vxy = fma(gx, dy, vxy);
            }
            a = an;
        }
    }
    double acc[7];
    if (VIR) {
        // this lane's share:  -(f - f_gh) (x) R_i + 1/2 sum_gh g (x) D   (symmetric part)
        const double *g = &s_gh[0][threadIdx.x];
        const double rx = fx - g[0 * EAMZ_T], ry = fy - g[1 * EAMZ_T], rz = fz - g[2 * EAMZ_T];
        acc[1] = fma(-rx, me.x, 0.5 * g[3 * EAMZ_T]);
        acc[2] = fma(-ry, me.y, 0.5 * g[4 * EAMZ_T]);
        acc[3] = fma(-rz, me.z, 0.5 * g[5 * EAMZ_T]);
        acc[4] = 0.5 * (g[6 * EAMZ_T] - fma(ry, me.z, rz * me.y));
        acc[5] = 0.5 * (g[7 * EAMZ_T] - fma(rx, me.z, rz * me.x));
        acc[6] = 0.5 * (g[8 * EAMZ_T] - fma(rx, me.y, ry * me.x));
    } else {
        acc[1] = 0.5 * vxx;
        acc[2] = 0.5 * vyy;
        acc[3] = 0.5 * vzz;
        acc[4] = 0.5 * vyz;
        acc[5] = 0.5 * vxz;
        acc[6] = 0.5 * vxy;
    }
    fx = lanes_sum<L>(fx);
    fy = lanes_sum<L>(fy);
    fz = lanes_sum<L>(fz);
    ep = lanes_sum<L>(ep);
    acc[0] = 0.0;
    bool mine = true;
    if (row.active) {
        const int o = perm[row.idx];
        if (own_mask) mine = own_mask[o] != 0;
        if (row.part == 0) {
            const double ei = fembed[row.idx] + 0.5 * ep;
            if (eatom) eatom[o] = ei;
            if (forces) {
                forces[3 * (size_t)o + 0] = fx;
                forces[3 * (size_t)o + 1] = fy;
                forces[3 * (size_t)o + 2] = fz;
            }
            acc[0] = ei;
        }
    }
    if (!mine || !row.active)
#pragma unroll
        for (int q = 0; q < 7; ++q) acc[q] = 0.0;
    block_partials(acc, partial);
}

// ---------------------------------------------------------------------------
// float32 ('medium'): fixed-point records, lengths in units of delta.
//   s_q = |D_q|^2 + eps / delta^2,  r = delta sqrt(s_q),  x = r / re
// Forces need no delta: (dE/dr / r) D = (dE/dr / r_q) D_q.
// ---------------------------------------------------------------------------
struct QScale {
    float x_per_q;     // delta / re
    float eps_q;       // 1e-8 / delta^2      (precision.py:114)
    float rc2_q;       // mask radius^2 in delta^2
    float pad_;
    double ox, oy, oz, ddelta;
};

__device__ __forceinline__ float f_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float f_rcp(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float f_rsqrt(float x) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <int L>
__global__ void __launch_bounds__(EAMZ_T, 8)
k_eamz_rho_f32(int n, const Rec16 *__restrict__ recs, const uint32_t *__restrict__ ls_ptr,
               const uint32_t *__restrict__ ls_col, const int *__restrict__ perm, ZPair z,
               QScale qs, tab_fn embed0, double *__restrict__ fprime,
               double *__restrict__ fembed, double *__restrict__ fprime_caller) {
    const LaneRow<L> row(n, ls_ptr, ls_col);
    float rho = 0.f;
    if (row.steps > 0) {
        Rec16 me;
        me.qx = me.qy = me.qz = 0;
        if (row.active) me = recs[row.idx];
        uint32_t c1 = row.cp[32];
        Rec16 a = recs[row.cp[0] & TAB_COL_IDX_MASK];
#pragma unroll 2
        for (int t = 0; t < row.steps; ++t) {
            const Rec16 an = recs[c1 & TAB_COL_IDX_MASK];
            c1 = row.cp[(size_t)(t + 2) * 32u];
            const float dx = (float)(a.qx - me.qx), dy = (float)(a.qy - me.qy),
                        dz = (float)(a.qz - me.qz);
            const float s = fmaf(dx, dx, fmaf(dy, dy, fmaf(dz, dz, qs.eps_q)));
            const float x = (s * f_rsqrt(s)) * qs.x_per_q;
            const float u = x - z.f_k_rho;
            const float u2 = u * u, u4 = u2 * u2, u8 = u4 * u4, u16 = u8 * u8;
            const float f = f_ex2(fmaf(x, z.f_yx_rho, z.f_yc_rho)) * f_rcp(fmaf(u16, u4, 1.f));
            rho += s < qs.rc2_q ? f : 0.f;
            a = an;
        }
    }
    rho = lanes_sum<L>(rho);
    if (row.active && row.part == 0) {
        float F, dF;
        zhou_embed<float>(embed0.p, embed0.kind == TAB_FN_ZHOU_EMBED_XC, rho, F, dF);
        fprime[row.idx] = (double)dF;
        fembed[row.idx] = (double)F;
        if (fprime_caller) fprime_caller[perm[row.idx]] = (double)dF;
    }
}

// Rec16.w = F'(rho) fe / B - 1/2 as in the float64 kernel.  The F (x) R form of the virial takes
// the positions relative to the block's first atom (float32 force sums times 200 A coordinates
// would cost two digits): the block also sums its real-pair forces, which carry the offset.
template <int L, int VIR>
__global__ void __launch_bounds__(EAMZ_T, 8)
k_eamz_force_f32(int n, int n_real, const Rec16 *__restrict__ recs,
                 const uint32_t *__restrict__ ls_ptr, const uint32_t *__restrict__ ls_col,
                 const int *__restrict__ perm, ZPair z, QScale qs,
                 const double *__restrict__ fembed, double *__restrict__ eatom,
                 double *__restrict__ forces, double *__restrict__ partial,
                 const int *__restrict__ own_mask) {
    __shared__ float s_gh[VIR ? 9 : 1][EAMZ_T];
    __shared__ double s_f3[EAMZ_T / 32][3];
    const LaneRow<L> row(n, ls_ptr, ls_col);
    float fx = 0, fy = 0, fz = 0, ep = 0;
    float vxx = 0, vyy = 0, vzz = 0, vyz = 0, vxz = 0, vxy = 0;
    if (VIR) {
#pragma unroll
        for (int q = 0; q < 9; ++q) s_gh[q][threadIdx.x] = 0.f;
    }
    Rec16 me;
    me.qx = me.qy = me.qz = 0;
    me.w = 0.f;
    if (row.active) me = recs[row.idx];
    if (row.steps > 0) {
        uint32_t c0 = row.cp[0];
        uint32_t c1 = row.cp[32];
        Rec16 a = recs[c0 & TAB_COL_IDX_MASK];
#pragma unroll 2
        for (int t = 0; t < row.steps; ++t) {
            const Rec16 an = recs[c1 & TAB_COL_IDX_MASK];
            const uint32_t cc = c0;
            c0 = c1;
            c1 = row.cp[(size_t)(t + 2) * 32u];
            const float dx = (float)(a.qx - me.qx), dy = (float)(a.qy - me.qy),
                        dz = (float)(a.qz - me.qz);
            const float s = fmaf(dx, dx, fmaf(dy, dy, fmaf(dz, dz, qs.eps_q)));
            const float rinv = f_rsqrt(s);           // 1 / r_q
            const float x = (s * rinv) * qs.x_per_q;
            // both denominators through one reciprocal
            const float ua = x - z.f_k_a, ub = x - z.f_k_b;
            const float ua2 = ua * ua, ua4 = ua2 * ua2, ua8 = ua4 * ua4, ua16 = ua8 * ua8;
            const float ub2 = ub * ub, ub4 = ub2 * ub2, ub8 = ub4 * ub4, ub16 = ub8 * ub8;
            const float da = fmaf(ua16, ua4, 1.f), db = fmaf(ub16, ub4, 1.f);
            const float rr = f_rcp(da * db);
            const float qa = db * rr, qb = da * rr;
            const float ga = f_ex2(fmaf(x, z.f_yx_a, z.f_yc_a)) * qa;
            const float gb = f_ex2(fmaf(x, z.f_yx_b, z.f_yc_b)) * qb;
            // df/dr = f (-(20 / re) u^19 q - b / re)
            const float dga = ga * fmaf(-(ua16 * ua2) * (ua * z.f_20_re), qa, z.f_nbre_a);
            const float dgb = gb * fmaf(-(ub16 * ub2) * (ub * z.f_20_re), qb, z.f_nbre_b);
            const bool in = s < qs.rc2_q;
            const float der = fmaf(me.w + a.w, dgb, dga);
            // (dE/dr / r) D = (dE/dr / (delta r_q)) (delta D_q): delta cancels
            const float sc = in ? der * rinv : 0.f;
            ep += in ? ga - gb : 0.f;
            if (VIR) {
                fx = fmaf(sc, dx, fx);
                fy = fmaf(sc, dy, fy);
                fz = fmaf(sc, dz, fz);
                if (in && (int)(cc & TAB_COL_IDX_MASK) >= n_real) {
                    const float gx = sc * dx, gy = sc * dy, gz = sc * dz;
                    volatile float *g = &s_gh[0][threadIdx.x];
                    g[0 * EAMZ_T] += gx;
                    g[1 * EAMZ_T] += gy;
                    g[2 * EAMZ_T] += gz;
                    g[3 * EAMZ_T] = fmaf(gx, dx, g[3 * EAMZ_T]);
                    g[4 * EAMZ_T] = fmaf(gy, dy, g[4 * EAMZ_T]);
                    g[5 * EAMZ_T] = fmaf(gz, dz, g[5 * EAMZ_T]);
                    g[6 * EAMZ_T] = fmaf(gy, dz, g[6 * EAMZ_T]);
                    g[7 * EAMZ_T] = fmaf(gx, dz, g[7 * EAMZ_T]);
                    g[8 * EAMZ_T] = fmaf(gx, dy, g[8 * EAMZ_T]);
                }
            } else {
                const float gx = sc * dx, gy = sc * dy, gz = sc * dz;
                fx += gx;
                fy += gy;
                fz += gz;
                vxx = fmaf(gx, dx, vxx);
                vyy = fmaf(gy, dy, vyy);
                vzz = fmaf(gz, dz, vzz);
                vyz = fmaf(gy, dz, vyz);
                vxz = fmaf(gx, dz, vxz);
                vxy = fmaf(gx, dy, vxy);
            }
            a = an;
        }
    }
    double acc[7];
    const double hd = 0.5 * qs.ddelta;
    double frx = 0.0, fry = 0.0, frz = 0.0;     // this lane's real-pair force sums
    if (VIR) {
        const float *g = &s_gh[0][threadIdx.x];
        // block origin = first atom of the block (always exists)
        const int first = (int)((blockIdx.x * blockDim.x) >> 5) * (32 / L);
        const Rec16 org = recs[first];
        const double px = (double)(me.qx - org.qx) * qs.ddelta,
                     py = (double)(me.qy - org.qy) * qs.ddelta,
                     pz = (double)(me.qz - org.qz) * qs.ddelta;
        if (row.active) {
            frx = (double)(fx - g[0 * EAMZ_T]);
            fry = (double)(fy - g[1 * EAMZ_T]);
            frz = (double)(fz - g[2 * EAMZ_T]);
        }
        acc[1] = fma(-frx, px, hd * (double)g[3 * EAMZ_T]);
        acc[2] = fma(-fry, py, hd * (double)g[4 * EAMZ_T]);
        acc[3] = fma(-frz, pz, hd * (double)g[5 * EAMZ_T]);
        acc[4] = fma(hd, (double)g[6 * EAMZ_T], -0.5 * fma(fry, pz, frz * py));
        acc[5] = fma(hd, (double)g[7 * EAMZ_T], -0.5 * fma(frx, pz, frz * px));
        acc[6] = fma(hd, (double)g[8 * EAMZ_T], -0.5 * fma(frx, py, fry * px));
    } else {
        acc[1] = hd * (double)vxx;
        acc[2] = hd * (double)vyy;
        acc[3] = hd * (double)vzz;
        acc[4] = hd * (double)vyz;
        acc[5] = hd * (double)vxz;
        acc[6] = hd * (double)vxy;
    }
    fx = lanes_sum<L>(fx);
    fy = lanes_sum<L>(fy);
    fz = lanes_sum<L>(fz);
    ep = lanes_sum<L>(ep);
    acc[0] = 0.0;
    bool mine = true;
    if (row.active) {
        const int o = perm[row.idx];
        if (own_mask) mine = own_mask[o] != 0;
        if (row.part == 0) {
            const double ei = fembed[row.idx] + 0.5 * (double)ep;
            if (eatom) eatom[o] = ei;
            if (forces) {
                forces[3 * (size_t)o + 0] = (double)fx;
                forces[3 * (size_t)o + 1] = (double)fy;
                forces[3 * (size_t)o + 2] = (double)fz;
            }
            acc[0] = ei;
        }
    }
    if (!mine || !row.active)
#pragma unroll
        for (int q = 0; q < 7; ++q) acc[q] = 0.0;
    if (VIR) {
        // - (sum of the block's real-pair forces) (x) block origin
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            frx += __shfl_xor_sync(0xffffffffu, frx, d);
            fry += __shfl_xor_sync(0xffffffffu, fry, d);
            frz += __shfl_xor_sync(0xffffffffu, frz, d);
        }
        if ((threadIdx.x & 31) == 0) {
            s_f3[threadIdx.x >> 5][0] = frx;
            s_f3[threadIdx.x >> 5][1] = fry;
            s_f3[threadIdx.x >> 5][2] = frz;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            double bx = 0, by = 0, bz = 0;
#pragma unroll
            for (int w = 0; w < EAMZ_T / 32; ++w) {
                bx += s_f3[w][0];
                by += s_f3[w][1];
                bz += s_f3[w][2];
            }
            const int first = (int)((blockIdx.x * blockDim.x) >> 5) * (32 / L);
            const Rec16 org = recs[first];
            const double ox = fma((double)org.qx, qs.ddelta, qs.ox),
                         oy = fma((double)org.qy, qs.ddelta, qs.oy),
                         oz = fma((double)org.qz, qs.ddelta, qs.oz);
            acc[1] -= bx * ox;
            acc[2] -= by * oy;
            acc[3] -= bz * oz;
            acc[4] -= 0.5 * (by * oz + bz * oy);
            acc[5] -= 0.5 * (bx * oz + bz * ox);
            acc[6] -= 0.5 * (bx * oy + by * ox);
        }
    }
    block_partials(acc, partial);
}

// F'(rho) fe / B - 1/2 into the float32 records (pass 2 of the 'medium' path reads Rec16.w)
__global__ void k_spread_w_rec(int n_owned, int n_loc, int n_ext, const double *__restrict__ v,
                               const double *__restrict__ halo_v, const int *__restrict__ perm,
                               const int *__restrict__ ghost_owner, double scale, double shift,
                               Rec16 *__restrict__ recs) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_ext) return;
    const int o = e < n_loc ? e : ghost_owner[e - n_loc];
    const double fp = o < n_owned ? v[o] : (halo_v ? halo_v[perm[o] - n_owned] : 0.0);
    recs[e].w = (float)fma(fp, scale, shift);
}
