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
//   * neighbour records prefetched D steps ahead through a register ring, the index stream two
//     ring lengths further; rows padded with a sentinel entry so that the loops carry no
//     per-lane trip counts and no predicated loads.
//   * r < rc mask: lists built with a skin hold entries beyond the model's cutoff; they
//     contribute exactly 0, so a reused list equals a fresh one (transformer/universal.py:58
//     rebuilds per call).  (Ordering the rows inside-first at build time and skipping the skin
//     part by warp vote was built and measured: the vote saves the arithmetic but not the
//     gather, pass 1 sits on the L1 data pipe and gained nothing, pass 2 gained 12 %, and the
//     1.4 ms ordering kernel per rebuild ate it -- profiles/README.md, r02e.)
//   * float32 ('medium'): 16-byte fixed-point records (tab_internal.h Rec16), geometry, functions
//     and per-atom sums in float32, block sums in float64.
#pragma once

#ifndef EAMZ_T
#define EAMZ_T 128
#endif
#ifndef EAMZ_L
#define EAMZ_L 1          // lanes per atom (template parameter of the kernels; A/B: profiles/r02*)
#endif
#ifndef EAMZ_MINB_RHO
#define EAMZ_MINB_RHO 6
#endif
#ifndef EAMZ_MINB_FORCE
#define EAMZ_MINB_FORCE 5
#endif
#ifndef EAMZ_PF64
#define EAMZ_PF64 1       // neighbour records in flight per thread (float64: 8 registers each)
#endif
#ifndef EAMZ_PF32
#define EAMZ_PF32 1       // float32: 4 registers each
#endif
#define EAMZ_PF_MAX (TAB_SPARE_ROWS / 2)   // 2 x this many rows are readable past the last group
#ifndef EAMZ_IDX_LD
#define EAMZ_IDX_LD(p) __ldcs(p)     // the index stream is read once: evict-first
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
        // the rows after the last group (the loops prefetch indices 2 x EAMZ_PF rows ahead)
        const int k = idx - n_pad;
        if (k < 32 * 2 * EAMZ_PF_MAX) ls_col[total_rows * 32u + k] = sentinel;
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

// The arrays the kernels traverse: for L = 1 the lists themselves when the builder left them
// sentinel-padded with spare rows (no copy), else the lane-split copy.
struct LsView {
    const uint32_t *ptr, *w, *col;
};

template <int L>
static int ensure_lanesplit(tab_nbr *nbr, cudaStream_t st, LsView &v) {
    if (L == 1 && nbr->col_padded) {
        v.ptr = nbr->slice_ptr.as<uint32_t>();
        v.w = nbr->slice_w.as<uint32_t>();
        v.col = nbr->col.as<uint32_t>();
        return TAB_OK;
    }
    if (nbr->ls_L == L) {
        v.ptr = nbr->ls_ptr.as<uint32_t>();
        v.w = nbr->ls_w.as<uint32_t>();
        v.col = nbr->ls_col.as<uint32_t>();
        return TAB_OK;
    }
    constexpr int G = 32 / L;
    const int n = nbr->n, n_groups = (n + G - 1) / G;
    TAB_TRY(nbr->ls_ptr.ensure(sizeof(uint32_t) * (size_t)(n_groups + 2)));
    TAB_TRY(nbr->ls_w.ensure(sizeof(uint32_t) * (size_t)(n_groups + 2)));
    TAB_TRY(nbr->stats.ensure(5 * sizeof(unsigned long long)));
    // widths of n_groups groups + one zero: the scan then leaves ls_ptr[n_groups] = total
    TAB_CUDA(cudaMemsetAsync(nbr->ls_w.as<uint32_t>() + n_groups, 0, sizeof(uint32_t), st));
    k_ls_widths<L><<<(n_groups + 127) / 128, 128, 0, st>>>(n, nbr->counts.as<int>(),
                                                           nbr->ls_w.as<uint32_t>());
    TAB_LAUNCH_CHECK();
    unsigned long long *d_total = nbr->stats.as<unsigned long long>() + 2;
    TAB_TRY(tab_scan_exclusive_u32(nbr->ls_w.as<uint32_t>(), nbr->ls_ptr.as<uint32_t>(),
                                   n_groups + 1, d_total, nbr->scan_tmp, st));
    unsigned long long total = 0;
    TAB_CUDA(cudaMemcpyAsync(&total, d_total, sizeof(total), cudaMemcpyDeviceToHost, st));
    TAB_CUDA(cudaStreamSynchronize(st));
    nbr->ls_rows = (long long)total;
    TAB_TRY(nbr->ls_col.ensure(sizeof(uint32_t) * 32 * (size_t)(total + 2 * EAMZ_PF_MAX)));
    const int n_pad = n_groups * G;
    k_ls_fill<L><<<(n_pad + 64 * EAMZ_PF_MAX + 127) / 128, 128, 0, st>>>(
        n, n_pad, (uint32_t)nbr->n_ext, nbr->counts.as<int>(), nbr->slice_ptr.as<uint32_t>(),
        nbr->col.as<uint32_t>(), nbr->ls_ptr.as<uint32_t>(), nbr->ls_col.as<uint32_t>(), total);
    TAB_LAUNCH_CHECK();
    v.ptr = nbr->ls_ptr.as<uint32_t>();      // (the buffers may just have been allocated)
    v.w = nbr->ls_w.as<uint32_t>();
    v.col = nbr->ls_col.as<uint32_t>();
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
                                       const uint32_t *__restrict__ ls_w,
                                       const uint32_t *__restrict__ ls_col) {
        constexpr int G = 32 / L;
        const int lane = threadIdx.x & 31;
        const int g = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
        idx = g * G + lane / L;
        part = lane % L;
        active = idx < n;
        uint32_t p0 = 0;
        steps = 0;
        if (g * G < n) {
            p0 = ls_ptr[g];
            steps = (int)ls_w[g];
        }
        cp = ls_col + ((size_t)p0 * 32u + lane);
    }
};

// Software-pipelined traversal of a lane's part of the row.  D records are in flight: a ring
// of D + 1 register slots, the record of step s + D is requested into the slot freed by step
// s - 1 before step s is evaluated (no register copies); the entries run two ring lengths
// further ahead (the index stream comes from DRAM).  Everything is unconditional: the rows past
// a group are readable (next group or the sentinel rows) and hold valid indices.
template <int D, typename Rec, typename F>
__device__ __forceinline__ void for_each_entry(const uint32_t *__restrict__ cp, int steps,
                                               const Rec *__restrict__ recs, F &&body) {
    constexpr int NB = D + 1;
    static_assert(3 * D + 2 <= 2 * EAMZ_PF_MAX, "spare rows after the last group");
    Rec buf[NB];
    uint32_t ci[NB], ci2[NB];
#pragma unroll
    for (int k = 0; k < D; ++k) buf[k] = recs[EAMZ_IDX_LD(cp + k * 32) & TAB_COL_IDX_MASK];
#pragma unroll
    for (int k = 0; k < NB; ++k) {
        // next use of slot k: step k + NB for k < D, step D for the free slot
        const int s1 = k < D ? k + NB : D;
        ci[k] = EAMZ_IDX_LD(cp + s1 * 32);
        ci2[k] = EAMZ_IDX_LD(cp + (s1 + NB) * 32);
    }
    for (int t = 0; t < steps; t += NB) {
#pragma unroll
        for (int d = 0; d < NB; ++d) {
            if (t + d < steps) {                 // warp-uniform
                const int sn = (d + D) % NB;     // slot of step s + D (freed by step s - 1)
                buf[sn] = recs[ci[sn] & TAB_COL_IDX_MASK];
                ci[sn] = ci2[sn];
                ci2[sn] = EAMZ_IDX_LD(cp + (size_t)(t + d + D + 2 * NB) * 32u);
                body(buf[d]);
            }
        }
    }
}

// The plain form of the same pipeline with one record in flight (the float32 kernels measured
// faster with it than with the ring: 0.175 / 0.286 vs 0.206 / 0.358 ms, profiles/r02a, r02c).
template <typename Rec, typename F>
__device__ __forceinline__ void for_each_entry_simple(const uint32_t *__restrict__ cp, int steps,
                                                      const Rec *__restrict__ recs, F &&body) {
    uint32_t c1 = EAMZ_IDX_LD(cp + 32);
    Rec a = recs[EAMZ_IDX_LD(cp) & TAB_COL_IDX_MASK];
#pragma unroll 2
    for (int t = 0; t < steps; ++t) {
        const Rec an = recs[c1 & TAB_COL_IDX_MASK];
        c1 = EAMZ_IDX_LD(cp + (size_t)(t + 2) * 32u);
        body(a);
        a = an;
    }
}

#ifndef EAMZ_RING32
#define EAMZ_RING32 0
#endif
#ifndef EAMZ_RING64
#define EAMZ_RING64 1
#endif

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
#ifndef EAMZ_RHO_TABLE
#define EAMZ_RHO_TABLE 1  // 0: table-free exponential in pass 1 (6 more FP64 ops, no LDS) --
#endif                    // measured 0.362 vs 0.346 ms with the table (profiles/r02c)

struct ZT2 {
    double tr, tc;        // natural exponent t = tr * r + tc  (table-free exp of pass 1)
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
    t.tr = -b / re;
    t.tc = b + log(a);
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
    double e = 1.2417843701716925e-12;               // c^5 / 120
    e = fma(e, fr, 5.7328516886404021e-10);          // c^4 / 24
    e = fma(e, fr, 2.1173137155464775e-07);          // c^3 / 6
    e = fma(e, fr, 5.8649049550561697e-05);          // c^2 / 2
    e = fma(e, fr, 1.0830424696249145e-02);          // c
    e = fma(e, fr, 1.0);
    e *= etab[n & 63];
    return __hiloint2double(__double2hiint(e) + ((n >> 6) << 20), __double2loint(e));
}

template <bool DERIV, bool TABLE = true>
__device__ __forceinline__ void zt2_eval(double r, const ZT2 &p,
                                         const double *__restrict__ etab, double &f,
                                         double &df) {
    const double u = fma(r, p.ire, p.nkappa);
    const double u2 = u * u, u4 = u2 * u2, u8 = u4 * u4, u16 = u8 * u8;
    const double q = tab_rcp(fma(u16, u4, 1.0));
    const double e = TABLE ? zexp64(fma(r, p.yr, p.yc), etab)
                           : zexp<0>(fma(r, p.tr, p.tc), nullptr);
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
           const uint32_t *__restrict__ ls_w, const uint32_t *__restrict__ ls_col, const int *__restrict__ perm, ZPair z,
           double rc2m, tab_fn embed0, double *__restrict__ fprime,
           double *__restrict__ fembed, double *__restrict__ fprime_caller) {
    __shared__ double s_etab[EAMZ_RHO_TABLE ? 64 : 1];
    if (EAMZ_RHO_TABLE) load_exp2_tab64(s_etab);
    const LaneRow<L> row(n, ls_ptr, ls_w, ls_col);
    double rho = 0.0;
    if (row.steps > 0) {
        Atom4 me;
        me.x = me.y = me.z = 0.0;
        if (row.active) me = atoms[row.idx];
        auto body = [&](const Atom4 &a) {
            const double dx = a.x - me.x, dy = a.y - me.y, dz = a.z - me.z;
            const double s = fma(dx, dx, fma(dy, dy, fma(dz, dz, 1e-14)));
            const double r = s * tab_rsqrt(s);
            double f, df;
            zt2_eval<false, EAMZ_RHO_TABLE != 0>(r, z.rho, s_etab, f, df);
            rho += s < rc2m ? f : 0.0;
        };
        if (EAMZ_RING64) for_each_entry<EAMZ_PF64>(row.cp, row.steps, atoms, body);
        else for_each_entry_simple(row.cp, row.steps, atoms, body);
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
// Virial: every undirected pair is visited from both ends, half of g (x) D each.
// (The form  W = -sum_i F_i^real (x) R_i + 1/2 sum_{image pairs} g (x) D  saves the six
// per-pair FMAs but was SLOWER on the 1 M-atom box, 0.79 vs 0.68 ms: the warps next to the
// box faces -- 17 % of all -- take the image branch on nearly every step.  profiles/r02a.)
// ---------------------------------------------------------------------------
template <int L>
__global__ void __launch_bounds__(EAMZ_T, EAMZ_MINB_FORCE)
k_eamz_force(int n, const Atom4 *__restrict__ atoms, const uint32_t *__restrict__ ls_ptr,
             const uint32_t *__restrict__ ls_w, const uint32_t *__restrict__ ls_col, const int *__restrict__ perm, ZPair z,
             double rc2m, const double *__restrict__ fembed, double *__restrict__ eatom,
             double *__restrict__ forces, double *__restrict__ partial,
             const int *__restrict__ own_mask) {
    __shared__ double s_etab[64];
    load_exp2_tab64(s_etab);
    const LaneRow<L> row(n, ls_ptr, ls_w, ls_col);
    double fx = 0, fy = 0, fz = 0, ep = 0;
    double vxx = 0, vyy = 0, vzz = 0, vyz = 0, vxz = 0, vxy = 0;
    if (row.steps > 0) {
        Atom4 me;
        me.x = me.y = me.z = me.w = 0.0;
        if (row.active) me = atoms[row.idx];
        auto body = [&](const Atom4 &a) {
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
        };
        if (EAMZ_RING64) for_each_entry<EAMZ_PF64>(row.cp, row.steps, atoms, body);
        else for_each_entry_simple(row.cp, row.steps, atoms, body);
    }
    double acc[7];
    acc[1] = 0.5 * vxx;
    acc[2] = 0.5 * vyy;
    acc[3] = 0.5 * vzz;
    acc[4] = 0.5 * vyz;
    acc[5] = 0.5 * vxz;
    acc[6] = 0.5 * vxy;
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
               const uint32_t *__restrict__ ls_w, const uint32_t *__restrict__ ls_col, const int *__restrict__ perm, ZPair z,
               QScale qs, tab_fn embed0, double *__restrict__ fprime,
               double *__restrict__ fembed, double *__restrict__ fprime_caller) {
    const LaneRow<L> row(n, ls_ptr, ls_w, ls_col);
    float rho = 0.f;
    if (row.steps > 0) {
        Rec16 me;
        me.qx = me.qy = me.qz = 0;
        if (row.active) me = recs[row.idx];
        auto body = [&](const Rec16 &a) {
            const float dx = (float)(a.qx - me.qx), dy = (float)(a.qy - me.qy),
                        dz = (float)(a.qz - me.qz);
            const float s = fmaf(dx, dx, fmaf(dy, dy, fmaf(dz, dz, qs.eps_q)));
            const float x = (s * f_rsqrt(s)) * qs.x_per_q;
            const float u = x - z.f_k_rho;
            const float u2 = u * u, u4 = u2 * u2, u8 = u4 * u4, u16 = u8 * u8;
            const float f = f_ex2(fmaf(x, z.f_yx_rho, z.f_yc_rho)) * f_rcp(fmaf(u16, u4, 1.f));
            rho += s < qs.rc2_q ? f : 0.f;
        };
        if (EAMZ_RING32) for_each_entry<EAMZ_PF32>(row.cp, row.steps, recs, body);
        else for_each_entry_simple(row.cp, row.steps, recs, body);
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

// Rec16.w = F'(rho) fe / B - 1/2 as in the float64 kernel.
template <int L>
__global__ void __launch_bounds__(EAMZ_T, 8)
k_eamz_force_f32(int n, const Rec16 *__restrict__ recs, const uint32_t *__restrict__ ls_ptr,
                 const uint32_t *__restrict__ ls_w, const uint32_t *__restrict__ ls_col, const int *__restrict__ perm, ZPair z,
                 QScale qs, const double *__restrict__ fembed, double *__restrict__ eatom,
                 double *__restrict__ forces, double *__restrict__ partial,
                 const int *__restrict__ own_mask) {
    const LaneRow<L> row(n, ls_ptr, ls_w, ls_col);
    float fx = 0, fy = 0, fz = 0, ep = 0;
    float vxx = 0, vyy = 0, vzz = 0, vyz = 0, vxz = 0, vxy = 0;
    if (row.steps > 0) {
        Rec16 me;
        me.qx = me.qy = me.qz = 0;
        me.w = 0.f;
        if (row.active) me = recs[row.idx];
        auto body = [&](const Rec16 &a) {
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
        };
        if (EAMZ_RING32) for_each_entry<EAMZ_PF32>(row.cp, row.steps, recs, body);
        else for_each_entry_simple(row.cp, row.steps, recs, body);
    }
    double acc[7];
    const double hd = 0.5 * qs.ddelta;      // g (x) D_q carries one factor delta
    acc[1] = hd * (double)vxx;
    acc[2] = hd * (double)vyy;
    acc[3] = hd * (double)vzz;
    acc[4] = hd * (double)vyz;
    acc[5] = hd * (double)vxz;
    acc[6] = hd * (double)vxy;
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
