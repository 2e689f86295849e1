// sf.cu -- Behler symmetry functions (G2 + G4) + per-element atomic MLPs:
// energies, forces and virial of the reference's AtomicNN.
//
// Replaces, for the reference:
//   nn/atomic/sf.py:79-119     G2_{i,T,tau} = sum_{p in T} exp(-eta (r-omega)^2/rc^2) fc(r)
//   nn/atomic/sf.py:121-182    G4_{i,T,tau} = 2^{1-zeta} sum_{j<k in T} (1+gamma cos)^zeta
//                              exp(-beta (rij^2+rik^2+rjk^2)/acut^2) fc fc fc
//   transformer/universal.py:115-233  triple enumeration (host Python loops) and the
//                              dense [12, Ta, N, nnl, ij2k] tensors -- never built here
//   nn/cutoff.py:20-85         cosine / polynomial cutoffs
//   nn/atomic/atomic.py:157-302, nn/convolutional.py:257-290   min-max + 1x1-conv MLP
//   nn/basic.py:276-331        F = -dE/dR, virial (TF autograd -> analytic backward)
//
// Structure (all atomic-free, deterministic):
//   k_sf_forward   block (128 threads) per centre atom; its neighbour row
//                  (species-sorted) is staged in shared memory; threads stride over
//                  neighbours (G2) and over the second index of the j<k pairs (G4);
//                  fixed-order block reduction.
//   k_mlp          warp per atom: forward + backward through the element's MLP,
//                  giving E_i and dE_i/dG_i.
//   k_sf_backward  block per centre atom: E_i depends on the D vectors of row i only,
//                  so g_p = dE_i/dD_p is computed for every entry p of the row by the
//                  thread that owns p (each unordered triple is visited from both of
//                  its legs -> no reduction, no atomics); g_p is stored per entry,
//                  sum_p g_p and sum_p g_p (x) D_p are accumulated on the fly.
//   k_sf_collect   thread per atom: F_i = sum_p g_p - sum_p g_rev(p)  (reverse-pair
//                  index built once per list, nbr.cu:k_build_reverse).
#include <stdlib.h>

#include "potentials.cuh"

#define SF_MAX_R 32      // radial parameter sets
#define SF_MAX_A 32      // angular parameter sets
#define SF_WARPS 4       // warps per block of the geometry kernels: ONE ATOM PER BLOCK
#define SF_TPA (SF_WARPS * 32)   // threads cooperating on one centre atom
#define MLP_WARPS 1      // k_mlp: one warp (= one atom) per block
#define MLP_MAX_LAYERS 8
#define MLP_MAX_WIDTH 256

struct SfDev {
    int n_el, n_r, n_a, angular, cutoff;   // cutoff: 0 cosine, 1 polynomial(gamma=5)
    int d_r, dim;                          // d_r = n_el*n_r*n_mom ; dim = full descriptor length
    // radial family (GRAP, nn/atomic/grap.py): 0 sf (= Behler G2), 1 morse, 2 density,
    // 3 pexp; moments = subset of {0, 1, 2, 3} in ascending order (layout: per term, per
    // tau, per moment).  New mode of the reference (grap.py:596-680): new_m0 = the m = 0
    // entry is sign(P) sqrt(P^2 + 1e-16); sym = traceless multiplicity tensor
    // (grap.py:485-494: m = 2 minus P_0^2 / 3, m = 3 minus 3/5 sum_a P_a^2).
    int rad_kind, n_mom, mom[4], has_m1, has_m2, has_m3, new_m0, sym;
    double rc, acut;
    double eta[SF_MAX_R], omega[SF_MAX_R], p3[SF_MAX_R];   // parameters 1, 2, 3 per set
    double beta[SF_MAX_A], gamma[SF_MAX_A], zeta[SF_MAX_A];
    double outer[SF_MAX_A];                // 2^(1 - zeta)
};

struct MlpDev {
    int n_layers;                          // hidden layers + output layer
    int act;                               // activation id
    int resnet, has_out_bias, has_minmax;
    int in[MLP_MAX_LAYERS], out[MLP_MAX_LAYERS];
    long long w_off[MLP_MAX_LAYERS], b_off[MLP_MAX_LAYERS];   // into the blob
    long long xlo_off, xhi_off;
};

#include "mlp_act.cuh"
#include "mlp_tc.cuh"

struct tab_atomic {
    SfDev sf;
    int n_el;
    MlpDev mlp[TAB_MAX_ELEMENTS];
    bool tc_ok = false;      // every element's network fits the tensor-core kernel
    DevBuf tc_nets;          // MlpTcDev [n_el]
    DevBuf tc_status;        // int: set by k_mlp_tc when an MMA barrier timed out
    DevBuf blob;        // double: weights, biases, xlo, xhi of every element + MlpDev table
    size_t mlp_table_off = 0;   // offset (in doubles) of the MlpDev table inside blob
    DevBuf G, dEdG, eat, gvec, fown;
    DevBuf mom;         // GRAP moment sums [n, n_el, n_r, 10 or 20] (forward -> backward)
};

// ---------------------------------------------------------------------------
// cutoff functions: value and derivative  (nn/cutoff.py:20-85)
// ---------------------------------------------------------------------------
// sin(pi x), cos(pi x) for x in [0, 1]: quadrant folding to |theta| <= pi/4 and the
// fdlibm kernel polynomials (|error| < 3e-18); ~20 FP64 instructions instead of the ~90
// of the library sincos -- the cutoff of r_jk is evaluated for every neighbour PAIR.
__device__ __forceinline__ void sincospi_unit(double x, double &sn, double &cs) {
    const int q = (int)(x * 2.0 + 0.5);                 // 0, 1, 2
    const double th = (x - 0.5 * (double)q) * 3.14159265358979323846;
    const double z = th * th;
    double ps = 1.58969099521155010221e-10;
    ps = fma(ps, z, -2.50507602534068634195e-08);
    ps = fma(ps, z, 2.75573137070700676789e-06);
    ps = fma(ps, z, -1.98412698298579493134e-04);
    ps = fma(ps, z, 8.33333333332248946124e-03);
    ps = fma(ps, z, -1.66666666666666324348e-01);
    const double s0 = fma(th * z, ps, th);
    double pc = -1.13596475577881948265e-11;
    pc = fma(pc, z, 2.08757232129817482790e-09);
    pc = fma(pc, z, -2.75573143513906633035e-07);
    pc = fma(pc, z, 2.48015872894767294178e-05);
    pc = fma(pc, z, -1.38888888888741095749e-03);
    pc = fma(pc, z, 4.16666666666666019037e-02);
    const double c0 = fma(z * z, pc, fma(-0.5, z, 1.0));
    sn = q == 1 ? c0 : (q == 2 ? -s0 : s0);
    cs = q == 1 ? -s0 : (q == 2 ? -c0 : c0);
}
__device__ __forceinline__ void sincospi_unit(float x, float &sn, float &cs) {
    sincospif(x, &sn, &cs);
}

template <typename Real>
__device__ __forceinline__ void cutoff_fn(int kind, Real r, Real rc, Real &f, Real &df) {
    if (r >= rc) {
        f = Real(0);
        df = Real(0);
        return;
    }
    const Real rci = Real(1) / rc;          // loop invariant at every call site
    const Real x = r * rci;
    if (kind == 0) {
        Real sn, cs;
        sincospi_unit(x, sn, cs);
        f = Real(0.5) * (cs + Real(1));
        df = Real(-0.5) * sn * Real(3.14159265358979323846) * rci;
    } else {
        const Real x2 = x * x, x4 = x2 * x2, x5 = x4 * x;
        f = Real(1) + Real(5) * x5 * x - Real(6) * x5;        // 1 + g x^(g+1) - (g+1) x^g, g=5
        df = (Real(30) * x5 - Real(30) * x4) * rci;
    }
}

template <typename Real>
__device__ __forceinline__ Real powz(Real base, Real z, Real &dpow) {
    // base^z and d/dbase; integer exponents by multiplication (sf.py uses pow)
    const int zi = (int)z;
    if ((Real)zi == z && zi >= 1 && zi <= 16) {
        Real p1 = Real(1);                     // base^(z-1)
        for (int k = 1; k < zi; ++k) p1 *= base;
        dpow = z * p1;
        return p1 * base;
    }
    const Real p = pow(base, z);
    dpow = (base != Real(0)) ? z * p / base : Real(0);
    return p;
}

// radial functions of the GRAP family, value and d/dr BEFORE the cutoff factor
// (grap.py:121-234; generic.py:15-30,87-100,120-168):
//   sf      exp(-eta (r - omega)^2 / rc^2)            (eta, omega)   = Behler G2
//   morse   D [e^{-2 g (r - r0)} - 2 e^{-g (r - r0)}]  (D, gamma, r0)
//   density A exp(-beta (r / re - 1))                  (A, beta, re)
//   pexp    exp(-(r / rl)^pl)                          (rl, pl)
template <typename Real>
__device__ __forceinline__ void rad_fn(const SfDev &sf, int tau, Real r, Real rc2i, Real &v,
                                       Real &dv) {
    const Real a = (Real)sf.eta[tau], b = (Real)sf.omega[tau];
    switch (sf.rad_kind) {
    case 0: {
        const Real d = r - b;
        v = Math<Real>::exp_(-a * d * d * rc2i);
        dv = Real(-2) * a * d * rc2i * v;
        break;
    }
    case 1: {
        const Real e1 = Math<Real>::exp_(-b * (r - (Real)sf.p3[tau]));
        v = a * (e1 * e1 - Real(2) * e1);
        dv = Real(-2) * a * b * (e1 * e1 - e1);
        break;
    }
    case 2: {
        const Real re = (Real)sf.p3[tau];
        v = a * Math<Real>::exp_(-b * (r / re - Real(1)));
        dv = -b / re * v;
        break;
    }
    default: {
        const Real x = r / a;
        Real xp1;                                    // x^(pl-1)
        if (b == Real(1)) xp1 = Real(1);
        else if (b == Real(2)) xp1 = x;
        else if (b == Real(3)) xp1 = x * x;
        else xp1 = Math<Real>::pow_(x, b - Real(1));
        v = Math<Real>::exp_(-xp1 * x);
        dv = -b * xp1 / a * v;
    }
    }
}

// moment sums of one (centre, term, tau): S0, Mx My Mz, Qxx Qyy Qzz Qyz Qxz Qxy and, for
// models with moment 3 (kernels instantiated with MW = 20), the ten unique third-order
// sums Txxx Txxy Txxz Txyy Txyz Txzz Tyyy Tyyz Tyzz Tzzz (multiplicities 1 3 3 3 6 3 1 3 3 1,
// grap.py:489-491)
#define GRAP_MOM_W 10
#define GRAP_MOM_W3 20
__device__ __forceinline__ double grap_mult3(int q) {     // q = 10..19
    return (q == 10 || q == 16 || q == 19) ? 1.0 : (q == 14 ? 6.0 : 3.0);
}
// any per-pair unit-vector moment (1, 2 or 3)
__device__ __forceinline__ bool grap_vec(const SfDev &sf) {
    return sf.has_m1 || sf.has_m2 || sf.has_m3;
}
// the backward / JVP kernels need the moment sums of the forward pass
__host__ __device__ __forceinline__ bool grap_need_mom(const SfDev &sf) {
    return sf.has_m1 || sf.has_m2 || sf.has_m3 || sf.new_m0;
}
// d/dP of sign(P) sqrt(P^2 + 1e-16)
__device__ __forceinline__ double grap_dm0(double P) {
    return fabs(P) / sqrt(P * P + 1e-16);
}

// shared-memory row layout per warp: 8 doubles per neighbour
//   0..2 D, 3 r, 4 fc(r; acut), 5 dfc(r; acut)/dr, 6 type (as double), 7 1/r (0 if r = 0)
#define ROW_W 8

template <typename Real>
__device__ __forceinline__ int stage_row(const SfDev &sf, int idx, int lane, int nthr,
                                         const Atom4 *__restrict__ atoms,
                                         const int *__restrict__ counts,
                                         const uint32_t *__restrict__ slice_ptr,
                                         const uint32_t *__restrict__ col,
                                         double *row) {
    const Atom4 me = atoms[idx];
    const int cnt = counts[idx];
    const uint32_t *cp = col + ((size_t)slice_ptr[idx >> 5] * 32u + (idx & 31));
    for (int k = lane; k < cnt; k += nthr) {
        const uint32_t c = cp[(size_t)k * 32u];
        const Atom4 a = atoms[c & TAB_COL_IDX_MASK];
        const double dx = a.x - me.x, dy = a.y - me.y, dz = a.z - me.z;
        const Real fx = (Real)dx, fy = (Real)dy, fz = (Real)dz;
        const Real r = sqrt(fx * fx + fy * fy + fz * fz + Math<Real>::eps());
        Real f, df;
        cutoff_fn<Real>(sf.cutoff, r, (Real)sf.acut, f, df);
        double *e = row + (size_t)k * ROW_W;
        e[0] = fx;
        e[1] = fy;
        e[2] = fz;
        e[3] = r;
        e[4] = f;
        e[5] = df;
        e[6] = (double)(c >> TAB_COL_TYPE_SHIFT);
        e[7] = r != Real(0) ? (double)(Real(1) / r) : 0.0;     // div_no_nan
    }
    __syncthreads();
    return cnt;
}

// sum over the SF_TPA threads of a block, result on every thread; fixed order
__device__ __forceinline__ double block_sum(double v, double *red) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < SF_WARPS; ++w) t += red[w];
    __syncthreads();
    return t;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

// index of the radial k-body term (centre ti, neighbour tj) inside the centre's
// term list [cc, c-x1, ...] (utils.py:262-273)
__device__ __forceinline__ int radial_term(int ti, int tj) {
    return ti == tj ? 0 : tj - (tj > ti ? 1 : 0) + 1;
}
// index of the sorted pair (a <= b) in the j<=k enumeration (utils.py:274-286)
__device__ __forceinline__ int pair_term(int a, int b, int nel) {
    return a * nel - a * (a - 1) / 2 + (b - a);
}

// ---------------------------------------------------------------------------
// forward: descriptors
// ---------------------------------------------------------------------------
template <typename Real, int MW>
__global__ void __launch_bounds__(SF_WARPS * 32)
k_sf_forward(int n, SfDev sf, int n_types, int row_cap,
             const Atom4 *__restrict__ atoms, const uint8_t *__restrict__ types_ext,
             const int *__restrict__ counts, const int *__restrict__ tcounts,
             const uint32_t *__restrict__ slice_ptr, const uint32_t *__restrict__ col,
             double *__restrict__ G, double *__restrict__ mom) {
    extern __shared__ __align__(16) double smem[];
    __shared__ double red[SF_WARPS];
    const int lane = threadIdx.x;           // thread index inside the atom's block
    const int idx = blockIdx.x;
    if (idx >= n) return;
    double *row = smem;
    const int cnt = stage_row<Real>(sf, idx, lane, SF_TPA, atoms, counts, slice_ptr, col, row);
    const int ti = (int)types_ext[idx];
    double *g = G + (size_t)idx * sf.dim;
    const Real rc = (Real)sf.rc, rc2i = Real(1) / (rc * rc);
    const Real ac2i = Real(1) / ((Real)sf.acut * (Real)sf.acut);
    // species segments of the row
    int seg[TAB_MAX_ELEMENTS + 1];
    seg[0] = 0;
    for (int t = 0; t < sf.n_el; ++t)
        seg[t + 1] = seg[t] + (t < n_types ? tcounts[(size_t)idx * n_types + t] : 0);
    (void)cnt;
    // ---- radial part: G2, or the GRAP family with multipole moments (grap.py:384-466)
    //      m = 0: sum_j w ;  m = 1: sum_a (sum_j w d_a / r)^2 ;
    //      m = 2: sum_ab (sum_j w d_a d_b / r^2)^2 ,   w = f(r) fc(r)
    for (int t = 0; t < sf.n_el; ++t) {
        const int term = radial_term(ti, t);
        for (int tau = 0; tau < sf.n_r; ++tau) {
            Real acc[MW];
#pragma unroll
            for (int q = 0; q < MW; ++q) acc[q] = Real(0);
            for (int k = seg[t] + lane; k < seg[t + 1]; k += SF_TPA) {
                const double *e = row + k * ROW_W;
                const Real r = (Real)e[3];
                Real f, df, v, dv;
                cutoff_fn<Real>(sf.cutoff, r, rc, f, df);
                rad_fn<Real>(sf, tau, r, rc2i, v, dv);
                const Real w = v * f;
                acc[0] += w;
                if (grap_vec(sf)) {
                    const Real ri = r != Real(0) ? Real(1) / r : Real(0);   // div_no_nan
                    const Real ux = (Real)e[0] * ri, uy = (Real)e[1] * ri, uz = (Real)e[2] * ri;
                    acc[1] += w * ux;
                    acc[2] += w * uy;
                    acc[3] += w * uz;
                    acc[4] += w * ux * ux;
                    acc[5] += w * uy * uy;
                    acc[6] += w * uz * uz;
                    acc[7] += w * uy * uz;
                    acc[8] += w * ux * uz;
                    acc[9] += w * ux * uy;
                    if (MW > GRAP_MOM_W) {
                        const Real wxx = w * ux * ux, wyy = w * uy * uy, wzz = w * uz * uz;
                        acc[10] += wxx * ux;
                        acc[11] += wxx * uy;
                        acc[12] += wxx * uz;
                        acc[13] += wyy * ux;
                        acc[14] += w * ux * uy * uz;
                        acc[15] += wzz * ux;
                        acc[16] += wyy * uy;
                        acc[17] += wyy * uz;
                        acc[18] += wzz * uy;
                        acc[19] += wzz * uz;
                    }
                }
            }
            double tot[MW];
            const int nsum = grap_need_mom(sf) ? MW : 1;
            for (int q = 0; q < nsum; ++q) tot[q] = block_sum((double)acc[q], red);
            if (lane == 0) {
                double *gg = g + (size_t)(term * sf.n_r + tau) * sf.n_mom;
                for (int mi = 0; mi < sf.n_mom; ++mi) {
                    const int mm = sf.mom[mi];
                    if (mm == 0)
                        gg[mi] = sf.new_m0 ? copysign(sqrt(tot[0] * tot[0] + 1e-16), tot[0]) *
                                                 (tot[0] != 0.0 ? 1.0 : 0.0)
                                           : tot[0];
                    else if (mm == 1) gg[mi] = tot[1] * tot[1] + tot[2] * tot[2] + tot[3] * tot[3];
                    else if (mm == 2)
                        gg[mi] = tot[4] * tot[4] + tot[5] * tot[5] + tot[6] * tot[6] +
                                 2.0 * (tot[7] * tot[7] + tot[8] * tot[8] + tot[9] * tot[9]) -
                                 (sf.sym ? tot[0] * tot[0] * (1.0 / 3.0) : 0.0);
                    else if (MW > GRAP_MOM_W) {
                        double g3 = 0.0;
                        for (int q = GRAP_MOM_W; q < MW; ++q) g3 += grap_mult3(q) * tot[q] * tot[q];
                        if (sf.sym)
                            g3 -= 0.6 * (tot[1] * tot[1] + tot[2] * tot[2] + tot[3] * tot[3]);
                        gg[mi] = g3;
                    }
                }
                if (mom && nsum > 1) {
                    double *mo = mom + ((size_t)idx * sf.n_el * sf.n_r + term * sf.n_r + tau) * MW;
                    for (int q = 0; q < MW; ++q) mo[q] = tot[q];
                }
            }
        }
    }
    if (!sf.angular) return;
    // ---- G4: species pairs a <= b, neighbour pairs p < q
    for (int a = 0; a < sf.n_el; ++a)
        for (int b = a; b < sf.n_el; ++b) {
            const int pt = pair_term(a, b, sf.n_el);
            Real acc[SF_MAX_A];
            for (int tau = 0; tau < sf.n_a; ++tau) acc[tau] = Real(0);
            // flattened pair index: a == b -> strict upper triangle of the segment,
            // a < b -> the full na x nb rectangle; threads stride over it
            const int na = seg[a + 1] - seg[a], nb = seg[b + 1] - seg[b];
            const long long npairs = (a == b) ? (long long)na * (na - 1) / 2
                                              : (long long)na * nb;
            for (long long t = lane; t < npairs; t += SF_TPA) {
                int p, q;
                if (a == b) {
                    // row-major upper triangle: solve for the row
                    const double nn1 = (double)na - 0.5;
                    int i = (int)floor(nn1 - sqrt(nn1 * nn1 - 2.0 * (double)t));
                    long long start = (long long)i * (2 * na - i - 1) / 2;
                    if (start > t) { --i; start = (long long)i * (2 * na - i - 1) / 2; }
                    else if (start + (na - i - 1) <= t) { start += na - i - 1; ++i; }
                    p = seg[a] + i;
                    q = p + 1 + (int)(t - start);
                } else {
                    p = seg[a] + (int)(t / nb);
                    q = seg[b] + (int)(t % nb);
                }
                {
                    const double *ep = row + p * ROW_W;
                    const Real fp = (Real)ep[4];
                    if (fp == Real(0)) continue;
                    const Real px = (Real)ep[0], py = (Real)ep[1], pz = (Real)ep[2],
                               r1 = (Real)ep[3];
                    const double *eq = row + q * ROW_W;
                    const Real fq = (Real)eq[4];
                    if (fq == Real(0)) continue;
                    const Real r2 = (Real)eq[3];

                    const Real jx = (Real)eq[0] - px, jy = (Real)eq[1] - py,
                               jz = (Real)eq[2] - pz;
                    const Real d3 = jx * jx + jy * jy + jz * jz + Math<Real>::eps();
                    const Real r3 = d3 * Math<Real>::rsqrt_(d3);
                    Real f3, df3;
                    cutoff_fn<Real>(sf.cutoff, r3, (Real)sf.acut, f3, df3);
                    if (f3 == Real(0)) continue;
                    const Real s2 = r1 * r1 + r2 * r2 + r3 * r3;
                    // cos(theta) = (r1^2 + r2^2 - r3^2) / (2 r1 r2) with the row's 1/r
                    // (zero when r = 0: divide_no_nan, sf.py:150)
                    const Real ct = (r1 * r1 + r2 * r2 - r3 * r3) *
                                    (Real(0.5) * (Real)ep[7] * (Real)eq[7]);
                    const Real fc = fp * (fq * f3);
                    Real E = Real(0);
                    for (int tau = 0; tau < sf.n_a; ++tau) {
                        Real dp;
                        const Real z = (Real)sf.zeta[tau];
                        const Real pw = powz<Real>(Real(1) + (Real)sf.gamma[tau] * ct, z, dp);
                        // the exponential depends on beta only (grid: beta outermost)
                        if (tau == 0 || sf.beta[tau] != sf.beta[tau - 1])
                            E = Math<Real>::exp_(-(Real)sf.beta[tau] * s2 * ac2i);
                        acc[tau] += pw * (E * fc) * (Real)sf.outer[tau];
                    }
                }
            }
            for (int tau = 0; tau < sf.n_a; ++tau) {
                const double tot = block_sum((double)acc[tau], red);
                if (lane == 0) g[sf.d_r + pt * sf.n_a + tau] = tot;
            }
        }
}

// ---------------------------------------------------------------------------
// MLP forward + backward, warp per atom  (convolutional.py:257-290)
// ---------------------------------------------------------------------------
// shared layout per warp: h[L][MLP_MAX_WIDTH] activations, dz[L][MLP_MAX_WIDTH]
// activation derivatives, x[dim] inputs, delta/ delta2 scratch
template <typename Real>
__global__ void __launch_bounds__(MLP_WARPS * 32)
k_mlp(int n, int dim, const uint8_t *__restrict__ types_ext, const MlpDev *__restrict__ mlps,
      const double *__restrict__ blob, const double *__restrict__ G,
      double *__restrict__ eat, double *__restrict__ dEdG) {
    extern __shared__ __align__(16) double smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int idx = blockIdx.x * MLP_WARPS + warp;
    if (idx >= n) return;
    const MlpDev &M = mlps[types_ext[idx]];
    const int L = M.n_layers;        // last one = output layer (no activation)
    const int stride = MLP_MAX_WIDTH;
    double *base = smem + (size_t)warp * ((2 * MLP_MAX_LAYERS + 2) * stride);
    double *h = base;                               // [L+1][stride] (h[0] = x)
    double *dz = base + (MLP_MAX_LAYERS + 1) * stride;   // [L][stride]
    double *dl = dz + MLP_MAX_LAYERS * stride;           // delta scratch [stride]
    const double *g = G + (size_t)idx * dim;
    // input (+ min-max normalisation, atomic.py:157-195)
    for (int k = lane; k < dim; k += 32) {
        double x = g[k];
        if (M.has_minmax) {
            const double lo = blob[M.xlo_off + k], hi = blob[M.xhi_off + k];
            const double den = hi - lo;
            x = den != 0.0 ? (hi - x) / den : 0.0;
        }
        h[k] = x;
    }
    __syncwarp();
    // forward
    for (int l = 0; l < L; ++l) {
        const double *W = blob + M.w_off[l];
        const double *bb = blob + M.b_off[l];
        const int ni = M.in[l], no = M.out[l];
        const double *hin = h + (size_t)l * stride;
        double *hout = h + (size_t)(l + 1) * stride;
        const bool last = l == L - 1;
        for (int o = lane; o < no; o += 32) {
            Real z = (last && !M.has_out_bias) ? Real(0) : (Real)bb[o];
            for (int k = 0; k < ni; ++k) z += (Real)hin[k] * (Real)W[(size_t)k * no + o];
            if (last) {
                hout[o] = z;
            } else {
                Real d;
                Real y = act_fn<Real>(M.act, z, d);
                dz[(size_t)l * stride + o] = d;
                if (l > 0 && M.resnet && no == ni) y += (Real)hin[o];
                hout[o] = y;
            }
        }
        __syncwarp();
    }
    if (lane == 0) eat[idx] = h[(size_t)L * stride];
    // backward: delta over the inputs of layer l+1
    const int nlast = M.in[L - 1];
    for (int k = lane; k < nlast; k += 32) dl[k] = blob[M.w_off[L - 1] + k];   // out width 1
    __syncwarp();
    for (int l = L - 2; l >= 0; --l) {
        const double *W = blob + M.w_off[l];
        const int ni = M.in[l], no = M.out[l];
        const bool res = l > 0 && M.resnet && no == ni;
        // dl holds dE/dh_{l+1} [no]; produce dE/dh_l [ni]
        double *tmp = h + (size_t)(l + 1) * stride;     // reuse as scratch: a = dl * act'
        for (int o = lane; o < no; o += 32) tmp[o] = dl[o] * dz[(size_t)l * stride + o];
        __syncwarp();
        double keep[MLP_MAX_WIDTH / 32];
        int c = 0;
        for (int k = lane; k < ni; k += 32, ++c) {
            Real s = Real(0);
            for (int o = 0; o < no; ++o) s += (Real)tmp[o] * (Real)W[(size_t)k * no + o];
            keep[c] = (double)s + (res ? dl[k] : 0.0);
        }
        __syncwarp();
        c = 0;
        for (int k = lane; k < ni; k += 32, ++c) dl[k] = keep[c];
        __syncwarp();
    }
    double *out = dEdG + (size_t)idx * dim;
    for (int k = lane; k < dim; k += 32) {
        double v = dl[k];
        if (M.has_minmax) {
            const double den = blob[M.xhi_off + k] - blob[M.xlo_off + k];
            v = den != 0.0 ? -v / den : 0.0;
        }
        out[k] = v;
    }
}

// ---------------------------------------------------------------------------
// backward: per-entry gradients g_p = dE_i/dD_p
// ---------------------------------------------------------------------------
template <typename Real, int MW>
__global__ void __launch_bounds__(SF_WARPS * 32)
k_sf_backward(int n, SfDev sf, int n_types, int row_cap,
              const Atom4 *__restrict__ atoms, const uint8_t *__restrict__ types_ext,
              const int *__restrict__ counts, const int *__restrict__ tcounts,
              const uint32_t *__restrict__ slice_ptr, const uint32_t *__restrict__ col,
              const double *__restrict__ dEdG, double *__restrict__ gvec,
              size_t plane, double *__restrict__ fown, double *__restrict__ partial,
              const double *__restrict__ mom) {
    extern __shared__ __align__(16) double smem[];
    __shared__ double red[SF_WARPS];
    const int lane = threadIdx.x;
    const int idx = blockIdx.x;
    double vir[6] = {0, 0, 0, 0, 0, 0};
    {
        double *row = smem;
        const int cnt = stage_row<Real>(sf, idx, lane, SF_TPA, atoms, counts, slice_ptr, col, row);
        const double *mo_atom = mom ? mom + (size_t)idx * sf.n_el * sf.n_r * MW : nullptr;
        const int ti = (int)types_ext[idx];
        const double *c = dEdG + (size_t)idx * sf.dim;
        const Real rc = (Real)sf.rc, rc2i = Real(1) / (rc * rc);
        const Real ac2i = Real(1) / ((Real)sf.acut * (Real)sf.acut);
        int seg[TAB_MAX_ELEMENTS + 1];
        seg[0] = 0;
        for (int t = 0; t < sf.n_el; ++t)
            seg[t + 1] = seg[t] + (t < n_types ? tcounts[(size_t)idx * n_types + t] : 0);
        const size_t ebase = (size_t)slice_ptr[idx >> 5] * 32u + (idx & 31);
        double fx = 0, fy = 0, fz = 0;
        // 4 sub-threads per row entry a: each covers every 4th partner b, the four
        // partial sums are combined with two shuffles (fixed order)
        const int sub = lane & 3;
        for (int a0 = 0; a0 < cnt; a0 += SF_TPA / 4) {
            const int a = a0 + (lane >> 2);
            const bool valid = a < cnt;
            const double *ea = row + (valid ? a : 0) * ROW_W;
            const Real ax = (Real)ea[0], ay = (Real)ea[1], az = (Real)ea[2], ra = (Real)ea[3];
            const Real fa = (Real)ea[4], dfa = (Real)ea[5], ira = (Real)ea[7];
            const int ta = (int)ea[6];
            Real sa = Real(0), wx = Real(0), wy = Real(0), wz = Real(0);
            if (valid && sf.angular && fa != Real(0)) {
                for (int b = sub; b < cnt; b += 4) {
                    if (b == a) continue;
                    const double *eb = row + b * ROW_W;
                    const Real fb = (Real)eb[4];
                    if (fb == Real(0)) continue;
                    const Real rb = (Real)eb[3];
                    const Real jx = ax - (Real)eb[0], jy = ay - (Real)eb[1],
                               jz = az - (Real)eb[2];
                    const Real dab = jx * jx + jy * jy + jz * jz + Math<Real>::eps();
                    const Real irab = Math<Real>::rsqrt_(dab);
                    const Real rab = dab * irab;
                    Real fab, dfab;
                    cutoff_fn<Real>(sf.cutoff, rab, (Real)sf.acut, fab, dfab);
                    if (fab == Real(0)) continue;
                    const int tb = (int)eb[6];
                    const int pt = ta <= tb ? pair_term(ta, tb, sf.n_el)
                                            : pair_term(tb, ta, sf.n_el);
                    const double *cc = c + sf.d_r + pt * sf.n_a;
                    const Real s2 = ra * ra + rb * rb + rab * rab;
                    // reciprocals from the row (0 when r = 0: divide_no_nan)
                    const Real irb = (Real)eb[7], iab2 = ira * irb;
                    const Real ct = (ra * ra + rb * rb - rab * rab) * (Real(0.5) * iab2);
                    const Real dct_da = iab2 != Real(0) ? irb - ct * ira : Real(0);
                    const Real dct_dab = -rab * iab2;
                    const Real F3 = fa * fb * fab;
                    // sum over the angular sets with the beta-independent factors pulled
                    // out: for one beta group, with A0 = sum K P and A1 = sum K gamma P',
                    //   d/dr_a  = E [F3 dct_da  A1 + A0 (fa' fb fab - 2 beta r_a  F3 / acut^2)]
                    //   d/dr_ab = E [F3 dct_dab A1 + A0 (fa fb fab' - 2 beta r_ab F3 / acut^2)]
                    // (the grid is beta-outermost, sf.py:49-51: one exponential per group)
                    Real va = Real(0), vab = Real(0), E = Real(0), A0 = Real(0), A1 = Real(0);
                    Real be_cur = Real(0);
                    const Real ta_ = dfa * fb * fab, tab_ = fa * fb * dfab;
                    for (int tau = 0; tau < sf.n_a; ++tau) {
                        const Real z = (Real)sf.zeta[tau], gm = (Real)sf.gamma[tau],
                                   be = (Real)sf.beta[tau];
                        if (tau == 0 || sf.beta[tau] != sf.beta[tau - 1]) {
                            if (tau > 0) {
                                const Real c0 = Real(-2) * be_cur * ac2i * F3;
                                va += E * (F3 * dct_da * A1 + A0 * (ta_ + c0 * ra));
                                vab += E * (F3 * dct_dab * A1 + A0 * (tab_ + c0 * rab));
                            }
                            E = Math<Real>::exp_(-be * s2 * ac2i);
                            A0 = Real(0);
                            A1 = Real(0);
                            be_cur = be;
                        }
                        Real dP;
                        const Real P = powz<Real>(Real(1) + gm * ct, z, dP);
                        const Real K = (Real)sf.outer[tau] * (Real)cc[tau];
                        A0 += K * P;
                        A1 += K * (dP * gm);
                    }
                    {
                        const Real c0 = Real(-2) * be_cur * ac2i * F3;
                        va += E * (F3 * dct_da * A1 + A0 * (ta_ + c0 * ra));
                        vab += E * (F3 * dct_dab * A1 + A0 * (tab_ + c0 * rab));
                    }
                    sa += va;
                    const Real q = vab * irab;
                    wx += q * jx;
                    wy += q * jy;
                    wz += q * jz;
                }
            }
            // combine the 4 sub-threads (all lanes take part in the shuffles)
#pragma unroll
            for (int d = 1; d <= 2; d <<= 1) {
                sa += __shfl_xor_sync(0xffffffffu, sa, d);
                wx += __shfl_xor_sync(0xffffffffu, wx, d);
                wy += __shfl_xor_sync(0xffffffffu, wy, d);
                wz += __shfl_xor_sync(0xffffffffu, wz, d);
            }
            if (!valid || sub != 0) continue;
            // ---- radial part (G2 / GRAP moments)
            Real s_r = Real(0);                       // dE/dr_a
            {
                Real f, df;
                cutoff_fn<Real>(sf.cutoff, ra, rc, f, df);
                const int term = radial_term(ti, ta);
                const Real ri = ira;
                const Real ux = ax * ri, uy = ay * ri, uz = az * ri;
                for (int tau = 0; tau < sf.n_r; ++tau) {
                    Real v, dv;
                    rad_fn<Real>(sf, tau, ra, rc2i, v, dv);
                    const Real w = v * f, dw = dv * f + v * df;
                    const double *cc = c + (size_t)(term * sf.n_r + tau) * sf.n_mom;
                    const double *mo = mo_atom ? mo_atom + (size_t)(term * sf.n_r + tau) * MW
                                               : nullptr;
                    for (int mi = 0; mi < sf.n_mom; ++mi) {
                        const int mm = sf.mom[mi];
                        const Real ck = (Real)cc[mi];
                        if (mm == 0) {
                            s_r += (sf.new_m0 && mo ? ck * (Real)grap_dm0(mo[0]) : ck) * dw;
                        } else if (mm == 1) {
                            // G = |M|^2, M = sum w u:  dG/dD = 2 [ (dw - w/r)(M.u) u + (w/r) M ]
                            const Real Mx = (Real)mo[1], My = (Real)mo[2], Mz = (Real)mo[3];
                            const Real mu = Mx * ux + My * uy + Mz * uz;
                            const Real wr = w * ri;
                            s_r += Real(2) * ck * (dw - wr) * mu;
                            wx += Real(2) * ck * wr * Mx;
                            wy += Real(2) * ck * wr * My;
                            wz += Real(2) * ck * wr * Mz;
                        } else if (mm == 2) {
                            // G = sum_ab Q_ab^2, Q = sum w u (x) u:
                            //   dG/dD = 2 [ (dw - 2 w/r)(u.Q.u) u + 2 (w/r) Q.u ]
                            const Real Qxx = (Real)mo[4], Qyy = (Real)mo[5], Qzz = (Real)mo[6],
                                       Qyz = (Real)mo[7], Qxz = (Real)mo[8], Qxy = (Real)mo[9];
                            const Real qx = Qxx * ux + Qxy * uy + Qxz * uz;
                            const Real qy = Qxy * ux + Qyy * uy + Qyz * uz;
                            const Real qz = Qxz * ux + Qyz * uy + Qzz * uz;
                            const Real uqu = qx * ux + qy * uy + qz * uz;
                            const Real wr = w * ri;
                            s_r += Real(2) * ck * (dw - Real(2) * wr) * uqu;
                            wx += Real(4) * ck * wr * qx;
                            wy += Real(4) * ck * wr * qy;
                            wz += Real(4) * ck * wr * qz;
                            if (sf.sym)       // minus P_0^2 / 3
                                s_r -= ck * Real(2.0 / 3.0) * (Real)mo[0] * dw;
                        } else if (MW > GRAP_MOM_W) {
                            // G = sum_abc T_abc^2, T = sum w u (x) u (x) u:
                            //   dG/dD = 2 [ (dw - 3 w/r)(T:uuu) u + 3 (w/r) T:uu ]
                            const Real Txxx = (Real)mo[10], Txxy = (Real)mo[11], Txxz = (Real)mo[12],
                                       Txyy = (Real)mo[13], Txyz = (Real)mo[14], Txzz = (Real)mo[15],
                                       Tyyy = (Real)mo[16], Tyyz = (Real)mo[17], Tyzz = (Real)mo[18],
                                       Tzzz = (Real)mo[19];
                            const Real xx = ux * ux, yy = uy * uy, zz = uz * uz, xy = ux * uy,
                                       xz = ux * uz, yz = uy * uz;
                            const Real tx = Txxx * xx + Txyy * yy + Txzz * zz +
                                            Real(2) * (Txxy * xy + Txxz * xz + Txyz * yz);
                            const Real ty = Txxy * xx + Tyyy * yy + Tyzz * zz +
                                            Real(2) * (Txyy * xy + Txyz * xz + Tyyz * yz);
                            const Real tz = Txxz * xx + Tyyz * yy + Tzzz * zz +
                                            Real(2) * (Txyz * xy + Txzz * xz + Tyzz * yz);
                            const Real tuuu = tx * ux + ty * uy + tz * uz;
                            const Real wr = w * ri;
                            s_r += Real(2) * ck * (dw - Real(3) * wr) * tuuu;
                            wx += Real(6) * ck * wr * tx;
                            wy += Real(6) * ck * wr * ty;
                            wz += Real(6) * ck * wr * tz;
                            if (sf.sym) {     // minus 3/5 |M|^2
                                const Real cs = Real(-0.6) * ck;
                                const Real Mx = (Real)mo[1], My = (Real)mo[2], Mz = (Real)mo[3];
                                const Real mu = Mx * ux + My * uy + Mz * uz;
                                s_r += Real(2) * cs * (dw - wr) * mu;
                                wx += Real(2) * cs * wr * Mx;
                                wy += Real(2) * cs * wr * My;
                                wz += Real(2) * cs * wr * Mz;
                            }
                        }
                    }
                }
            }
            const Real q = (s_r + sa) * ira;
            const Real gx = q * ax + wx, gy = q * ay + wy, gz = q * az + wz;
            const size_t e = ebase + (size_t)a * 32u;
            gvec[e] = (double)gx;
            gvec[plane + e] = (double)gy;
            gvec[2 * plane + e] = (double)gz;
            fx += (double)gx;
            fy += (double)gy;
            fz += (double)gz;
            vir[0] += (double)(gx * ax);
            vir[1] += (double)(gy * ay);
            vir[2] += (double)(gz * az);
            vir[3] += 0.5 * (double)(gy * az + gz * ay);
            vir[4] += 0.5 * (double)(gx * az + gz * ax);
            vir[5] += 0.5 * (double)(gx * ay + gy * ax);
        }
        fx = block_sum(fx, red);
        fy = block_sum(fy, red);
        fz = block_sum(fz, red);
        if (lane == 0) {
            fown[3 * (size_t)idx + 0] = fx;
            fown[3 * (size_t)idx + 1] = fy;
            fown[3 * (size_t)idx + 2] = fz;
        }
    }
#pragma unroll
    for (int q = 0; q < 6; ++q) {
        const double v = block_sum(vir[q], red);
        if (lane == 0) partial[(size_t)blockIdx.x * 8 + 1 + q] = v;
    }
}

// F_i = sum_p g_p - sum_p g_rev(p); per-atom energies to caller order; energy
// partials.  One warp per atom (lanes stride over the row), 4 atoms per block.
__global__ void __launch_bounds__(128)
k_sf_collect(int n, int n_loc, const int *__restrict__ counts,
             const uint32_t *__restrict__ slice_ptr, const uint32_t *__restrict__ col,
             const uint32_t *__restrict__ rev, const int *__restrict__ ghost_owner,
             const int *__restrict__ perm, const double *__restrict__ gvec, size_t plane,
             const double *__restrict__ fown, const double *__restrict__ eat,
             double *__restrict__ eatom, double *__restrict__ forces,
             double *__restrict__ partial_e) {
    __shared__ double red[4];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int idx = blockIdx.x * 4 + warp;
    double e = 0.0;
    if (idx < n) {
        double fx = 0.0, fy = 0.0, fz = 0.0;
        const size_t base = (size_t)slice_ptr[idx >> 5] * 32u + (idx & 31);
        const int cnt = counts[idx];
        for (int k = lane; rev && k < cnt; k += 32) {
            const size_t ent = base + (size_t)k * 32u;
            const uint32_t q = rev[ent];
            if (q == 0xFFFFFFFFu) continue;
            int o = (int)(col[ent] & TAB_COL_IDX_MASK);
            if (o >= n_loc) o = ghost_owner[o - n_loc];
            const size_t re = (size_t)slice_ptr[o >> 5] * 32u + (o & 31) + (size_t)q * 32u;
            fx -= gvec[re];
            fy -= gvec[plane + re];
            fz -= gvec[2 * plane + re];
        }
        fx = warp_sum(fx);
        fy = warp_sum(fy);
        fz = warp_sum(fz);
        if (lane == 0) {
            const int o = perm[idx];
            if (forces) {
                forces[3 * (size_t)o + 0] = fx + fown[3 * (size_t)idx];
                forces[3 * (size_t)o + 1] = fy + fown[3 * (size_t)idx + 1];
                forces[3 * (size_t)o + 2] = fz + fown[3 * (size_t)idx + 2];
            }
            e = eat[idx];
            if (eatom) eatom[o] = e;
        }
    }
    if (lane == 0) red[warp] = e;
    __syncthreads();
    if (threadIdx.x == 0) partial_e[(size_t)blockIdx.x] = red[0] + red[1] + red[2] + red[3];
}

// batch handles: one block per structure sums its atoms' energies (eat, sorted = caller
// order) and the per-atom virial rows of k_sf_backward (partial[idx*8+1..6]) in fixed order
__global__ void __launch_bounds__(128)
k_sf_reduce_batch(const int *__restrict__ struct_off, const double *__restrict__ eat,
                  const double *__restrict__ partial, double *__restrict__ energy,
                  double *__restrict__ virial) {
    __shared__ double sm[128][7];
    const int s = blockIdx.x;
    const int lo = struct_off[s], hi = struct_off[s + 1];
    double a[7] = {0, 0, 0, 0, 0, 0, 0};
    for (int i = lo + threadIdx.x; i < hi; i += 128) {
        if (eat) a[0] += eat[i];
        if (partial)
#pragma unroll
            for (int q = 1; q < 7; ++q) a[q] += partial[(size_t)i * 8 + q];
    }
#pragma unroll
    for (int q = 0; q < 7; ++q) sm[threadIdx.x][q] = a[q];
    __syncthreads();
    for (int w = 64; w > 0; w >>= 1) {
        if (threadIdx.x < w)
#pragma unroll
            for (int q = 0; q < 7; ++q) sm[threadIdx.x][q] += sm[threadIdx.x + w][q];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        if (energy) energy[s] = sm[0][0];
        if (virial) {
            double *v = virial + 9 * (size_t)s;
            const double xx = sm[0][1], yy = sm[0][2], zz = sm[0][3], yz = sm[0][4],
                         xz = sm[0][5], xy = sm[0][6];
            v[0] = xx; v[1] = xy; v[2] = xz;
            v[3] = xy; v[4] = yy; v[5] = yz;
            v[6] = xz; v[7] = yz; v[8] = zz;
        }
    }
}

// spatial decomposition: only the atoms with mask[caller index] != 0 (the rank's OWN atoms;
// the other rows belong to the inner halo and are recomputed by their owner ranks) count
// towards this rank's partial energy / virial
__global__ void __launch_bounds__(256)
k_sf_reduce_masked(int n, const int *__restrict__ perm, const int *__restrict__ mask,
                   const double *__restrict__ eat, const double *__restrict__ partial,
                   double *__restrict__ energy, double *__restrict__ virial) {
    __shared__ double sm[256][7];
    double a[7] = {0, 0, 0, 0, 0, 0, 0};
    for (int i = threadIdx.x; i < n; i += 256) {
        if (!mask[perm[i]]) continue;
        a[0] += eat[i];
        if (partial)
#pragma unroll
            for (int q = 1; q < 7; ++q) a[q] += partial[(size_t)i * 8 + q];
    }
#pragma unroll
    for (int q = 0; q < 7; ++q) sm[threadIdx.x][q] = a[q];
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if (threadIdx.x < w)
#pragma unroll
            for (int q = 0; q < 7; ++q) sm[threadIdx.x][q] += sm[threadIdx.x + w][q];
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

// fixed-order final sums: energy from partial_e[nb_e], virial from partial[nb_v*8+1..6]
__global__ void __launch_bounds__(256)
k_sf_reduce(int nb_e, const double *__restrict__ partial_e, int nb_v,
            const double *__restrict__ partial, double *__restrict__ energy,
            double *__restrict__ virial) {
    __shared__ double sm[256][7];
    double a[7] = {0, 0, 0, 0, 0, 0, 0};
    for (int b = threadIdx.x; b < nb_e; b += 256) a[0] += partial_e[b];
    for (int b = threadIdx.x; b < nb_v; b += 256)
#pragma unroll
        for (int q = 1; q < 7; ++q) a[q] += partial[(size_t)b * 8 + q];
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
// JVP: T[i, f] = sum_{p in row i} dG_{i,f}/dD_p . dD_p  with the per-pair
// displacement  dD_p = u_i - u_owner(j) + A D_p.
// This is the transpose of the force/virial assembly
//   F = J^T c  (forces), W = sum_p g_p (x) D_p   with c = dE/dG:
//   d/dc [ sum_a F_a.u_a + sum_ab A_ab W_ab ] = T.
// It is the backward of the force op in training (force / stress losses need
// d(loss)/d(parameters) through dE/dG) -- the reference gets it from TF's
// second-order autograd (nn/opt.py:132-157).
// ---------------------------------------------------------------------------
template <typename Real, int MW>
__global__ void __launch_bounds__(SF_WARPS * 32)
k_sf_jvp(int n, int n_loc, SfDev sf, int n_types, int row_cap,
         const Atom4 *__restrict__ atoms, const uint8_t *__restrict__ types_ext,
         const int *__restrict__ counts, const int *__restrict__ tcounts,
         const uint32_t *__restrict__ slice_ptr, const uint32_t *__restrict__ col,
         const int *__restrict__ ghost_owner, const double *__restrict__ u,
         const double *__restrict__ A, const int *__restrict__ struct_of,
         const double *__restrict__ mom, double *__restrict__ T) {
    extern __shared__ __align__(16) double smem[];
    __shared__ double red[SF_WARPS];
    const int lane = threadIdx.x;
    const int idx = blockIdx.x;
    if (idx >= n) return;
    if (struct_of) A += 9 * (size_t)struct_of[idx];     // batch: A [n_struct, 9]
    double *row = smem;
    double *dd = row + (size_t)row_cap * ROW_W;          // [row_cap][4]: dD_p
    const int cnt = stage_row<Real>(sf, idx, lane, SF_TPA, atoms, counts, slice_ptr, col, row);
    {
        const uint32_t *cp = col + ((size_t)slice_ptr[idx >> 5] * 32u + (idx & 31));
        const double ux = u[3 * (size_t)idx], uy = u[3 * (size_t)idx + 1],
                     uz = u[3 * (size_t)idx + 2];
        for (int k = lane; k < cnt; k += SF_TPA) {
            int j = (int)(cp[(size_t)k * 32u] & TAB_COL_IDX_MASK);
            if (j >= n_loc) j = ghost_owner[j - n_loc];
            const double *e = row + (size_t)k * ROW_W;
            dd[k * 4 + 0] = ux - u[3 * (size_t)j] + A[0] * e[0] + A[1] * e[1] + A[2] * e[2];
            dd[k * 4 + 1] = uy - u[3 * (size_t)j + 1] + A[3] * e[0] + A[4] * e[1] + A[5] * e[2];
            dd[k * 4 + 2] = uz - u[3 * (size_t)j + 2] + A[6] * e[0] + A[7] * e[1] + A[8] * e[2];
        }
        __syncthreads();
    }
    const int ti = (int)types_ext[idx];
    double *t = T + (size_t)idx * sf.dim;
    const Real rc = (Real)sf.rc, rc2i = Real(1) / (rc * rc);
    const Real ac2i = Real(1) / ((Real)sf.acut * (Real)sf.acut);
    int seg[TAB_MAX_ELEMENTS + 1];
    seg[0] = 0;
    for (int s = 0; s < sf.n_el; ++s)
        seg[s + 1] = seg[s] + (s < n_types ? tcounts[(size_t)idx * n_types + s] : 0);
    // ---- radial part: G2 or the GRAP families with multipole moments.  With
    //      w = f(r) fc(r), u = D / r, s = u . dD and dD_perp = dD - s u:
    //        m = 0:  dG = sum w' s
    //        m = 1:  G = |M|^2,      dG = 2 M . dM,  dM = sum [w' s u + (w / r) dD_perp]
    //        m = 2:  G = sum Q_ab^2, dG = 2 Q : dQ,
    //                dQ = sum [w' s u (x) u + (w / r) (dD_perp (x) u + u (x) dD_perp)]
    //      (M, Q = the moment sums of the forward pass, `mom`)
    //        m = 3:  G = sum T_abc^2, dG = 2 T : dT,
    //                dT = sum [w' s u(x)u(x)u + (w / r) sym(dD_perp (x) u (x) u)]
    const bool grap = grap_vec(sf);
    const double *mo_atom = (grap_need_mom(sf) && mom)
                                ? mom + (size_t)idx * sf.n_el * sf.n_r * MW : nullptr;
    for (int s = 0; s < sf.n_el; ++s) {
        const int term = radial_term(ti, s);
        for (int tau = 0; tau < sf.n_r; ++tau) {
            Real acc[MW];
#pragma unroll
            for (int q = 0; q < MW; ++q) acc[q] = Real(0);
            for (int k = seg[s] + lane; k < seg[s + 1]; k += SF_TPA) {
                const double *e = row + k * ROW_W;
                const Real r = (Real)e[3], ri = (Real)e[7];
                Real f, df, v, dv;
                cutoff_fn<Real>(sf.cutoff, r, rc, f, df);
                rad_fn<Real>(sf, tau, r, rc2i, v, dv);
                const Real w = v * f, dw = dv * f + v * df;
                const Real ux = (Real)e[0] * ri, uy = (Real)e[1] * ri, uz = (Real)e[2] * ri;
                const Real dx = (Real)dd[k * 4], dy = (Real)dd[k * 4 + 1], dz = (Real)dd[k * 4 + 2];
                const Real sp = ux * dx + uy * dy + uz * dz;
                acc[0] += dw * sp;
                if (grap) {
                    const Real wr = w * ri;
                    const Real px = dx - sp * ux, py = dy - sp * uy, pz = dz - sp * uz;
                    const Real c1 = dw * sp;
                    acc[1] += c1 * ux + wr * px;
                    acc[2] += c1 * uy + wr * py;
                    acc[3] += c1 * uz + wr * pz;
                    acc[4] += c1 * ux * ux + Real(2) * wr * px * ux;
                    acc[5] += c1 * uy * uy + Real(2) * wr * py * uy;
                    acc[6] += c1 * uz * uz + Real(2) * wr * pz * uz;
                    acc[7] += c1 * uy * uz + wr * (py * uz + pz * uy);
                    acc[8] += c1 * ux * uz + wr * (px * uz + pz * ux);
                    acc[9] += c1 * ux * uy + wr * (px * uy + py * ux);
                    if (MW > GRAP_MOM_W) {
                        const Real xx = ux * ux, yy = uy * uy, zz = uz * uz;
                        acc[10] += c1 * xx * ux + Real(3) * wr * px * xx;
                        acc[11] += c1 * xx * uy + wr * (Real(2) * px * ux * uy + xx * py);
                        acc[12] += c1 * xx * uz + wr * (Real(2) * px * ux * uz + xx * pz);
                        acc[13] += c1 * ux * yy + wr * (px * yy + Real(2) * ux * py * uy);
                        acc[14] += c1 * ux * uy * uz +
                                   wr * (px * uy * uz + ux * py * uz + ux * uy * pz);
                        acc[15] += c1 * ux * zz + wr * (px * zz + Real(2) * ux * pz * uz);
                        acc[16] += c1 * yy * uy + Real(3) * wr * py * yy;
                        acc[17] += c1 * yy * uz + wr * (Real(2) * py * uy * uz + yy * pz);
                        acc[18] += c1 * uy * zz + wr * (py * zz + Real(2) * uy * pz * uz);
                        acc[19] += c1 * zz * uz + Real(3) * wr * pz * zz;
                    }
                }
            }
            double tot[MW];
            const int nsum = grap ? MW : 1;
            for (int q = 0; q < nsum; ++q) tot[q] = block_sum((double)acc[q], red);
            if (lane == 0) {
                double *tt = t + (size_t)(term * sf.n_r + tau) * sf.n_mom;
                const double *mo = mo_atom ? mo_atom + (size_t)(term * sf.n_r + tau) * MW
                                           : nullptr;
                for (int mi = 0; mi < sf.n_mom; ++mi) {
                    const int mm = sf.mom[mi];
                    if (mm == 0) tt[mi] = (sf.new_m0 && mo) ? grap_dm0(mo[0]) * tot[0] : tot[0];
                    else if (mm == 1)
                        tt[mi] = 2.0 * (mo[1] * tot[1] + mo[2] * tot[2] + mo[3] * tot[3]);
                    else if (mm == 2)
                        tt[mi] = 2.0 * (mo[4] * tot[4] + mo[5] * tot[5] + mo[6] * tot[6] +
                                        2.0 * (mo[7] * tot[7] + mo[8] * tot[8] + mo[9] * tot[9])) -
                                 (sf.sym ? (2.0 / 3.0) * mo[0] * tot[0] : 0.0);
                    else if (MW > GRAP_MOM_W) {
                        double g3 = 0.0;
                        for (int q = GRAP_MOM_W; q < MW; ++q) g3 += grap_mult3(q) * mo[q] * tot[q];
                        if (sf.sym)
                            g3 -= 0.6 * (mo[1] * tot[1] + mo[2] * tot[2] + mo[3] * tot[3]);
                        tt[mi] = 2.0 * g3;
                    }
                }
            }
        }
    }
    if (!sf.angular) return;
    // ---- G4
    for (int a = 0; a < sf.n_el; ++a)
        for (int b = a; b < sf.n_el; ++b) {
            const int pt = pair_term(a, b, sf.n_el);
            Real acc[SF_MAX_A];
            for (int tau = 0; tau < sf.n_a; ++tau) acc[tau] = Real(0);
            for (int p = seg[a]; p < seg[a + 1]; ++p) {
                const double *ep = row + p * ROW_W;
                const Real fp = (Real)ep[4], dfp = (Real)ep[5];
                if (fp == Real(0)) continue;
                const Real px = (Real)ep[0], py = (Real)ep[1], pz = (Real)ep[2],
                           r1 = (Real)ep[3];
                const Real ir1 = (Real)ep[7];
                const Real s1 = (px * (Real)dd[p * 4] + py * (Real)dd[p * 4 + 1] +
                                 pz * (Real)dd[p * 4 + 2]) * ir1;
                const int q0 = (a == b) ? p + 1 : seg[b];
                for (int q = q0 + lane; q < seg[b + 1]; q += SF_TPA) {
                    const double *eq = row + q * ROW_W;
                    const Real fq = (Real)eq[4], dfq = (Real)eq[5];
                    if (fq == Real(0)) continue;
                    const Real qx = (Real)eq[0], qy = (Real)eq[1], qz = (Real)eq[2],
                               r2 = (Real)eq[3];
                    const Real jx = qx - px, jy = qy - py, jz = qz - pz;
                    const Real d3 = jx * jx + jy * jy + jz * jz + Math<Real>::eps();
                    const Real ir3 = Math<Real>::rsqrt_(d3);
                    const Real r3 = d3 * ir3;
                    Real f3, df3;
                    cutoff_fn<Real>(sf.cutoff, r3, (Real)sf.acut, f3, df3);
                    if (f3 == Real(0)) continue;
                    const Real ir2 = (Real)eq[7];
                    const Real s2 = (qx * (Real)dd[q * 4] + qy * (Real)dd[q * 4 + 1] +
                                     qz * (Real)dd[q * 4 + 2]) * ir2;
                    const Real s3 = (jx * ((Real)dd[q * 4] - (Real)dd[p * 4]) +
                                     jy * ((Real)dd[q * 4 + 1] - (Real)dd[p * 4 + 1]) +
                                     jz * ((Real)dd[q * 4 + 2] - (Real)dd[p * 4 + 2])) * ir3;
                    const Real ss = r1 * r1 + r2 * r2 + r3 * r3;
                    const Real i12 = ir1 * ir2;
                    const bool ok = i12 != Real(0);
                    const Real ct = (r1 * r1 + r2 * r2 - r3 * r3) * (Real(0.5) * i12);
                    const Real dc1 = ok ? ir2 - ct * ir1 : Real(0);
                    const Real dc2 = ok ? ir1 - ct * ir2 : Real(0);
                    const Real dc3 = -r3 * i12;
                    const Real F3 = fp * fq * f3;
                    // directional derivative of the geometry-only factors
                    const Real dct = dc1 * s1 + dc2 * s2 + dc3 * s3;
                    const Real dss = Real(2) * (r1 * s1 + r2 * s2 + r3 * s3);
                    const Real dF3 = dfp * s1 * fq * f3 + fp * dfq * s2 * f3 + fp * fq * df3 * s3;
                    Real E = Real(0);
                    for (int tau = 0; tau < sf.n_a; ++tau) {
                        const Real z = (Real)sf.zeta[tau], gm = (Real)sf.gamma[tau],
                                   be = (Real)sf.beta[tau];
                        Real dP;
                        const Real P = powz<Real>(Real(1) + gm * ct, z, dP);
                        if (tau == 0 || sf.beta[tau] != sf.beta[tau - 1])
                            E = Math<Real>::exp_(-be * ss * ac2i);
                        acc[tau] += (Real)sf.outer[tau] *
                                    (dP * gm * dct * E * F3 - be * ac2i * dss * P * E * F3 +
                                     P * E * dF3);
                    }
                }
            }
            for (int tau = 0; tau < sf.n_a; ++tau) {
                const double tot = block_sum((double)acc[tau], red);
                if (lane == 0) t[sf.d_r + pt * sf.n_a + tau] = tot;
            }
        }
}

// caller order <-> sorted order row permutations
__global__ void k_rows_to_sorted(int n, int w, const int *__restrict__ perm,
                                 const double *__restrict__ src, double *__restrict__ dst) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)n * w) return;
    const int idx = (int)(t / w), k = (int)(t % w);
    dst[t] = src[(size_t)perm[idx] * w + k];
}

// sorted -> caller order copy of the descriptors
__global__ void k_sf_export(int n, int dim, const int *__restrict__ perm,
                            const double *__restrict__ G, double *__restrict__ out) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)n * dim) return;
    const int idx = (int)(t / dim), k = (int)(t % dim);
    out[(size_t)perm[idx] * dim + k] = G[t];
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
extern "C" int tab_atomic_create(tab_atomic **out, const tab_sf_desc *d,
                                 const tab_mlp_desc *mlps) {
    if (!out || !d || !mlps || d->n_el < 1 || d->n_el > TAB_MAX_ELEMENTS) {
        tab_set_error("tab_atomic_create: bad argument");
        return TAB_EINVAL;
    }
    if (d->n_r < 0 || d->n_r > SF_MAX_R || d->n_a < 0 || d->n_a > SF_MAX_A) {
        tab_set_error("tab_atomic_create: too many symmetry-function parameter sets");
        return TAB_EUNSUPPORTED;
    }
    tab_atomic *m = new tab_atomic();
    SfDev &sf = m->sf;
    memset(&sf, 0, sizeof(sf));
    sf.n_el = m->n_el = d->n_el;
    sf.n_r = d->n_r;
    sf.n_a = d->angular ? d->n_a : 0;
    sf.angular = d->angular ? 1 : 0;
    sf.cutoff = d->cutoff;
    sf.rc = d->rc;
    sf.acut = d->angular ? d->acut : d->rc;
    for (int k = 0; k < d->n_r; ++k) {
        sf.eta[k] = d->eta[k];
        sf.omega[k] = d->omega[k];
        sf.p3[k] = d->p3 ? d->p3[k] : 0.0;
    }
    sf.rad_kind = d->radial_kind;
    sf.n_mom = d->n_moments > 0 ? d->n_moments : 1;
    if (sf.rad_kind < 0 || sf.rad_kind > 3 || sf.n_mom > 4) {
        tab_set_error("tab_atomic_create: unknown radial family / moments");
        delete m;
        return TAB_EINVAL;
    }
    for (int k = 0; k < sf.n_mom; ++k) {
        sf.mom[k] = d->n_moments > 0 ? d->moments[k] : 0;
        if (sf.mom[k] < 0 || sf.mom[k] > 3) {
            tab_set_error("tab_atomic_create: moment %d is not supported (0 .. 3)", sf.mom[k]);
            delete m;
            return TAB_EUNSUPPORTED;
        }
        if (sf.mom[k] == 1) sf.has_m1 = 1;
        if (sf.mom[k] == 2) sf.has_m2 = 1;
        if (sf.mom[k] == 3) sf.has_m3 = 1;
    }
    sf.new_m0 = (d->grap_flags & TAB_GRAP_SIGNED_SQRT_M0) ? 1 : 0;
    sf.sym = (d->grap_flags & TAB_GRAP_TRACELESS) ? 1 : 0;
    if (sf.sym) {
        // the traceless entries subtract lower moments of the same list: they must be there
        bool ok = true;
        for (int k = 0; k < sf.n_mom; ++k) ok = ok && sf.mom[k] == k;
        if (!ok) {
            tab_set_error("tab_atomic_create: TAB_GRAP_TRACELESS needs moments 0..max");
            delete m;
            return TAB_EINVAL;
        }
    }
    for (int k = 0; k < sf.n_a; ++k) {
        sf.beta[k] = d->beta[k];
        sf.gamma[k] = d->gamma[k];
        sf.zeta[k] = d->zeta[k];
        sf.outer[k] = pow(2.0, 1.0 - d->zeta[k]);
    }
    sf.d_r = sf.n_el * sf.n_r * sf.n_mom;
    sf.dim = sf.d_r + (sf.angular ? sf.n_el * (sf.n_el + 1) / 2 * sf.n_a : 0);
    // pack the MLP blobs
    size_t total = 0;
    for (int e = 0; e < d->n_el; ++e) {
        const tab_mlp_desc &q = mlps[e];
        if (q.n_layers < 1 || q.n_layers > MLP_MAX_LAYERS || q.sizes[0] != sf.dim ||
            q.sizes[q.n_layers] != 1) {
            tab_set_error("tab_atomic_create: element %d: bad layer layout "
                          "(input must be %d wide, output 1)", e, sf.dim);
            delete m;
            return TAB_EINVAL;
        }
        for (int l = 0; l < q.n_layers; ++l) {
            if (q.sizes[l] > MLP_MAX_WIDTH || q.sizes[l + 1] > MLP_MAX_WIDTH) {
                tab_set_error("tab_atomic_create: layer wider than %d", MLP_MAX_WIDTH);
                delete m;
                return TAB_EUNSUPPORTED;
            }
            total += (size_t)q.sizes[l] * q.sizes[l + 1] + q.sizes[l + 1];
        }
        total += 2 * (size_t)sf.dim;
    }
    double *host = new double[total + TAB_MAX_ELEMENTS * sizeof(MlpDev) / 8 + 8];
    size_t off = 0;
    for (int e = 0; e < d->n_el; ++e) {
        const tab_mlp_desc &q = mlps[e];
        MlpDev &M = m->mlp[e];
        memset(&M, 0, sizeof(M));
        M.n_layers = q.n_layers;
        M.act = q.activation;
        M.resnet = q.use_resnet_dt;
        M.has_out_bias = q.output_bias;
        M.has_minmax = (q.xlo && q.xhi) ? 1 : 0;
        for (int l = 0; l < q.n_layers; ++l) {
            const int ni = q.sizes[l], no = q.sizes[l + 1];
            M.in[l] = ni;
            M.out[l] = no;
            M.w_off[l] = (long long)off;
            memcpy(host + off, q.weights[l], sizeof(double) * ni * no);
            off += (size_t)ni * no;
            M.b_off[l] = (long long)off;
            if (q.biases[l]) memcpy(host + off, q.biases[l], sizeof(double) * no);
            else memset(host + off, 0, sizeof(double) * no);
            off += no;
        }
        M.xlo_off = (long long)off;
        if (M.has_minmax) memcpy(host + off, q.xlo, sizeof(double) * sf.dim);
        off += sf.dim;
        M.xhi_off = (long long)off;
        if (M.has_minmax) memcpy(host + off, q.xhi, sizeof(double) * sf.dim);
        off += sf.dim;
    }
    // MlpDev table appended (8-byte aligned)
    const size_t mlp_off = off;
    memcpy(host + off, m->mlp, sizeof(MlpDev) * d->n_el);
    const size_t bytes = off * 8 + sizeof(MlpDev) * d->n_el;
    int rc = m->blob.ensure(bytes);
    if (rc == TAB_OK) {
        cudaError_t e = cudaMemcpy(m->blob.p, host, bytes, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) {
            tab_set_error("tab_atomic_create: cudaMemcpy -> %s", cudaGetErrorString(e));
            rc = TAB_ECUDA;
        }
    }
    delete[] host;
    if (rc != TAB_OK) {
        m->blob.release();
        delete m;
        return rc;
    }
    m->mlp_table_off = mlp_off;
    // tensor-core variant (mlp_tc.cuh): two hidden layers, D <= 64, H1 <= 64, H2 <= 32,
    // no ResNet link
    {
        MlpTcDev nets[TAB_MAX_ELEMENTS];
        bool ok = true;
        auto pad16 = [](int v) { return (v + 15) & ~15; };
        for (int e = 0; e < m->n_el && ok; ++e) {
            const MlpDev &q = m->mlp[e];
            if (q.n_layers != 3 || q.resnet || q.in[0] > 64 || q.out[0] > 64 || q.out[1] > 32 ||
                q.out[2] != 1) {
                ok = false;
                break;
            }
            MlpTcDev &t = nets[e];
            t.dim = q.in[0];
            t.dp = pad16(q.in[0]);
            t.h1 = q.out[0];
            t.h1p = pad16(q.out[0]);
            t.h2 = q.out[1];
            t.h2p = pad16(q.out[1]);
            t.act = q.act;
            t.has_out_bias = q.has_out_bias;
            t.has_minmax = q.has_minmax;
            t.w1 = q.w_off[0];
            t.b1 = q.b_off[0];
            t.w2 = q.w_off[1];
            t.b2 = q.b_off[1];
            t.w3 = q.w_off[2];
            t.b3 = q.b_off[2];
            t.xlo = q.xlo_off;
            t.xhi = q.xhi_off;
        }
        if (ok && m->tc_nets.ensure(sizeof(MlpTcDev) * m->n_el) == TAB_OK &&
            m->tc_status.ensure(sizeof(int)) == TAB_OK &&
            cudaMemcpy(m->tc_nets.p, nets, sizeof(MlpTcDev) * m->n_el,
                       cudaMemcpyHostToDevice) == cudaSuccess &&
            cudaMemset(m->tc_status.p, 0, sizeof(int)) == cudaSuccess)
            m->tc_ok = true;
    }
    *out = m;
    return TAB_OK;
}

extern "C" int tab_atomic_free(tab_atomic *m) {
    if (!m) return TAB_OK;
    DevBuf *bufs[] = {&m->blob, &m->G, &m->dEdG, &m->eat, &m->gvec, &m->fown, &m->tc_nets,
                      &m->tc_status, &m->mom};
    for (DevBuf *b : bufs) b->release();
    delete m;
    return TAB_OK;
}

extern "C" int tab_atomic_dim(const tab_atomic *m) { return m ? m->sf.dim : 0; }

// the geometry kernels exist in two widths of the moment-sum record: 10 (moments <= 2) and
// 20 (moment 3); `sf` and `Real` are taken from the calling scope
#define SF_MOM_W(sf) ((sf).has_m3 ? GRAP_MOM_W3 : GRAP_MOM_W)
#define SF_LAUNCH(kern, grid, smem, st, ...)                                              \
    do {                                                                                  \
        if (sf.has_m3)                                                                    \
            kern<Real, GRAP_MOM_W3><<<grid, SF_WARPS * 32, smem, st>>>(__VA_ARGS__);      \
        else                                                                              \
            kern<Real, GRAP_MOM_W><<<grid, SF_WARPS * 32, smem, st>>>(__VA_ARGS__);       \
    } while (0)
#define SF_SMEM_ATTR(kern)                                                                \
    do {                                                                                  \
        TAB_CUDA(cudaFuncSetAttribute(kern<Real, GRAP_MOM_W>,                             \
                                      cudaFuncAttributeMaxDynamicSharedMemorySize,        \
                                      220 * 1024));                                       \
        TAB_CUDA(cudaFuncSetAttribute(kern<Real, GRAP_MOM_W3>,                            \
                                      cudaFuncAttributeMaxDynamicSharedMemorySize,        \
                                      220 * 1024));                                       \
    } while (0)

template <typename Real>
static int atomic_run(tab_atomic *m, tab_nbr *nbr, double *d_energy, double *d_eatom,
                      double *d_forces, double *d_virial, double *d_desc,
                      cudaStream_t st, const int *d_mask = nullptr) {
    const int n = nbr->n;
    const SfDev &sf = m->sf;
    if (nbr->n_halo > 0 && !d_mask && !d_desc) {
        tab_set_error("AtomicNN on lists with halo atoms: use tab_atomic_eval_dd "
                      "(the rows of the inner halo must be masked out of the sums)");
        return TAB_EUNSUPPORTED;
    }
    if (d_mask && nbr->n_struct > 0) {
        tab_set_error("tab_atomic_eval_dd: batch handles are not decomposed");
        return TAB_EUNSUPPORTED;
    }
    if (nbr->n_types > sf.n_el) {
        tab_set_error("structure holds element index %d, model has %d elements",
                      nbr->n_types - 1, sf.n_el);
        return TAB_EINVAL;
    }
    const double need_rc = sf.angular && sf.acut > sf.rc ? sf.acut : sf.rc;
    if (nbr->grid.rc + 1e-12 < need_rc) {
        tab_set_error("neighbour lists were built with rc=%g < model cutoff %g",
                      nbr->grid.rc, need_rc);
        return TAB_ESTATE;
    }
    const int row_cap = nbr->nnl_max > 0 ? nbr->nnl_max : 1;
    const size_t smem_row = (size_t)row_cap * ROW_W * sizeof(double);
    if (smem_row > 200 * 1024) {
        tab_set_error("neighbour rows of %d entries exceed the shared-memory budget", row_cap);
        return TAB_EUNSUPPORTED;
    }
    const size_t smem_mlp = (size_t)MLP_WARPS * (2 * MLP_MAX_LAYERS + 2) * MLP_MAX_WIDTH *
                            sizeof(double);
    static bool attr_done[2] = {false, false};
    const int ai = sizeof(Real) == 8 ? 0 : 1;
    if (!attr_done[ai]) {
        SF_SMEM_ATTR(k_sf_forward);
        SF_SMEM_ATTR(k_sf_backward);
        TAB_CUDA(cudaFuncSetAttribute(k_mlp<Real>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        attr_done[ai] = true;
    }
    const int nblk = n;                       // one block per atom
    const int nblk_m = (n + MLP_WARPS - 1) / MLP_WARPS;
    const int nblk_c = (n + 3) / 4;
    TAB_TRY(m->G.ensure(sizeof(double) * (size_t)n * sf.dim));
    TAB_TRY(m->dEdG.ensure(sizeof(double) * (size_t)n * sf.dim));
    TAB_TRY(m->eat.ensure(sizeof(double) * (size_t)n));
    TAB_TRY(m->fown.ensure(sizeof(double) * 3 * (size_t)n));
    TAB_TRY(nbr->partial.ensure(sizeof(double) * (8 * (size_t)nblk + nblk_c + 8)));
    const Atom4 *atoms = nbr->atoms.as<Atom4>();
    const bool grap_moments = grap_need_mom(sf);
    if (grap_moments)
        TAB_TRY(m->mom.ensure(sizeof(double) * (size_t)n * sf.n_el * sf.n_r * SF_MOM_W(sf)));
    SF_LAUNCH(k_sf_forward, nblk, smem_row, st,
        n, sf, nbr->n_types, row_cap, atoms, nbr->types_ext.as<uint8_t>(),
        nbr->counts.as<int>(), nbr->tcounts.as<int>(), nbr->slice_ptr.as<uint32_t>(),
        nbr->col.as<uint32_t>(), m->G.as<double>(),
        grap_moments ? m->mom.as<double>() : nullptr);
    TAB_LAUNCH_CHECK();
    if (d_desc) {
        const size_t tot = (size_t)n * sf.dim;
        k_sf_export<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(
            n, sf.dim, nbr->perm.as<int>(), m->G.as<double>(), d_desc);
        TAB_LAUNCH_CHECK();
        if (!d_energy && !d_eatom && !d_forces && !d_virial) return TAB_OK;
    }
    const double *blob = m->blob.as<double>();
    const MlpDev *mlps = reinterpret_cast<const MlpDev *>(blob + m->mlp_table_off);
    // 'medium' precision, enough atoms to fill tiles: the MLP runs on the tensor cores
    // (TAB_MLP_TC=0 forces the warp-per-atom kernel, TAB_MLP_TC=1 forces the tensor
    // cores for any size, for tests)
    bool use_tc = sizeof(Real) == 4 && m->tc_ok && n >= 4096;
    if (const char *env = getenv("TAB_MLP_TC")) {
        if (env[0] == '0') use_tc = false;
        if (env[0] == '1') use_tc = sizeof(Real) == 4 && m->tc_ok;
    }
    if (use_tc) {
        static bool tc_attr = false;
        if (!tc_attr) {
            TAB_CUDA(cudaFuncSetAttribute(k_mlp_tc, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          TC_SMEM_BYTES));
            tc_attr = true;
        }
        const int n_tiles = (n + TC_ROWS - 1) / TC_ROWS;
        dim3 grid((unsigned)(n_tiles < 148 ? n_tiles : 148), (unsigned)m->n_el);
        k_mlp_tc<<<grid, TC_ROWS, TC_SMEM_BYTES, st>>>(
            n, sf.dim, nbr->types_ext.as<uint8_t>(), m->tc_nets.as<MlpTcDev>(), blob,
            m->G.as<double>(), m->eat.as<double>(), m->dEdG.as<double>(),
            m->tc_status.as<int>());
    } else {
        k_mlp<Real><<<nblk_m, MLP_WARPS * 32, smem_mlp, st>>>(
            n, sf.dim, nbr->types_ext.as<uint8_t>(), mlps, blob, m->G.as<double>(),
            m->eat.as<double>(), m->dEdG.as<double>());
    }
    TAB_LAUNCH_CHECK();
    const size_t plane = (size_t)nbr->ell_rows * 32;
    double *partial = nbr->partial.as<double>();
    double *partial_e = partial + 8 * (size_t)nblk;
    const bool need_grad = d_forces || d_virial;
    if (need_grad) {
        TAB_TRY(tab_nbr_ensure_reverse(nbr, st));
        TAB_TRY(m->gvec.ensure(sizeof(double) * 3 * (plane + 32)));
        SF_LAUNCH(k_sf_backward, nblk, smem_row, st,
            n, sf, nbr->n_types, row_cap, atoms, nbr->types_ext.as<uint8_t>(),
            nbr->counts.as<int>(), nbr->tcounts.as<int>(), nbr->slice_ptr.as<uint32_t>(),
            nbr->col.as<uint32_t>(), m->dEdG.as<double>(), m->gvec.as<double>(), plane,
            m->fown.as<double>(), partial, grap_moments ? m->mom.as<double>() : nullptr);
        TAB_LAUNCH_CHECK();
    }
    k_sf_collect<<<nblk_c, 128, 0, st>>>(
        n, nbr->n_loc, nbr->counts.as<int>(), nbr->slice_ptr.as<uint32_t>(),
        nbr->col.as<uint32_t>(), need_grad ? nbr->rev.as<uint32_t>() : nullptr,
        nbr->ghost_owner.as<int>(), nbr->perm.as<int>(), m->gvec.as<double>(), plane,
        m->fown.as<double>(), m->eat.as<double>(), d_eatom, need_grad ? d_forces : nullptr,
        partial_e);
    TAB_LAUNCH_CHECK();
    if (d_mask)
        k_sf_reduce_masked<<<1, 256, 0, st>>>(n, nbr->perm.as<int>(), d_mask,
                                             m->eat.as<double>(), need_grad ? partial : nullptr,
                                             d_energy, need_grad ? d_virial : nullptr);
    else if (nbr->n_struct > 0)
        k_sf_reduce_batch<<<nbr->n_struct, 128, 0, st>>>(
            nbr->struct_off.as<int>(), d_energy ? m->eat.as<double>() : nullptr,
            need_grad ? partial : nullptr, d_energy, need_grad ? d_virial : nullptr);
    else
        k_sf_reduce<<<1, 256, 0, st>>>(nblk_c, partial_e, need_grad ? nblk : 0, partial,
                                      d_energy, need_grad ? d_virial : nullptr);
    TAB_LAUNCH_CHECK();
    return TAB_OK;
}

extern "C" int tab_atomic_eval(tab_atomic *m, tab_nbr *nbr, int32_t precision,
                               double *d_energy, double *d_eatom, double *d_forces,
                               double *d_virial, void *stream) {
    if (!m || !nbr) {
        tab_set_error("tab_atomic_eval: null handle");
        return TAB_EINVAL;
    }
    if (!nbr->built) {
        tab_set_error("tab_atomic_eval before tab_nbr_build");
        return TAB_ESTATE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (precision == TAB_PRECISION_HIGH)
        return atomic_run<double>(m, nbr, d_energy, d_eatom, d_forces, d_virial, nullptr, st);
    return atomic_run<float>(m, nbr, d_energy, d_eatom, d_forces, d_virial, nullptr, st);
}

extern "C" int tab_atomic_eval_dd(tab_atomic *m, tab_nbr *nbr, int32_t precision,
                                  const int32_t *d_mask, double *d_energy, double *d_eatom,
                                  double *d_forces, double *d_virial, void *stream) {
    if (!m || !nbr || !d_mask) {
        tab_set_error("tab_atomic_eval_dd: null argument");
        return TAB_EINVAL;
    }
    if (!nbr->built) {
        tab_set_error("tab_atomic_eval_dd before tab_nbr_build_dd");
        return TAB_ESTATE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (precision == TAB_PRECISION_HIGH)
        return atomic_run<double>(m, nbr, d_energy, d_eatom, d_forces, d_virial, nullptr, st,
                                  d_mask);
    return atomic_run<float>(m, nbr, d_energy, d_eatom, d_forces, d_virial, nullptr, st, d_mask);
}

extern "C" int tab_atomic_descriptors(tab_atomic *m, tab_nbr *nbr, int32_t precision,
                                      double *d_desc, void *stream) {
    if (!m || !nbr || !d_desc) return TAB_EINVAL;
    if (!nbr->built) return TAB_ESTATE;
    cudaStream_t st = (cudaStream_t)stream;
    if (precision == TAB_PRECISION_HIGH)
        return atomic_run<double>(m, nbr, nullptr, nullptr, nullptr, nullptr, d_desc, st);
    return atomic_run<float>(m, nbr, nullptr, nullptr, nullptr, nullptr, d_desc, st);
}

// ---------------------------------------------------------------------------
// training entry points: force/virial as a linear operator of c = dE/dG
// ---------------------------------------------------------------------------
static int atomic_common_checks(tab_atomic *m, tab_nbr *nbr, const char *who) {
    if (!m || !nbr) {
        tab_set_error("%s: null handle", who);
        return TAB_EINVAL;
    }
    if (!nbr->built) {
        tab_set_error("%s before tab_nbr_build", who);
        return TAB_ESTATE;
    }
    if (nbr->n_halo > 0) {
        tab_set_error("%s: halo atoms are not supported", who);
        return TAB_EUNSUPPORTED;
    }
    return TAB_OK;
}

// GRAP moments 1 / 2: the backward and JVP kernels need the moment sums M, Q of the forward
// pass for THESE lists (m->mom is per model, not per structure): recompute them.
template <typename Real>
static int forward_moments(tab_atomic *m, tab_nbr *nbr, cudaStream_t st) {
    const SfDev &sf = m->sf;
    const int n = nbr->n;
    const int row_cap = nbr->nnl_max > 0 ? nbr->nnl_max : 1;
    const size_t smem_row = (size_t)row_cap * ROW_W * sizeof(double);
    if (smem_row > 200 * 1024) {
        tab_set_error("neighbour rows of %d entries exceed the shared-memory budget", row_cap);
        return TAB_EUNSUPPORTED;
    }
    SF_SMEM_ATTR(k_sf_forward);
    TAB_TRY(m->G.ensure(sizeof(double) * (size_t)n * sf.dim));
    TAB_TRY(m->mom.ensure(sizeof(double) * (size_t)n * sf.n_el * sf.n_r * SF_MOM_W(sf)));
    SF_LAUNCH(k_sf_forward, n, smem_row, st,
        n, sf, nbr->n_types, row_cap, nbr->atoms.as<Atom4>(), nbr->types_ext.as<uint8_t>(),
        nbr->counts.as<int>(), nbr->tcounts.as<int>(), nbr->slice_ptr.as<uint32_t>(),
        nbr->col.as<uint32_t>(), m->G.as<double>(), m->mom.as<double>());
    TAB_LAUNCH_CHECK();
    return TAB_OK;
}

template <typename Real>
static int atomic_forces_from(tab_atomic *m, tab_nbr *nbr, const double *d_dedg,
                              double *d_forces, double *d_virial, cudaStream_t st) {
    const int n = nbr->n;
    const SfDev &sf = m->sf;
    const int row_cap = nbr->nnl_max > 0 ? nbr->nnl_max : 1;
    const size_t smem_row = (size_t)row_cap * ROW_W * sizeof(double);
    SF_SMEM_ATTR(k_sf_backward);
    const int nblk = n, nblk_c = (n + 3) / 4;
    TAB_TRY(m->dEdG.ensure(sizeof(double) * (size_t)n * sf.dim));
    TAB_TRY(m->eat.ensure(sizeof(double) * (size_t)n));
    TAB_TRY(m->fown.ensure(sizeof(double) * 3 * (size_t)n));
    TAB_TRY(nbr->partial.ensure(sizeof(double) * (8 * (size_t)nblk + nblk_c + 8)));
    TAB_TRY(tab_nbr_ensure_reverse(nbr, st));
    const size_t plane = (size_t)nbr->ell_rows * 32;
    TAB_TRY(m->gvec.ensure(sizeof(double) * 3 * (plane + 32)));
    const bool grap_moments = grap_need_mom(sf);
    if (grap_moments) TAB_TRY(forward_moments<Real>(m, nbr, st));
    const size_t tot = (size_t)n * sf.dim;
    k_rows_to_sorted<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(
        n, sf.dim, nbr->perm.as<int>(), d_dedg, m->dEdG.as<double>());
    TAB_LAUNCH_CHECK();
    TAB_CUDA(cudaMemsetAsync(m->eat.p, 0, sizeof(double) * (size_t)n, st));
    double *partial = nbr->partial.as<double>();
    SF_LAUNCH(k_sf_backward, nblk, smem_row, st,
        n, sf, nbr->n_types, row_cap, nbr->atoms.as<Atom4>(), nbr->types_ext.as<uint8_t>(),
        nbr->counts.as<int>(), nbr->tcounts.as<int>(), nbr->slice_ptr.as<uint32_t>(),
        nbr->col.as<uint32_t>(), m->dEdG.as<double>(), m->gvec.as<double>(), plane,
        m->fown.as<double>(), partial, grap_moments ? m->mom.as<double>() : nullptr);
    TAB_LAUNCH_CHECK();
    k_sf_collect<<<nblk_c, 128, 0, st>>>(
        n, nbr->n_loc, nbr->counts.as<int>(), nbr->slice_ptr.as<uint32_t>(),
        nbr->col.as<uint32_t>(), nbr->rev.as<uint32_t>(), nbr->ghost_owner.as<int>(),
        nbr->perm.as<int>(), m->gvec.as<double>(), plane, m->fown.as<double>(),
        m->eat.as<double>(), nullptr, d_forces, partial + 8 * (size_t)nblk);
    TAB_LAUNCH_CHECK();
    if (nbr->n_struct > 0)
        k_sf_reduce_batch<<<nbr->n_struct, 128, 0, st>>>(nbr->struct_off.as<int>(), nullptr,
                                                         partial, nullptr, d_virial);
    else
        k_sf_reduce<<<1, 256, 0, st>>>(0, nullptr, nblk, partial, nullptr, d_virial);
    TAB_LAUNCH_CHECK();
    return TAB_OK;
}

extern "C" int tab_atomic_forces(tab_atomic *m, tab_nbr *nbr, int32_t precision,
                                 const double *d_dedg, double *d_forces,
                                 double *d_virial, void *stream) {
    TAB_TRY(atomic_common_checks(m, nbr, "tab_atomic_forces"));
    if (!d_dedg) return TAB_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    if (precision == TAB_PRECISION_HIGH)
        return atomic_forces_from<double>(m, nbr, d_dedg, d_forces, d_virial, st);
    return atomic_forces_from<float>(m, nbr, d_dedg, d_forces, d_virial, st);
}

template <typename Real>
static int atomic_jvp(tab_atomic *m, tab_nbr *nbr, const double *d_u, const double *d_A,
                      double *d_out, cudaStream_t st) {
    const int n = nbr->n;
    const SfDev &sf = m->sf;
    const int row_cap = nbr->nnl_max > 0 ? nbr->nnl_max : 1;
    const size_t smem = (size_t)row_cap * (ROW_W + 4) * sizeof(double);
    if (smem > 200 * 1024) {
        tab_set_error("neighbour rows of %d entries exceed the shared-memory budget", row_cap);
        return TAB_EUNSUPPORTED;
    }
    SF_SMEM_ATTR(k_sf_jvp);
    const int nblk = n;
    const bool grap_moments = grap_need_mom(sf);
    if (grap_moments) TAB_TRY(forward_moments<Real>(m, nbr, st));   // before G is reused as T
    TAB_TRY(m->fown.ensure(sizeof(double) * 3 * (size_t)n));
    TAB_TRY(m->G.ensure(sizeof(double) * (size_t)n * sf.dim));
    // u to sorted order
    k_rows_to_sorted<<<(unsigned)(((size_t)n * 3 + 255) / 256), 256, 0, st>>>(
        n, 3, nbr->perm.as<int>(), d_u, m->fown.as<double>());
    TAB_LAUNCH_CHECK();
    SF_LAUNCH(k_sf_jvp, nblk, smem, st,
        n, nbr->n_loc, sf, nbr->n_types, row_cap, nbr->atoms.as<Atom4>(),
        nbr->types_ext.as<uint8_t>(), nbr->counts.as<int>(), nbr->tcounts.as<int>(),
        nbr->slice_ptr.as<uint32_t>(), nbr->col.as<uint32_t>(), nbr->ghost_owner.as<int>(),
        m->fown.as<double>(), d_A, nbr->n_struct > 0 ? nbr->struct_of.as<int>() : nullptr,
        grap_moments ? m->mom.as<double>() : nullptr, m->G.as<double>());
    TAB_LAUNCH_CHECK();
    const size_t tot = (size_t)n * sf.dim;
    k_sf_export<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(
        n, sf.dim, nbr->perm.as<int>(), m->G.as<double>(), d_out);
    TAB_LAUNCH_CHECK();
    return TAB_OK;
}

extern "C" int tab_atomic_jvp(tab_atomic *m, tab_nbr *nbr, int32_t precision,
                              const double *d_u, const double *d_A, double *d_out,
                              void *stream) {
    TAB_TRY(atomic_common_checks(m, nbr, "tab_atomic_jvp"));
    if (!d_u || !d_A || !d_out) return TAB_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    if (precision == TAB_PRECISION_HIGH) return atomic_jvp<double>(m, nbr, d_u, d_A, d_out, st);
    return atomic_jvp<float>(m, nbr, d_u, d_A, d_out, st);
}
