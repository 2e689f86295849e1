// potentials.cuh -- device restatement of the analytic potential functions of
// the reference (tensoralloy/nn/eam/potentials/*), value AND first derivative.
// Each evaluator cites the reference lines it follows.  `Real` is the working
// precision of the pair arithmetic (double = 'high', float = 'medium').
#pragma once
#include "tab_internal.h"

// ---------------------------------------------------------------------------
// arithmetic building blocks.  float64: MUFU seed + Newton iterations and a
// branch-free exp (arguments here are bounded: no overflow / denormal / NaN
// paths needed); every routine is accurate to <= 2 ulp, three orders below the
// 1e-10 eV/atom parity tolerance.  float32: hardware approximations (the
// 'medium' tolerance is 1e-5 relative).
// ---------------------------------------------------------------------------
// MUFU seeds carry >= 20 good bits (relative error e <= 2^-20); ONE third-order
// step leaves e^3 ~ 1e-18, so a second Newton iteration is not needed.
__device__ __forceinline__ double tab_rcp(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-x, y, 1.0);          // 1/x = y / (1 - e) = y (1 + e + e^2 + ...)
    return fma(y, fma(e, e, e), y);
}

// 1/sqrt(x) for normal positive x
__device__ __forceinline__ double tab_rsqrt(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-x, y * y, 1.0);      // x^-1/2 = y (1 - e)^-1/2 = y (1 + e/2 + 3e^2/8 + ...)
    return fma(y, e * fma(0.375, e, 0.5), y);
}

// exp(t) for t < 700 (t < -700 -> 0): n = rint(t log2 e), f = t - n ln2, degree-12 Taylor
// polynomial on |f| <= 0.347 (truncation 1.7e-16), exponent patched in.
__device__ __forceinline__ double tab_exp(double t) {
    const double SHIFT = 6755399441055744.0;   // 1.5 * 2^52
    const double z = fma(t, 1.4426950408889634074, SHIFT);
    const int n = __double2loint(z);
    const double nf = z - SHIFT;
    double f = fma(nf, -6.93147180369123816490e-01, t);
    f = fma(nf, -1.90821492927058770002e-10, f);
    double p = 2.08767569878680989792e-09;      // 1/12!
    p = fma(p, f, 2.50521083854417187751e-08);  // 1/11!
    p = fma(p, f, 2.75573192239858906526e-07);  // 1/10!
    p = fma(p, f, 2.75573192239858906526e-06);  // 1/9!
    p = fma(p, f, 2.48015873015873015873e-05);  // 1/8!
    p = fma(p, f, 1.98412698412698412698e-04);  // 1/7!
    p = fma(p, f, 1.38888888888888888889e-03);  // 1/6!
    p = fma(p, f, 8.33333333333333333333e-03);  // 1/5!
    p = fma(p, f, 4.16666666666666666667e-02);  // 1/4!
    p = fma(p, f, 1.66666666666666666667e-01);  // 1/3!
    p = fma(p, f, 0.5);
    p = fma(p, f, 1.0);
    p = fma(p, f, 1.0);
    const double y = __hiloint2double(__double2hiint(p) + (n << 20), __double2loint(p));
    return t < -700.0 ? 0.0 : y;     // underflow guard (erf tails, far pairs)
}

template <typename Real> struct Math;
template <> struct Math<double> {
    static __device__ __forceinline__ double exp_(double x) { return tab_exp(x); }
    static __device__ __forceinline__ double log_(double x) { return log(x); }
    static __device__ __forceinline__ double pow_(double x, double y) { return pow(x, y); }
    static __device__ __forceinline__ double rcp_(double x) { return tab_rcp(x); }
    static __device__ __forceinline__ double rsqrt_(double x) { return tab_rsqrt(x); }
    static __device__ __forceinline__ double sqrt_(double x) { return sqrt(x); }
    static __device__ __forceinline__ double erf_(double x) { return erf(x); }
    static __device__ __forceinline__ double eps() { return 1e-14; }   // precision.py:113
};
template <> struct Math<float> {
    static __device__ __forceinline__ float exp_(float x) { return __expf(x); }
    static __device__ __forceinline__ float log_(float x) { return logf(x); }
    static __device__ __forceinline__ float pow_(float x, float y) { return powf(x, y); }
    static __device__ __forceinline__ float rcp_(float x) { return __frcp_rn(x); }
    static __device__ __forceinline__ float rsqrt_(float x) { return rsqrtf(x); }
    static __device__ __forceinline__ float sqrt_(float x) { return sqrtf(x); }
    static __device__ __forceinline__ float erf_(float x) { return erff(x); }
    static __device__ __forceinline__ float eps() { return 1e-8f; }    // precision.py:114
};

// generic.py:102-117  f(r) = a exp(-b (r/re - 1)) / (1 + (r/re - c)^20)
// returns f and df/dr.  `inv_re` = 1/re.  (x-c)^20 is formed by repeated
// squaring; the reference calls pow(x-c, 20.0); the two agree to a few ulp.
template <typename Real>
__device__ __forceinline__ void zhou_exp(Real r, Real a, Real b, Real c, Real inv_re,
                                         Real &f, Real &df) {
    const Real x = r * inv_re;
    const Real u = x - c;
    const Real u2 = u * u, u4 = u2 * u2, u8 = u4 * u4, u16 = u8 * u8;
    const Real u19 = u16 * u2 * u;
    const Real q = Math<Real>::rcp_(fma(u19, u, Real(1)));
    const Real e = a * Math<Real>::exp_(fma(-b, x, b));
    f = e * q;
    df = -(f * inv_re) * fma(Real(20) * u19, q, b);
}

// ---------------------------------------------------------------------------
// Folded form of one zhou_exp term for the single-element float64 fast path:
//   f(r) = a exp(-b (x - 1)) / (1 + (x - c)^20),  x = r / re
// with the prefactor moved into the exponent (t = -b x + (b + ln a)), the
// denominator formed as 1 + u^16 u^4, and (20/re) u^19 built from
// u' = (20/re) u = r (20/re^2) - 20 c/re, so that
//   df/dr = f * ( -(20/re) u^19 q - b/re ).
// 25 FP64 instructions for f and df (the generic zhou_exp above: 35); the
// exponential is table driven (below) and has no underflow guard -- the host only
// takes this path when b (rc/re - 1) < 600.
// ---------------------------------------------------------------------------
struct ZTerm {
    double nb;      // -b
    double c;       // b + ln a
    double kappa;   // c of the formula
    double c20;     // 20 / re^2
    double k20;     // -20 c / re
    double nb_re;   // -b / re
};

// 2^(j/64), j = 0..63, correctly rounded (exp2 table of the folded float64 terms).
// The kernels copy a stride of it to shared memory once per block.
#define TAB_EXP_TAB 64
__device__ const double c_exp2_tab[TAB_EXP_TAB] = {
    0x1.0000000000000p+0, 0x1.02c9a3e778061p+0, 0x1.059b0d3158574p+0, 0x1.0874518759bc8p+0,
    0x1.0b5586cf9890fp+0, 0x1.0e3ec32d3d1a2p+0, 0x1.11301d0125b51p+0, 0x1.1429aaea92de0p+0,
    0x1.172b83c7d517bp+0, 0x1.1a35beb6fcb75p+0, 0x1.1d4873168b9aap+0, 0x1.2063b88628cd6p+0,
    0x1.2387a6e756238p+0, 0x1.26b4565e27cddp+0, 0x1.29e9df51fdee1p+0, 0x1.2d285a6e4030bp+0,
    0x1.306fe0a31b715p+0, 0x1.33c08b26416ffp+0, 0x1.371a7373aa9cbp+0, 0x1.3a7db34e59ff7p+0,
    0x1.3dea64c123422p+0, 0x1.4160a21f72e2ap+0, 0x1.44e086061892dp+0, 0x1.486a2b5c13cd0p+0,
    0x1.4bfdad5362a27p+0, 0x1.4f9b2769d2ca7p+0, 0x1.5342b569d4f82p+0, 0x1.56f4736b527dap+0,
    0x1.5ab07dd485429p+0, 0x1.5e76f15ad2148p+0, 0x1.6247eb03a5585p+0, 0x1.6623882552225p+0,
    0x1.6a09e667f3bcdp+0, 0x1.6dfb23c651a2fp+0, 0x1.71f75e8ec5f74p+0, 0x1.75feb564267c9p+0,
    0x1.7a11473eb0187p+0, 0x1.7e2f336cf4e62p+0, 0x1.82589994cce13p+0, 0x1.868d99b4492edp+0,
    0x1.8ace5422aa0dbp+0, 0x1.8f1ae99157736p+0, 0x1.93737b0cdc5e5p+0, 0x1.97d829fde4e50p+0,
    0x1.9c49182a3f090p+0, 0x1.a0c667b5de565p+0, 0x1.a5503b23e255dp+0, 0x1.a9e6b5579fdbfp+0,
    0x1.ae89f995ad3adp+0, 0x1.b33a2b84f15fbp+0, 0x1.b7f76f2fb5e47p+0, 0x1.bcc1e904bc1d2p+0,
    0x1.c199bdd85529cp+0, 0x1.c67f12e57d14bp+0, 0x1.cb720dcef9069p+0, 0x1.d072d4a07897cp+0,
    0x1.d5818dcfba487p+0, 0x1.da9e603db3285p+0, 0x1.dfc97337b9b5fp+0, 0x1.e502ee78b3ff6p+0,
    0x1.ea4afa2a490dap+0, 0x1.efa1bee615a27p+0, 0x1.f50765b6e4540p+0, 0x1.fa7c1819e90d8p+0,
};

// LB = log2(table entries): 0 = no table (degree-11 polynomial on |fr| <= ln2/2),
// 4 = 16 entries (one 128-byte shared-memory row: conflict-free for any index pattern,
// degree 7), 5 = 32 entries (degree 6), 6 = 64 entries (degree 5).
template <int LB>
__device__ __forceinline__ void load_exp2_tab(double *s_tab) {
    if (LB > 0) {
        if (threadIdx.x < (1 << LB)) s_tab[threadIdx.x] = c_exp2_tab[threadIdx.x << (6 - LB)];
        __syncthreads();
    }
}

// exp(t) of the folded terms.  Table driven (LB > 0): t = (2^LB m + j) ln2/2^LB + fr,
// |fr| <= ln2/2^(LB+1), exp(t) = 2^m * 2^(j/2^LB) * P(fr) with a Taylor polynomial whose
// truncation error is < 2e-17 relative, the table entry by one 8-byte shared-memory load.
template <int LB>
__device__ __forceinline__ double zexp(double t, const double *__restrict__ etab) {
    const double SHIFT = 6755399441055744.0;   // 1.5 * 2^52
    constexpr double S = (double)(1 << LB);
    const double zz = fma(t, S * 1.4426950408889634074, SHIFT);
    const int n = __double2loint(zz);
    const double nf = zz - SHIFT;
    double fr = fma(nf, -6.93147180369123816490e-01 / S, t);
    fr = fma(nf, -1.90821492927058770002e-10 / S, fr);
    double e;
    if (LB == 0) {
        e = 2.511015364692893e-08;
        e = fma(e, fr, 2.763279054211718e-07);
        e = fma(e, fr, 2.7557240604663896e-06);
        e = fma(e, fr, 2.4801485074105115e-05);
        e = fma(e, fr, 0.00019841269890340666);
        e = fma(e, fr, 0.0013888888952696516);
        e = fma(e, fr, 0.00833333333331949);
        e = fma(e, fr, 0.04166666666648666);
        e = fma(e, fr, 0.1666666666666668);
        e = fma(e, fr, 0.5000000000000019);
        e = fma(e, fr, 1.0);
        e = fma(e, fr, 1.0);   // (an Estrin split of this chain measured 4 % slower: +3 FP64 ops)
    } else {
        if (LB == 4) {
            e = 1.98412698412698412698e-04;                 // 1/7!
            e = fma(e, fr, 1.38888888888888888889e-03);     // 1/6!
            e = fma(e, fr, 8.33333333333333333333e-03);     // 1/5!
        } else if (LB == 5) {
            e = 1.38888888888888888889e-03;
            e = fma(e, fr, 8.33333333333333333333e-03);
        } else {
            e = 8.33333333333333333333e-03;
        }
        e = fma(e, fr, 4.16666666666666666667e-02);
        e = fma(e, fr, 1.66666666666666666667e-01);
        e = fma(e, fr, 0.5);
        e = fma(e, fr, 1.0);
        e = fma(e, fr, 1.0);
        e *= etab[n & ((1 << LB) - 1)];
    }
    return __hiloint2double(__double2hiint(e) + ((n >> LB) << 20), __double2loint(e));
}

template <bool DERIV, int LB>
__device__ __forceinline__ void zterm_eval(double r, double x, const ZTerm &p,
                                           const double *__restrict__ etab, double &f,
                                           double &df) {
    const double u = x - p.kappa;
    const double u2 = u * u, u4 = u2 * u2, u8 = u4 * u4, u16 = u8 * u8;
    const double q = tab_rcp(fma(u16, u4, 1.0));
    const double e = zexp<LB>(fma(x, p.nb, p.c), etab);
    f = e * q;
    if (DERIV) {
        const double up = fma(r, p.c20, p.k20);
        const double u19p = (u16 * u2) * up;
        df = f * fma(-u19p, q, p.nb_re);
    }
}

// zjw04.py:279-389 (piecewise) and :440-550 (xc, sigmoid blended).
// p = Fn0..Fn3, F0..F3, eta, Fe, rho_e, rho_s
template <typename Real>
__device__ __forceinline__ void zhou_embed(const double *p, bool blended, Real rho,
                                           Real &F, Real &dF) {
    const Real Fn0 = (Real)p[0], Fn1 = (Real)p[1], Fn2 = (Real)p[2], Fn3 = (Real)p[3];
    const Real F0 = (Real)p[4], F1 = (Real)p[5], F2 = (Real)p[6], F3 = (Real)p[7];
    const Real eta = (Real)p[8], Fe = (Real)p[9], rho_e = (Real)p[10],
               rho_s = (Real)p[11];
    const Real rho_n = Real(0.85) * rho_e;   // computed in working precision, as
    const Real rho_0 = Real(1.15) * rho_e;   // the reference does (zjw04.py:319-322)
    auto e1 = [&](Real &y, Real &dy) {
        const Real x = rho / rho_n - Real(1);
        y = Fn0 + (Fn1 * x + Fn2 * x * x + Fn3 * x * x * x);
        dy = (Fn1 + Real(2) * Fn2 * x + Real(3) * Fn3 * x * x) / rho_n;
    };
    auto e2 = [&](Real &y, Real &dy) {
        const Real x = rho / rho_e - Real(1);
        y = F0 + (F1 * x + F2 * x * x + F3 * x * x * x);
        dy = (F1 + Real(2) * F2 * x + Real(3) * F3 * x * x) / rho_e;
    };
    auto e3 = [&](Real shift, Real &y, Real &dy) {
        const Real x = rho / rho_s + shift;
        const Real lnx = Math<Real>::log_(x);
        const Real xe = Math<Real>::pow_(x, eta);
        y = Fe * (Real(1) - eta * lnx) * xe;
        // d/dx [Fe (1 - eta ln x) x^eta] = -Fe eta^2 ln(x) x^(eta-1)
        dy = -Fe * eta * eta * lnx * xe / x / rho_s;
    };
    if (!blended) {
        if (rho < rho_n) e1(F, dF);
        else if (rho < rho_0) e2(F, dF);
        else e3(Real(0), F, dF);
        return;
    }
    Real y1, d1, y2, d2, y3, d3;
    e1(y1, d1);
    e2(y2, d2);
    e3(Real(1e-8), y3, d3);
    const Real c1 = Real(1) / (Real(1) + Math<Real>::exp_(-Real(2) * (rho_n - rho)));
    const Real c3 = Real(1) / (Real(1) + Math<Real>::exp_(-Real(2) * (rho - rho_0)));
    const Real c2 = Real(1) - (c1 + c3);
    const Real dc1 = -Real(2) * c1 * (Real(1) - c1);
    const Real dc3 = Real(2) * c3 * (Real(1) - c3);
    const Real dc2 = -(dc1 + dc3);
    F = c1 * y1 + c2 * y2 + c3 * y3;
    dF = dc1 * y1 + c1 * d1 + dc2 * y2 + c2 * d2 + dc3 * y3 + c3 * d3;
}

// ---------------------------------------------------------------------------
// 'nn' functions (eam.py:174-190 -> convolution1x1 on a scalar): y = MLP(x) with
// hidden layers (bias + activation) and a linear output unit without bias.  Value
// and dy/dx by forward-mode differentiation.  Weights live in the model's pool at
// pool + 4*aux:  layer 0: w[h0], b[h0];  layer l: W[h_{l-1} x h_l] (in-major), b[h_l];
// output: w[h_last].   p = n_hidden, activation id, h0, h1, ...
// ---------------------------------------------------------------------------
#define TAB_MLP_FN_MAXW 64
#define TAB_MLP_FN_MAXL 4

template <typename Real>
__device__ __forceinline__ Real mlp_fn_act(int kind, Real z, Real &d) {
    switch (kind) {
    case 0: {   // softplus
        const Real e = Math<Real>::exp_(-fabs(z));
        d = z >= Real(0) ? Real(1) / (Real(1) + e) : e / (Real(1) + e);
        return (z > Real(0) ? z : Real(0)) + Math<Real>::log_(Real(1) + e);
    }
    case 1: {   // tanh
        const Real t = tanh(z);
        d = Real(1) - t * t;
        return t;
    }
    case 2:
        d = z > Real(0) ? Real(1) : Real(0);
        return z > Real(0) ? z : Real(0);
    case 3:
        d = z > Real(0) ? Real(1) : Real(0.2);
        return z > Real(0) ? z : Real(0.2) * z;
    case 4: {
        const Real sg = Real(1) / (Real(1) + Math<Real>::exp_(-z));
        d = sg * (Real(1) - sg);
        return sg;
    }
    case 5: {
        const Real q = Real(1) + fabs(z);
        d = Real(1) / (q * q);
        return z / q;
    }
    case 6: {
        const Real e = Math<Real>::exp_(z);
        d = z > Real(0) ? Real(1) : e;
        return z > Real(0) ? z : e - Real(1);
    }
    default: {
        const Real sq = Math<Real>::sqrt_(z * z + Real(4));
        d = Real(0.5) * (Real(1) + z / sq);
        return Real(0.5) * (z + sq);
    }
    }
}

template <typename Real>
__device__ __noinline__ void mlp_fn_eval(const tab_fn &fn, const double *__restrict__ pool,
                                         Real x, Real &f, Real &df) {
    const int nh = (int)fn.p[0], act = (int)fn.p[1];
    const double *w = pool + (size_t)fn.aux * 4;
    Real h[TAB_MLP_FN_MAXW], dh[TAB_MLP_FN_MAXW], g[TAB_MLP_FN_MAXW], dg[TAB_MLP_FN_MAXW];
    int n0 = (int)fn.p[2];
    for (int o = 0; o < n0; ++o) {
        const Real wo = (Real)w[o];
        Real d;
        h[o] = mlp_fn_act<Real>(act, wo * x + (Real)w[n0 + o], d);
        dh[o] = d * wo;
    }
    w += 2 * n0;
    for (int l = 1; l < nh; ++l) {
        const int n1 = (int)fn.p[2 + l];
        const double *b = w + (size_t)n0 * n1;
        for (int o = 0; o < n1; ++o) {
            Real z = (Real)b[o], dz = Real(0);
            for (int k = 0; k < n0; ++k) {
                const Real wk = (Real)w[(size_t)k * n1 + o];
                z += h[k] * wk;
                dz += dh[k] * wk;
            }
            Real d;
            g[o] = mlp_fn_act<Real>(act, z, d);
            dg[o] = d * dz;
        }
        for (int o = 0; o < n1; ++o) {
            h[o] = g[o];
            dh[o] = dg[o];
        }
        w = b + n1;
        n0 = n1;
    }
    Real y = Real(0), dy = Real(0);
    for (int k = 0; k < n0; ++k) {
        y += h[k] * (Real)w[k];
        dy += dh[k] * (Real)w[k];
    }
    f = y;
    df = dy;
}

// NOTE: tab_eam_create stores 1/r_eq in the r_eq slots of the device tables.
// One table entry -> value and d/dr.  The switch is warp-uniform for
// single-species systems and cheap next to the transcendental work otherwise.
// cubic spline table: value and derivative; out-of-range arguments use the edge
// interval's polynomial (LAMMPS clamps the same way)
template <typename Real>
__device__ __forceinline__ void spline_eval(const tab_fn &fn, const double *__restrict__ pool,
                                            Real x, Real &f, Real &df) {
    const double t = ((double)x - fn.p[0]) * fn.p[1];
    int k = (int)floor(t);
    const int last = (int)fn.p[2] - 1;
    k = k < 0 ? 0 : (k > last ? last : k);
    const double *c = pool + ((size_t)fn.aux + k) * 4;
    const Real d = (Real)((double)x - (fn.p[0] + (double)k / fn.p[1]));
    const Real c0 = (Real)c[0], c1 = (Real)c[1], c2 = (Real)c[2], c3 = (Real)c[3];
    f = ((c3 * d + c2) * d + c1) * d + c0;
    df = (Real(3) * c3 * d + Real(2) * c2) * d + c1;
}

// x^n and n x^(n-1) for small integer n
template <typename Real>
__device__ __forceinline__ void ipow_d(Real x, int n, Real &v, Real &dv) {
    Real p1 = Real(1);
    for (int k = 1; k < n; ++k) p1 *= x;
    dv = (Real)n * p1;
    v = p1 * x;
}

// msah11.py:52-157 -- the pair function of the Mendelev Al-Fe potential: a screened-Coulomb
// head (c0 / r) sum b_i exp(c_i r) on [lo, hi), an exp(cubic) bridge, and tails
// sum_k a_k (hi - r)^n_k on [lo, hi).  Constants in the coefficient pool (layout: tab200.h).
template <typename Real>
__device__ __forceinline__ void msah_phi(const double *__restrict__ c, Real r, Real &f,
                                         Real &df) {
    f = Real(0);
    df = Real(0);
    const int n_poly = (int)c[0];
    const double *q = c + 1;
    if ((double)r >= q[0] && (double)r < q[1]) {
        Real s = Real(0), ds = Real(0);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const Real b = (Real)q[3 + 2 * i], ci = (Real)q[4 + 2 * i];
            const Real e = b * Math<Real>::exp_(ci * r);
            s += e;
            ds += ci * e;
        }
        const Real sc = (Real)q[2] / r;
        f += sc * s;
        df += sc * ds - sc * s / r;
    }
    q += 11;
    if ((double)r >= q[0] && (double)r < q[1]) {
        const Real c1 = (Real)q[3], c2 = (Real)q[4], c3 = (Real)q[5];
        const Real e = Math<Real>::exp_((Real)q[2] + r * (c1 + r * (c2 + r * c3)));
        f += e;
        df += e * (c1 + r * (Real(2) * c2 + Real(3) * c3 * r));
    }
    q += 6;
    for (int g = 0; g < n_poly; ++g) {
        const int nt = (int)q[2];
        if ((double)r >= q[0] && (double)r < q[1]) {
            const Real x = (Real)q[1] - r;
            for (int k = 0; k < nt; ++k) {
                Real v, dv;
                ipow_d<Real>(x, (int)q[4 + 2 * k], v, dv);
                f += (Real)q[3 + 2 * k] * v;
                df -= (Real)q[3 + 2 * k] * dv;
            }
        }
        q += 3 + 2 * nt;
    }
}

// NN = the model holds 'nn' (MLP) functions: only then is mlp_fn_eval (local arrays, a
// stack frame and ~2x the registers) compiled into the calling kernel.
template <typename Real, bool NN = false>
__device__ __forceinline__ void eval_pair_fn(const tab_fn &fn, Real r, Real &f,
                                             Real &df,
                                             const double *__restrict__ pool = nullptr) {
    const double *p = fn.p;
    if (NN && fn.kind == TAB_FN_MLP) {
        mlp_fn_eval<Real>(fn, pool, r, f, df);
        return;
    }
    switch (fn.kind) {
    case TAB_FN_SPLINE:
        spline_eval<Real>(fn, pool, r, f, df);
        break;
    case TAB_FN_MSAH_PHI:   // msah11.py:52-301
        msah_phi<Real>(pool + (size_t)fn.aux * 4, r, f, df);
        break;
    case TAB_FN_POWCUT_RHO: {   // msah11.py:303-352
        const int order = (int)p[0], n = (int)p[1];
        f = Real(0);
        df = Real(0);
        for (int i = 0; i < n; ++i) {
            const Real x = (Real)p[3 + 2 * i] - r;
            if (x > Real(0)) {
                Real v, dv;
                ipow_d<Real>(x, order, v, dv);
                f += (Real)p[2 + 2 * i] * v;
                df -= (Real)p[2 + 2 * i] * dv;
            }
        }
        break;
    }
    case TAB_FN_ZHOU_RHO:   // zjw04.py:245-277
        zhou_exp<Real>(r, (Real)p[0], (Real)p[1], (Real)p[2], (Real)p[3], f, df);
        break;
    case TAB_FN_ZHOU_PHI: { // zjw04.py:187-227
        Real fa, dfa, fb, dfb;
        zhou_exp<Real>(r, (Real)p[0], (Real)p[1], (Real)p[2], (Real)p[6], fa, dfa);
        zhou_exp<Real>(r, (Real)p[3], (Real)p[4], (Real)p[5], (Real)p[6], fb, dfb);
        f = fa - fb;
        df = dfa - dfb;
        break;
    }
    case TAB_FN_ZHOU_PHI_MIX: { // zjw04.py:229-243
        // p[0..6] phi of a, p[7..10] rho of a, p[11..17] phi of b, p[18..21] rho of b
        Real t0, d0, t1, d1, pa, dpa, pb, dpb, ra, dra, rb, drb;
        zhou_exp<Real>(r, (Real)p[0], (Real)p[1], (Real)p[2], (Real)p[6], t0, d0);
        zhou_exp<Real>(r, (Real)p[3], (Real)p[4], (Real)p[5], (Real)p[6], t1, d1);
        pa = t0 - t1;
        dpa = d0 - d1;
        zhou_exp<Real>(r, (Real)p[7], (Real)p[8], (Real)p[9], (Real)p[10], ra, dra);
        zhou_exp<Real>(r, (Real)p[11], (Real)p[12], (Real)p[13], (Real)p[17], t0, d0);
        zhou_exp<Real>(r, (Real)p[14], (Real)p[15], (Real)p[16], (Real)p[17], t1, d1);
        pb = t0 - t1;
        dpb = d0 - d1;
        zhou_exp<Real>(r, (Real)p[18], (Real)p[19], (Real)p[20], (Real)p[21], rb, drb);
        const Real qab = ra / rb, qba = rb / ra;
        const Real dqab = (dra - qab * drb) / rb;
        const Real dqba = (drb - qba * dra) / ra;
        f = Real(0.5) * (qab * pb + qba * pa);
        df = Real(0.5) * (dqab * pb + qab * dpb + dqba * pa + qba * dpa);
        break;
    }
    case TAB_FN_SUTTON_RHO: {   // sutton90.py:61-78  (a/r)^6
        const Real q = (Real)p[0] / r, q2 = q * q;
        f = q2 * q2 * q2;
        df = Real(-6) * f / r;
        break;
    }
    case TAB_FN_SUTTON_PHI: {   // sutton90.py:43-59  (b/r)^12
        const Real q = (Real)p[0] / r, q2 = q * q, q4 = q2 * q2;
        f = q4 * q4 * q4;
        df = Real(-12) * f / r;
        break;
    }
    case TAB_FN_AGRAWAL_RHO: {  // agrawal.py:57-83
        const Real A = (Real)p[0], Bq = (Real)p[1], re = (Real)p[2], rc = (Real)p[3],
                   m = (Real)p[4];
        const Real rho0 = A * Math<Real>::exp_(-Bq * (r - re));
        const Real tail = A * Math<Real>::exp_(-Bq * (rc - re));
        const Real drho = -Bq * tail;
        const Real zm1 = Math<Real>::pow_(r / rc, m - Real(1));
        f = rho0 - tail + rc / m * (Real(1) - zm1 * (r / rc)) * drho;
        df = -Bq * rho0 - zm1 * drho;
        break;
    }
    case TAB_FN_AGRAWAL_PHI: {  // agrawal.py:124-152 (generic.py:15-30 morse)
        const Real D = (Real)p[0], al = (Real)p[1], re = (Real)p[2], rc = (Real)p[3],
                   m = (Real)p[4];
        const Real e1 = Math<Real>::exp_(-al * (r - re)), e1c = Math<Real>::exp_(-al * (rc - re));
        const Real phi0 = D * (e1 * e1 - Real(2) * e1);
        const Real phic = D * (e1c * e1c - Real(2) * e1c);
        const Real dphic = Real(2) * D * al * (e1c - e1c * e1c);
        const Real zm1 = Math<Real>::pow_(r / rc, m - Real(1));
        f = phi0 - phic + rc / m * (Real(1) - zm1 * (r / rc)) * dphic;
        df = Real(2) * D * al * (e1 - e1 * e1) - zm1 * dphic;
        break;
    }
    case TAB_FN_GRIMES_RHO: {   // grimmes.py:62-84  n/r^8 * (1+erf(20(r-1.5)))/2
        const Real r2 = r * r, r4 = r2 * r2, r8 = r4 * r4;
        const Real left = (Real)p[0] / r8;
        const Real t = Real(20) * (r - Real(1.5));
        const Real right = Real(0.5) + Real(0.5) * Math<Real>::erf_(t);
        f = left * right;
        df = Real(-8) * f / r +
             left * Real(20) * Real(0.56418958354775628695) * Math<Real>::exp_(-t * t);
        break;
    }
    case TAB_FN_GRIMES_PHI: {   // grimmes.py:41-60: morse + buckingham
        const Real A = (Real)p[0], rho = (Real)p[1], C = (Real)p[2], D = (Real)p[3],
                   gm = (Real)p[4], r0 = (Real)p[5];
        const Real e1 = Math<Real>::exp_(-gm * (r - r0));
        const Real eb = A * Math<Real>::exp_(-r / rho);
        const Real r2 = r * r, r6 = r2 * r2 * r2;
        f = D * (e1 * e1 - Real(2) * e1) + eb - C / r6;
        df = Real(2) * D * gm * (e1 - e1 * e1) - eb / rho + Real(6) * C / (r6 * r);
        break;
    }
    case TAB_FN_MISHIN_POLAR: { // generic.py:52-84: (p1 e^{-p2 r} + p3) psi((r-rc)/h)
        const Real p1 = (Real)p[0], p2 = (Real)p[1], p3 = (Real)p[2], rc = (Real)p[3],
                   h = (Real)p[4];
        const Real z = (r - rc) / h;
        if (z >= Real(0)) {
            f = Real(0);
            df = Real(0);
        } else {
            const Real z2 = z * z, z4 = z2 * z2;
            const Real den = Real(1) / (Real(1) + z4);
            const Real psi = z4 * den;
            const Real dpsi = Real(4) * z2 * z * den * den / h;
            const Real ex = p1 * Math<Real>::exp_(-p2 * r);
            f = (ex + p3) * psi;
            df = -p2 * ex * psi + (ex + p3) * dpsi;
        }
        break;
    }
    default:
        f = Real(0);
        df = Real(0);
    }
}

template <typename Real, bool NN = false>
__device__ __forceinline__ void eval_embed_fn(const tab_fn &fn, Real rho, Real &F,
                                              Real &dF,
                                              const double *__restrict__ pool = nullptr) {
    if (NN && fn.kind == TAB_FN_MLP) {
        mlp_fn_eval<Real>(fn, pool, rho, F, dF);
        return;
    }
    switch (fn.kind) {
    case TAB_FN_SPLINE:
        spline_eval<Real>(fn, pool, rho, F, dF);
        break;
    case TAB_FN_MSAH_EMBED_AL: {   // msah11.py:400-411
        if (rho >= Real(1e-12)) {
            const Real c1 = (Real)fn.p[0], c2 = (Real)fn.p[1];
            const Real sq = Math<Real>::sqrt_(rho), lg = Math<Real>::log_(rho);
            F = -sq + c1 * rho * rho - c2 * rho * lg;
            dF = Real(-0.5) / sq + Real(2) * c1 * rho - c2 * (lg + Real(1));
        } else {
            F = Real(0);
            dF = Real(0);
        }
        break;
    }
    case TAB_FN_MSAH_EMBED_FE: {   // msah11.py:412-420
        const Real c3 = (Real)fn.p[0], c4 = (Real)fn.p[1];
        const Real r2 = rho * rho;
        if (rho > Real(0)) {
            const Real sq = Math<Real>::sqrt_(rho);
            F = -sq - c3 * r2 + c4 * r2 * r2;
            dF = Real(-0.5) / sq - Real(2) * c3 * rho + Real(4) * c4 * r2 * rho;
        } else {
            F = Real(0);
            dF = Real(0);
        }
        break;
    }
    case TAB_FN_ZHOU_EMBED:
        zhou_embed<Real>(fn.p, false, rho, F, dF);
        break;
    case TAB_FN_ZHOU_EMBED_XC:
        zhou_embed<Real>(fn.p, true, rho, F, dF);
        break;
    case TAB_FN_SQRT_EMBED: {      // sutton90.py:80-97 (G = 1), grimmes.py:86-101
        const Real G = (Real)fn.p[0];
        if (rho > Real(0)) {
            const Real sq = Math<Real>::sqrt_(rho);
            F = -G * sq;
            dF = -G * Real(0.5) / sq;
        } else {
            F = Real(0);
            dF = Real(0);
        }
        break;
    }
    case TAB_FN_AGRAWAL_EMBED: {   // agrawal.py:85-122
        const Real F0 = (Real)fn.p[0], F1 = (Real)fn.p[1], be = (Real)fn.p[2],
                   ga = (Real)fn.p[3];
        if (rho > Real(0)) {
            const Real lg = Math<Real>::log_(rho > Real(1e-12) ? rho : Real(1e-12));
            const Real xb = Math<Real>::pow_(rho, be), xg = Math<Real>::pow_(rho, ga);
            F = F0 * (Real(1) - be * lg) * xb + F1 * xg;
            const Real dlg = rho > Real(1e-12) ? Real(1) / rho : Real(0);
            dF = F0 * (-be * dlg * xb + (Real(1) - be * lg) * be * xb / rho) +
                 F1 * ga * xg / rho;
        } else {
            F = Real(0);
            dF = Real(0);
        }
        break;
    }
    case TAB_FN_MISHIN_EMBED: {    // mishin.py:196-260
        const double *s = fn.p;
        const Real s1 = (Real)s[0], s2 = (Real)s[1], s3 = (Real)s[2], s4 = (Real)s[3],
                   s5 = (Real)s[4], s6 = (Real)s[5], s7 = (Real)s[6], eps = (Real)s[7];
        const Real r2 = rho * rho, r3 = r2 * rho, r4 = r2 * r2;
        const Real re = rho + eps;
        const Real pw = Math<Real>::pow_(re, s5);
        const Real S = s1 * rho + s2 * r2 + s3 * r3 - s4 * pw;
        const Real dS = s1 + Real(2) * s2 * rho + Real(3) * s3 * r2 - s4 * s5 * pw / re;
        const Real a = Real(1) - s6 * r2, b = Real(1) + s7 * r4;
        const Real om = Real(1) - a / b;
        const Real dom = -((-Real(2) * s6 * rho) * b - a * (Real(4) * s7 * r3)) / (b * b);
        F = S * om;
        dF = dS * om + S * dom;
        break;
    }
    default:
        F = Real(0);
        dF = Real(0);
    }
}
