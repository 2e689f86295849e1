// hessian.cu -- analytic Hessian d2E/dR_a dR_b of the EAM / FS models.
//
// Replaces BasicNN._get_hessian_op (nn/basic.py:410-421: tf.hessians(E, R), i.e.
// 3*Nvap sequential reverse-mode passes) by closed-form second derivatives:
//   E = sum_i F_i(rho_i) + 1/2 sum_p phi_p(r_p),  rho_i = sum_{p from i} rho_p(r_p)
//   pair kernel   K_p = f'' n(x)n + f'/r (I - n(x)n)         for f = rho_p, phi_p
//   H_ab = sum_p (d_aj - d_ai)(d_bj - d_bi) [F'_i K^rho_p + 1/2 K^phi_p]
//        + sum_c F''_c g_c^a (x) g_c^b ,   g_c^x = sum_{p from c} rho'_p n_p (d_{x,j} - d_{x,c})
// with j = OWNER of the neighbour (periodic images of one atom add up; an atom's
// own image contributes nothing, as moving the atom moves its image).
//
// First and second derivatives of every potential kind come from ONE value
// formula evaluated on a forward-mode dual number (value, d/dx, d2/dx2).
//
// Output: dense H [N,3,N,3] float64 in caller atom order.  Row block a is written
// by thread a only (atomic-free, deterministic).
#include "potentials.cuh"

// ---------------------------------------------------------------------------
// second-order dual numbers
// ---------------------------------------------------------------------------
struct D2 {
    double v, d, dd;
    __device__ D2() {}
    __device__ D2(double a) : v(a), d(0), dd(0) {}
    __device__ D2(double a, double b, double c) : v(a), d(b), dd(c) {}
};
__device__ inline D2 operator+(D2 a, D2 b) { return D2(a.v + b.v, a.d + b.d, a.dd + b.dd); }
__device__ inline D2 operator-(D2 a, D2 b) { return D2(a.v - b.v, a.d - b.d, a.dd - b.dd); }
__device__ inline D2 operator-(D2 a) { return D2(-a.v, -a.d, -a.dd); }
__device__ inline D2 operator*(D2 a, D2 b) {
    return D2(a.v * b.v, a.d * b.v + a.v * b.d, a.dd * b.v + 2.0 * a.d * b.d + a.v * b.dd);
}
// g(a) with g', g'' known at a.v
__device__ inline D2 chain(D2 a, double g, double g1, double g2) {
    return D2(g, g1 * a.d, g2 * a.d * a.d + g1 * a.dd);
}
__device__ inline D2 rcp(D2 a) {
    const double i = 1.0 / a.v;
    return chain(a, i, -i * i, 2.0 * i * i * i);
}
__device__ inline D2 operator/(D2 a, D2 b) { return a * rcp(b); }
__device__ inline D2 dexp(D2 a) {
    const double e = exp(a.v);
    return chain(a, e, e, e);
}
__device__ inline D2 dlog(D2 a) {
    const double i = 1.0 / a.v;
    return chain(a, log(a.v), i, -i * i);
}
__device__ inline D2 dsqrt(D2 a) {
    const double s = sqrt(a.v);
    return chain(a, s, 0.5 / s, -0.25 / (s * a.v));
}
__device__ inline D2 dpow(D2 a, double y) {
    const int yi = (int)y;
    if ((double)yi == y && yi >= 0 && yi <= 32) {     // integer powers: any sign of base
        double p2 = 1.0;                              // a^(y-2)
        for (int k = 2; k < yi; ++k) p2 *= a.v;
        const double p1 = yi >= 2 ? p2 * a.v : (yi == 1 ? 1.0 : 0.0);
        const double p0 = yi >= 1 ? p1 * a.v : 1.0;
        return chain(a, p0, y * p1, yi >= 2 ? y * (y - 1.0) * p2 : 0.0);
    }
    const double p = pow(a.v, y);
    return chain(a, p, y * p / a.v, y * (y - 1.0) * p / (a.v * a.v));
}
__device__ inline D2 derf(D2 a) {
    const double g1 = 1.1283791670955125739 * exp(-a.v * a.v);
    return chain(a, erf(a.v), g1, -2.0 * a.v * g1);
}
__device__ inline D2 dsigmoid(D2 a) {
    const double s = 1.0 / (1.0 + exp(-a.v));
    return chain(a, s, s * (1.0 - s), s * (1.0 - s) * (1.0 - 2.0 * s));
}

// ---------------------------------------------------------------------------
// value formulas (same references as potentials.cuh)
// ---------------------------------------------------------------------------
__device__ inline D2 v_zhou(D2 r, double a, double b, double c, double inv_re) {
    const D2 x = r * D2(inv_re);
    return D2(a) * dexp(D2(b) - D2(b) * x) / (D2(1.0) + dpow(x - D2(c), 20.0));
}
__device__ inline D2 v_zhou_phi(D2 r, const double *p) {   // A,alpha,kappa,B,beta,lamda,1/re
    return v_zhou(r, p[0], p[1], p[2], p[6]) - v_zhou(r, p[3], p[4], p[5], p[6]);
}
__device__ inline D2 v_morse(D2 r, double d, double g, double r0) {
    const D2 e = dexp(D2(-g) * (r - D2(r0)));
    return D2(d) * (e * e - D2(2.0) * e);
}

__device__ inline D2 spline_d2(const tab_fn &fn, const double *pool, D2 x) {
    const double t = (x.v - fn.p[0]) * fn.p[1];
    int k = (int)floor(t);
    const int last = (int)fn.p[2] - 1;
    k = k < 0 ? 0 : (k > last ? last : k);
    const double *c = pool + ((size_t)fn.aux + k) * 4;
    const double d = x.v - (fn.p[0] + (double)k / fn.p[1]);
    return chain(x, ((c[3] * d + c[2]) * d + c[1]) * d + c[0],
                 (3.0 * c[3] * d + 2.0 * c[2]) * d + c[1], 6.0 * c[3] * d + 2.0 * c[2]);
}

__device__ const double *g_hess_pool = nullptr;   // set per launch (single stream use)

// msah11.py:52-301 (layout of the constants: tab200.h, TAB_FN_MSAH_PHI)
__device__ inline D2 msah_phi_d2(const double *c, D2 r) {
    D2 out(0.0);
    const int n_poly = (int)c[0];
    const double *q = c + 1;
    if (r.v >= q[0] && r.v < q[1]) {
        D2 s(0.0);
        for (int i = 0; i < 4; ++i) s = s + D2(q[3 + 2 * i]) * dexp(D2(q[4 + 2 * i]) * r);
        out = out + D2(q[2]) * s / r;
    }
    q += 11;
    if (r.v >= q[0] && r.v < q[1])
        out = out + dexp(D2(q[2]) + r * (D2(q[3]) + r * (D2(q[4]) + r * D2(q[5]))));
    q += 6;
    for (int g = 0; g < n_poly; ++g) {
        const int nt = (int)q[2];
        if (r.v >= q[0] && r.v < q[1]) {
            const D2 x = D2(q[1]) - r;
            for (int k = 0; k < nt; ++k) out = out + D2(q[3 + 2 * k]) * dpow(x, q[4 + 2 * k]);
        }
        q += 3 + 2 * nt;
    }
    return out;
}

__device__ inline D2 eval_pair_d2(const tab_fn &fn, D2 r) {
    const double *p = fn.p;
    switch (fn.kind) {
    case TAB_FN_SPLINE:
        return spline_d2(fn, g_hess_pool, r);
    case TAB_FN_MSAH_PHI:
        return msah_phi_d2(g_hess_pool + (size_t)fn.aux * 4, r);
    case TAB_FN_POWCUT_RHO: {   // msah11.py:303-352
        D2 out(0.0);
        const int n = (int)p[1];
        for (int i = 0; i < n; ++i)
            if (p[3 + 2 * i] - r.v > 0.0)
                out = out + D2(p[2 + 2 * i]) * dpow(D2(p[3 + 2 * i]) - r, p[0]);
        return out;
    }
    case TAB_FN_ZHOU_RHO:
        return v_zhou(r, p[0], p[1], p[2], p[3]);
    case TAB_FN_ZHOU_PHI:
        return v_zhou_phi(r, p);
    case TAB_FN_ZHOU_PHI_MIX: {
        const D2 pa = v_zhou_phi(r, p), ra = v_zhou(r, p[7], p[8], p[9], p[10]);
        const D2 pb = v_zhou_phi(r, p + 11), rb = v_zhou(r, p[18], p[19], p[20], p[21]);
        return D2(0.5) * (ra / rb * pb + rb / ra * pa);
    }
    case TAB_FN_SUTTON_RHO:
        return dpow(D2(p[0]) / r, 6.0);
    case TAB_FN_SUTTON_PHI:
        return dpow(D2(p[0]) / r, 12.0);
    case TAB_FN_AGRAWAL_RHO: {
        const double A = p[0], B = p[1], re = p[2], rc = p[3], m = p[4];
        const double tail = A * exp(-B * (rc - re)), drho = -B * tail;
        return D2(A) * dexp(D2(-B) * (r - D2(re))) - D2(tail) +
               D2(rc / m * drho) * (D2(1.0) - dpow(r / D2(rc), m));
    }
    case TAB_FN_AGRAWAL_PHI: {
        const double D = p[0], al = p[1], re = p[2], rc = p[3], m = p[4];
        const double e1c = exp(-al * (rc - re));
        const double phic = D * (e1c * e1c - 2.0 * e1c);
        const double dphic = 2.0 * D * al * (e1c - e1c * e1c);
        return v_morse(r, D, al, re) - D2(phic) +
               D2(rc / m * dphic) * (D2(1.0) - dpow(r / D2(rc), m));
    }
    case TAB_FN_GRIMES_RHO:
        return D2(p[0]) / dpow(r, 8.0) *
               (D2(0.5) + D2(0.5) * derf(D2(20.0) * (r - D2(1.5))));
    case TAB_FN_GRIMES_PHI:
        return v_morse(r, p[3], p[4], p[5]) + D2(p[0]) * dexp(-r / D2(p[1])) -
               D2(p[2]) / dpow(r, 6.0);
    case TAB_FN_MISHIN_POLAR: {
        const D2 z = (r - D2(p[3])) / D2(p[4]);
        if (z.v >= 0.0) return D2(0.0);
        const D2 z4 = dpow(z, 4.0);
        return (D2(p[0]) * dexp(D2(-p[1]) * r) + D2(p[2])) * (z4 / (D2(1.0) + z4));
    }
    default:
        return D2(0.0);
    }
}

__device__ inline D2 v_zhou_embed(const double *p, bool blended, D2 rho) {
    const double rho_e = p[10], rho_s = p[11];
    const double rho_n = 0.85 * rho_e, rho_0 = 1.15 * rho_e;
    auto e1 = [&]() {
        const D2 x = rho / D2(rho_n) - D2(1.0);
        return D2(p[0]) + (D2(p[1]) * x + D2(p[2]) * x * x + D2(p[3]) * x * x * x);
    };
    auto e2 = [&]() {
        const D2 x = rho / D2(rho_e) - D2(1.0);
        return D2(p[4]) + (D2(p[5]) * x + D2(p[6]) * x * x + D2(p[7]) * x * x * x);
    };
    auto e3 = [&](double shift) {
        const D2 x = rho / D2(rho_s) + D2(shift);
        return D2(p[9]) * (D2(1.0) - D2(p[8]) * dlog(x)) * dpow(x, p[8]);
    };
    if (!blended) {
        if (rho.v < rho_n) return e1();
        if (rho.v < rho_0) return e2();
        return e3(0.0);
    }
    const D2 c1 = dsigmoid(D2(2.0) * (D2(rho_n) - rho));
    const D2 c3 = dsigmoid(D2(2.0) * (rho - D2(rho_0)));
    const D2 c2 = D2(1.0) - (c1 + c3);
    return c1 * e1() + c2 * e2() + c3 * e3(1e-8);
}

__device__ inline D2 eval_embed_d2(const tab_fn &fn, D2 rho) {
    const double *p = fn.p;
    switch (fn.kind) {
    case TAB_FN_SPLINE:
        return spline_d2(fn, g_hess_pool, rho);
    case TAB_FN_MSAH_EMBED_AL:   // msah11.py:400-411
        if (rho.v < 1e-12) return D2(0.0);
        return -dsqrt(rho) + D2(p[0]) * rho * rho - D2(p[1]) * rho * dlog(rho);
    case TAB_FN_MSAH_EMBED_FE: { // msah11.py:412-420
        if (!(rho.v > 0.0)) return D2(0.0);
        const D2 r2 = rho * rho;
        return -dsqrt(rho) - D2(p[0]) * r2 + D2(p[1]) * r2 * r2;
    }
    case TAB_FN_ZHOU_EMBED:
        return v_zhou_embed(p, false, rho);
    case TAB_FN_ZHOU_EMBED_XC:
        return v_zhou_embed(p, true, rho);
    case TAB_FN_SQRT_EMBED:
        return rho.v > 0.0 ? -(D2(p[0]) * dsqrt(rho)) : D2(0.0);
    case TAB_FN_AGRAWAL_EMBED: {
        if (!(rho.v > 0.0)) return D2(0.0);
        const D2 lg = rho.v > 1e-12 ? dlog(rho) : D2(log(1e-12));
        return D2(p[0]) * (D2(1.0) - D2(p[2]) * lg) * dpow(rho, p[2]) +
               D2(p[1]) * dpow(rho, p[3]);
    }
    case TAB_FN_MISHIN_EMBED: {
        const D2 r2 = rho * rho, r3 = r2 * rho, r4 = r2 * r2;
        const D2 S = D2(p[0]) * rho + D2(p[1]) * r2 + D2(p[2]) * r3 -
                     D2(p[3]) * dpow(rho + D2(p[7]), p[4]);
        const D2 om = D2(1.0) - (D2(1.0) - D2(p[5]) * r2) / (D2(1.0) + D2(p[6]) * r4);
        return S * om;
    }
    default:
        return D2(0.0);
    }
}

// ---------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------
struct HessCtx {
    int n, n_loc, n_el;
    const Atom4 *atoms;
    const uint8_t *types_ext;
    const int *counts;
    const uint32_t *slice_ptr;
    const uint32_t *col;
    const int *ghost_owner;
    const int *perm;
    const tab_fn *rho, *phi, *embed;
};

__device__ inline int owner_of(const HessCtx &c, int j) {
    return j < c.n_loc ? j : c.ghost_owner[j - c.n_loc];
}

// per atom: rho, F', F''  (sorted order)
__global__ void k_hess_rho(HessCtx c, double *__restrict__ fp, double *__restrict__ fpp) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c.n) return;
    const Atom4 me = c.atoms[i];
    const int ti = c.types_ext[i];
    const size_t base = (size_t)c.slice_ptr[i >> 5] * 32u + (i & 31);
    double rho = 0.0;
    for (int k = 0; k < c.counts[i]; ++k) {
        const uint32_t e = c.col[base + (size_t)k * 32u];
        const Atom4 a = c.atoms[e & TAB_COL_IDX_MASK];
        const double dx = a.x - me.x, dy = a.y - me.y, dz = a.z - me.z;
        const double r = sqrt(dx * dx + dy * dy + dz * dz + 1e-14);
        rho += eval_pair_d2(c.rho[ti * c.n_el + (int)(e >> TAB_COL_TYPE_SHIFT)],
                            D2(r, 1.0, 0.0)).v;
    }
    const D2 F = eval_embed_d2(c.embed[ti], D2(rho, 1.0, 0.0));
    fp[i] = F.d;
    fpp[i] = F.dd;
}

// block (3x3) accumulate into row a of H (caller order): H[oa, :, ob, :]
__device__ inline void add_block(double *H, int n, int oa, int ob, const double *K,
                                 double s) {
#pragma unroll
    for (int al = 0; al < 3; ++al)
#pragma unroll
        for (int be = 0; be < 3; ++be)
            H[(((size_t)oa * 3 + al) * n + ob) * 3 + be] += s * K[al * 3 + be];
}

__device__ inline void pair_geom(const Atom4 &me, const Atom4 &a, double *nv, double &r) {
    const double dx = a.x - me.x, dy = a.y - me.y, dz = a.z - me.z;
    // r = sqrt(D.D + eps): derivatives of the reference include the eps
    r = sqrt(dx * dx + dy * dy + dz * dz + 1e-14);
    nv[0] = dx / r;
    nv[1] = dy / r;
    nv[2] = dz / r;
}

// K = f2 n(x)n + f1/r (I - n(x)n).  With r = sqrt(D.D+eps), d r/dD = D/r = n and
// d n/dD = (I - n(x)n)/r exactly, so the formula holds with |n| slightly < 1.
__device__ inline void pair_kernel(const double *nv, double r, double f1, double f2,
                                   double *K) {
    const double a = f1 / r;
#pragma unroll
    for (int al = 0; al < 3; ++al)
#pragma unroll
        for (int be = 0; be < 3; ++be)
            K[al * 3 + be] = (f2 - a) * nv[al] * nv[be] + (al == be ? a : 0.0);
}

__global__ void k_hessian(HessCtx c, const double *__restrict__ fp,
                          const double *__restrict__ fpp, double *__restrict__ H) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= c.n) return;
    const int n = c.n;
    const int oa = c.perm[a];
    const Atom4 me = c.atoms[a];
    const int ta = c.types_ext[a];
    const size_t base = (size_t)c.slice_ptr[a >> 5] * 32u + (a & 31);
    const int cnt = c.counts[a];
    double s_a[3] = {0, 0, 0};          // s_a = sum_q v_q over row a
    // ---- pair terms + gather of s_a
    for (int k = 0; k < cnt; ++k) {
        const uint32_t e = c.col[base + (size_t)k * 32u];
        const int j = (int)(e & TAB_COL_IDX_MASK), tj = (int)(e >> TAB_COL_TYPE_SHIFT);
        const int o = owner_of(c, j);
        double nv[3], r;
        pair_geom(me, c.atoms[j], nv, r);
        const D2 rij = eval_pair_d2(c.rho[ta * c.n_el + tj], D2(r, 1.0, 0.0));
        s_a[0] += rij.d * nv[0];
        s_a[1] += rij.d * nv[1];
        s_a[2] += rij.d * nv[2];
        if (o == a) continue;           // own periodic image: rigid, no contribution
        const D2 rji = ta == tj ? rij : eval_pair_d2(c.rho[tj * c.n_el + ta], D2(r, 1.0, 0.0));
        const D2 ph = eval_pair_d2(c.phi[ta * c.n_el + tj], D2(r, 1.0, 0.0));
        const double f1 = fp[a] * rij.d + fp[o] * rji.d + ph.d;
        const double f2 = fp[a] * rij.dd + fp[o] * rji.dd + ph.dd;
        double K[9];
        pair_kernel(nv, r, f1, f2, K);
        add_block(H, n, oa, oa, K, 1.0);
        add_block(H, n, oa, c.perm[o], K, -1.0);
    }
    // ---- embedding term, centre c = a:  F''_a g_a^a (x) g_a^b
    //      g_a^a = -(s_a - self images) ; handled uniformly through
    //      g_c^x = sum_p v_p (d_{x,o(p)} - d_{x,c})
    {
        const double w = fpp[a];
        for (int k = 0; k < cnt; ++k) {
            const uint32_t e = c.col[base + (size_t)k * 32u];
            const int j = (int)(e & TAB_COL_IDX_MASK), tj = (int)(e >> TAB_COL_TYPE_SHIFT);
            double nv[3], r;
            pair_geom(me, c.atoms[j], nv, r);
            const double d1 = eval_pair_d2(c.rho[ta * c.n_el + tj], D2(r, 1.0, 0.0)).d;
            double K[9];
#pragma unroll
            for (int al = 0; al < 3; ++al)
#pragma unroll
                for (int be = 0; be < 3; ++be) K[al * 3 + be] = s_a[al] * d1 * nv[be];
            add_block(H, n, oa, c.perm[owner_of(c, j)], K, -w);   // term3
        }
        double K[9];
#pragma unroll
        for (int al = 0; al < 3; ++al)
#pragma unroll
            for (int be = 0; be < 3; ++be) K[al * 3 + be] = s_a[al] * s_a[be];
        add_block(H, n, oa, oa, K, w);                            // term4
    }
    // ---- embedding term, centres c = owner of each entry of row a
    for (int k = 0; k < cnt; ++k) {
        const uint32_t e = c.col[base + (size_t)k * 32u];
        const int j = (int)(e & TAB_COL_IDX_MASK), tc = (int)(e >> TAB_COL_TYPE_SHIFT);
        const int cc = owner_of(c, j);
        if (cc >= c.n) continue;        // halo owner: not supported here
        double nv[3], r;
        pair_geom(me, c.atoms[j], nv, r);
        // reverse entry p = (c -> a): v_p = rho'_{c<-a}(r) * (-n)
        const double d1 = eval_pair_d2(c.rho[tc * c.n_el + ta], D2(r, 1.0, 0.0)).d;
        const double vp[3] = {-d1 * nv[0], -d1 * nv[1], -d1 * nv[2]};
        const double w = fpp[cc];
        const Atom4 pc = c.atoms[cc];
        const size_t cbase = (size_t)c.slice_ptr[cc >> 5] * 32u + (cc & 31);
        const int ccnt = c.counts[cc];
        double s_c[3] = {0, 0, 0};
        for (int q = 0; q < ccnt; ++q) {
            const uint32_t eq = c.col[cbase + (size_t)q * 32u];
            const int jq = (int)(eq & TAB_COL_IDX_MASK), tq = (int)(eq >> TAB_COL_TYPE_SHIFT);
            double nq[3], rq;
            pair_geom(pc, c.atoms[jq], nq, rq);
            const double dq = eval_pair_d2(c.rho[tc * c.n_el + tq], D2(rq, 1.0, 0.0)).d;
            double K[9];
#pragma unroll
            for (int al = 0; al < 3; ++al)
#pragma unroll
                for (int be = 0; be < 3; ++be) K[al * 3 + be] = vp[al] * dq * nq[be];
            add_block(H, n, oa, c.perm[owner_of(c, jq)], K, w);   // term1
            s_c[0] += dq * nq[0];
            s_c[1] += dq * nq[1];
            s_c[2] += dq * nq[2];
        }
        double K[9];
#pragma unroll
        for (int al = 0; al < 3; ++al)
#pragma unroll
            for (int be = 0; be < 3; ++be) K[al * 3 + be] = vp[al] * s_c[be];
        add_block(H, n, oa, c.perm[cc], K, -w);                   // term2
    }
}

// ---------------------------------------------------------------------------
// Elastic-constant op of the reference (nn/constraint/elastic.py:24-91):
//   C_ijkl = [ (d virial_ij / d h)^T h ]_kl / V / GPa ,  virial_ij = sum_p g_p,i D_p,j ,
// the derivative w.r.t. the lattice h taken at FIXED Cartesian positions: D_p = R_j - R_i + S_p h
// depends on h only through the image shift, d D_p,a / d h_mk = S_p,m d_ak.  With T_p = S_p h
// (the shift vector of the pair) and the pair-space second derivatives of the EAM energy
//   d g_p,i / d D_q,k = d_pq A_p,ik + [same centre c] F''_c v_p,i v_q,k ,
//   A_p = F'_c K^rho_p + 1/2 K^phi_p ,  v_p = rho'_p n_p ,  g_p = (F'_c rho'_p + 1/2 phi'_p) n_p
// (directed pairs p of centre c):
//   V GPa C_ijkl = sum_p A_p,ik T_p,l D_p,j + sum_c F''_c (sum_p v_p,i D_p,j)(sum_q v_q,k T_q,l)
//                  + d_jk sum_p g_p,i T_p,l .
// One thread per centre; out[c][vi * 6 + vj] = its share in Voigt pairs (xx yy zz yz xz xy).
// ---------------------------------------------------------------------------
struct ShiftCtx {           // what turns a list entry into its integer image shift S_p
    const int *s0;          // packed wrap shift per atom, caller order
    const int *ghost_S;     // packed image shift per ghost record
    double h[9];            // lattice, rows = vectors
};

__global__ void k_elastic(HessCtx c, ShiftCtx sc, const double *__restrict__ fp,
                          const double *__restrict__ fpp, double *__restrict__ out) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= c.n) return;
    const int VI[6] = {0, 1, 2, 1, 0, 0}, VJ[6] = {0, 1, 2, 2, 2, 1};
    const Atom4 me = c.atoms[a];
    const int ta = c.types_ext[a];
    const size_t base = (size_t)c.slice_ptr[a >> 5] * 32u + (a & 31);
    const int cnt = c.counts[a];
    double C[36];
    for (int q = 0; q < 36; ++q) C[q] = 0.0;
    double U[9], W[9], G[9];        // sum v_i D_j, sum v_k T_l, sum g_i T_l
    for (int q = 0; q < 9; ++q) U[q] = W[q] = G[q] = 0.0;
    for (int k = 0; k < cnt; ++k) {
        const uint32_t e = c.col[base + (size_t)k * 32u];
        const int j = (int)(e & TAB_COL_IDX_MASK), tj = (int)(e >> TAB_COL_TYPE_SHIFT);
        const Atom4 aj = c.atoms[j];
        double nv[3], r;
        pair_geom(me, aj, nv, r);
        const double D[3] = {aj.x - me.x, aj.y - me.y, aj.z - me.z};
        const D2 rho = eval_pair_d2(c.rho[ta * c.n_el + tj], D2(r, 1.0, 0.0));
        const D2 ph = eval_pair_d2(c.phi[ta * c.n_el + tj], D2(r, 1.0, 0.0));
        const double v[3] = {rho.d * nv[0], rho.d * nv[1], rho.d * nv[2]};
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int jj = 0; jj < 3; ++jj) U[i * 3 + jj] += v[i] * D[jj];
        // S_p of the pair w.r.t. the CALLER's positions (the reference's n1): the image shift of
        // the ghost record minus the wrap shifts the library applied to the two atoms
        int Sa = 0, Sb = 0, Sc = 0, ia, ib, ic, ja, jb, jc;
        if (j >= c.n_loc) tab_unpack_shift(sc.ghost_S[j - c.n_loc], Sa, Sb, Sc);
        tab_unpack_shift(sc.s0[c.perm[a]], ia, ib, ic);
        tab_unpack_shift(sc.s0[c.perm[owner_of(c, j)]], ja, jb, jc);
        Sa += ia - ja;
        Sb += ib - jb;
        Sc += ic - jc;
        if (Sa == 0 && Sb == 0 && Sc == 0) continue;      // the pair does not move with h
        const double T[3] = {Sa * sc.h[0] + Sb * sc.h[3] + Sc * sc.h[6],
                             Sa * sc.h[1] + Sb * sc.h[4] + Sc * sc.h[7],
                             Sa * sc.h[2] + Sb * sc.h[5] + Sc * sc.h[8]};
        const double g1 = fp[a] * rho.d + 0.5 * ph.d;
        double A[9];
        pair_kernel(nv, r, g1, fp[a] * rho.dd + 0.5 * ph.dd, A);
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int l = 0; l < 3; ++l) {
                W[i * 3 + l] += v[i] * T[l];
                G[i * 3 + l] += g1 * nv[i] * T[l];
            }
        for (int vi = 0; vi < 6; ++vi)
            for (int vj = 0; vj < 6; ++vj)
                C[vi * 6 + vj] += A[VI[vi] * 3 + VI[vj]] * T[VJ[vj]] * D[VJ[vi]];
    }
    const double w = fpp[a];
    for (int vi = 0; vi < 6; ++vi)
        for (int vj = 0; vj < 6; ++vj) {
            const int i = VI[vi], jj = VJ[vi], kk = VI[vj], l = VJ[vj];
            double x = C[vi * 6 + vj] + w * U[i * 3 + jj] * W[kk * 3 + l];
            if (jj == kk) x += G[i * 3 + l];
            out[(size_t)c.perm[a] * 36 + vi * 6 + vj] = x;
        }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
struct tab_model_view {      // layout prefix of tab_model (eam.cu)
    int family, kind, n_el;
};
int tab_eam_tables(tab_model *m, const tab_fn **rho, const tab_fn **phi,
                   const tab_fn **embed, int *n_el, int *kind);   // eam.cu
const double *tab_eam_pool(tab_model *m);                          // eam.cu

extern "C" int tab_eam_hessian(tab_model *m, tab_nbr *nbr, double *d_hessian,
                               void *stream) {
    if (!m || !nbr || !d_hessian) {
        tab_set_error("tab_eam_hessian: bad argument");
        return TAB_EINVAL;
    }
    if (!nbr->built) {
        tab_set_error("tab_eam_hessian before tab_nbr_build");
        return TAB_ESTATE;
    }
    if (nbr->n_halo > 0) {
        tab_set_error("tab_eam_hessian: halo atoms (domain decomposition) not supported");
        return TAB_EUNSUPPORTED;
    }
    if (nbr->skin_built > 0.0) {
        tab_set_error("tab_eam_hessian: the lists carry a skin (entries beyond rc); build with "
                      "skin = 0");
        return TAB_ESTATE;
    }
    if (nbr->n_struct > 0) {
        tab_set_error("tab_eam_hessian: batch handles are not supported (one structure per call)");
        return TAB_EUNSUPPORTED;
    }
    cudaStream_t st = (cudaStream_t)stream;
    HessCtx c;
    int kind = 0;
    TAB_TRY(tab_eam_tables(m, &c.rho, &c.phi, &c.embed, &c.n_el, &kind));
    if (kind == TAB_EAM_ADP) {
        tab_set_error("tab_eam_hessian: ADP not supported");
        return TAB_EUNSUPPORTED;
    }
    c.n = nbr->n;
    c.n_loc = nbr->n_loc;
    c.atoms = nbr->atoms.as<Atom4>();
    c.types_ext = nbr->types_ext.as<uint8_t>();
    c.counts = nbr->counts.as<int>();
    c.slice_ptr = nbr->slice_ptr.as<uint32_t>();
    c.col = nbr->col.as<uint32_t>();
    c.ghost_owner = nbr->ghost_owner.as<int>();
    c.perm = nbr->perm.as<int>();
    const int n = nbr->n;
    TAB_TRY(nbr->rho.ensure(sizeof(double) * 2 * (size_t)n));
    double *fp = nbr->rho.as<double>(), *fpp = fp + n;
    const double *pool = tab_eam_pool(m);
    TAB_CUDA(cudaMemcpyToSymbolAsync(g_hess_pool, &pool, sizeof(pool), 0,
                                     cudaMemcpyHostToDevice, st));
    TAB_CUDA(cudaMemsetAsync(d_hessian, 0, sizeof(double) * 9 * (size_t)n * n, st));
    k_hess_rho<<<(n + 127) / 128, 128, 0, st>>>(c, fp, fpp);
    TAB_LAUNCH_CHECK();
    k_hessian<<<(n + 63) / 64, 64, 0, st>>>(c, fp, fpp, d_hessian);
    TAB_LAUNCH_CHECK();
    return TAB_OK;
}

extern "C" int tab_eam_elastic(tab_model *m, tab_nbr *nbr, double *d_out, void *stream) {
    if (!m || !nbr || !d_out) {
        tab_set_error("tab_eam_elastic: bad argument");
        return TAB_EINVAL;
    }
    if (!nbr->built) {
        tab_set_error("tab_eam_elastic before tab_nbr_build");
        return TAB_ESTATE;
    }
    if (nbr->n_halo > 0 || nbr->n_struct > 0) {
        tab_set_error("tab_eam_elastic: one undecomposed structure per call");
        return TAB_EUNSUPPORTED;
    }
    if (nbr->skin_built > 0.0) {
        tab_set_error("tab_eam_elastic: the lists carry a skin (entries beyond rc); build with "
                      "skin = 0");
        return TAB_ESTATE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    HessCtx c;
    int kind = 0;
    TAB_TRY(tab_eam_tables(m, &c.rho, &c.phi, &c.embed, &c.n_el, &kind));
    if (kind == TAB_EAM_ADP) {
        tab_set_error("tab_eam_elastic: ADP not supported");
        return TAB_EUNSUPPORTED;
    }
    c.n = nbr->n;
    c.n_loc = nbr->n_loc;
    c.atoms = nbr->atoms.as<Atom4>();
    c.types_ext = nbr->types_ext.as<uint8_t>();
    c.counts = nbr->counts.as<int>();
    c.slice_ptr = nbr->slice_ptr.as<uint32_t>();
    c.col = nbr->col.as<uint32_t>();
    c.ghost_owner = nbr->ghost_owner.as<int>();
    c.perm = nbr->perm.as<int>();
    const int n = nbr->n;
    TAB_TRY(nbr->rho.ensure(sizeof(double) * 2 * (size_t)n));
    double *fp = nbr->rho.as<double>(), *fpp = fp + n;
    const double *pool = tab_eam_pool(m);
    TAB_CUDA(cudaMemcpyToSymbolAsync(g_hess_pool, &pool, sizeof(pool), 0,
                                     cudaMemcpyHostToDevice, st));
    k_hess_rho<<<(n + 127) / 128, 128, 0, st>>>(c, fp, fpp);
    TAB_LAUNCH_CHECK();
    ShiftCtx sc;
    sc.s0 = nbr->s0.as<int>();
    sc.ghost_S = nbr->ghost_S.as<int>();
    memcpy(sc.h, nbr->grid.h, sizeof(sc.h));
    k_elastic<<<(n + 63) / 64, 64, 0, st>>>(c, sc, fp, fpp, d_out);
    TAB_LAUNCH_CHECK();
    return TAB_OK;
}
