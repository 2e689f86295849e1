// Temperature-dependent heads of the finite-temperature AtomicNN, fused in one kernel
// (reference nn/atomic/finite_temperature.py:92-304, nn/atomic/special/beryllium.py:23-77):
//
//   per atom (element e, descriptors x [dim], electron temperature T):
//     x'  = (xhi - x) / (xhi - xlo)                       min-max map (atomic.py:157-195)
//     H   = net_H(x')                                     last layer linear, width nH
//     Ht  = [H, T]
//     s   = net_S(Ht), u = net_U(Ht)                      last layers linear, width 1
//     S   = s | s T (Sommerfeld) | S_fe(T) softplus(s) (Be)
//     F   = u - T S
//   out:  U = u, S, F per atom and dF/dx (what the force kernels contract with dG/dR)
//
// The reference runs three chains of 1x1 convolutions and leaves dF/dx to tf.gradients.  Here
// one block takes a tile of A atoms of ONE element: forward through the three networks,
// the entropy model, and the reverse pass through S, U and H, with every activation and
// activation derivative of the tile kept in shared memory.
//
// Work split: a thread owns output column o (forward) or input row k (backward) of a layer and
// carries A accumulators, so one weight load feeds A FMAs; the tile's activations are stored
// [feature][atom] (atoms innermost) and read as 16-byte shared-memory broadcasts.  Weights come
// from L2 / L1 (a layer is read once per tile, not once per atom).  Atoms are bucketed by
// element on the device (k_td_bucket) so that tiles are element-pure; the order inside a bucket
// is arbitrary and irrelevant (atoms are independent).
#include <cstring>

#include "tab200.h"
#include "tab_internal.h"
#include "mlp_act.cuh"

#define TD_THREADS 128
#define TD_MAX_LAYERS 8
#define TD_MAX_WIDTH 1024
#define TD_SMEM_MAX (200 * 1024)

struct TdNet {
    int n_layers, act, resnet, has_out_bias;
    int in[TD_MAX_LAYERS], out[TD_MAX_LAYERS];
    long long w_off[TD_MAX_LAYERS], b_off[TD_MAX_LAYERS];      // into the blob
    int h_off[TD_MAX_LAYERS + 1];      // per-atom offsets of the activations h[0..n_layers]
    int dz_off[TD_MAX_LAYERS];         // ... of the activation derivatives of the hidden layers
};

struct TdElem {
    TdNet H, S, U;
    int has_minmax;
    long long xlo_off, xhi_off;
};

struct TdPlan {
    int dim, nH, algo, special;
    int d_off[3];          // per-atom offsets of three scratch vectors of width `wmax`
    int per_atom;          // values per atom in shared memory
};

struct tab_td {
    int n_el = 0;
    TdPlan plan;
    TdElem elem[TAB_MAX_ELEMENTS];
    DevBuf blob, blobf, elems_dev;     // weights as float64 and float32, TdElem table
    DevBuf list, cnt;                  // element buckets
    int tile[2] = {1, 1};              // atoms per block for float64 / float32
};

__global__ void k_td_bucket(int n, int n_el, const int32_t *__restrict__ types,
                            int *__restrict__ cnt, int *__restrict__ list) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int t = types[i];
    if (t < 0 || t >= n_el) return;
    const int slot = atomicAdd(&cnt[t], 1);
    list[(size_t)t * n + slot] = i;
}

// A consecutive values of the tile ([feature][atom] layout) as 16-byte shared-memory loads
template <typename Real, int A>
__device__ __forceinline__ void td_load(const Real *p, Real (&v)[A]) {
    if constexpr (A == 1) {
        v[0] = p[0];
    } else if constexpr (sizeof(Real) * A >= 16) {
        constexpr int NV = (int)(sizeof(Real) * A / 16);
        const uint4 *q = reinterpret_cast<const uint4 *>(p);
        uint4 raw[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i) raw[i] = q[i];
        memcpy(v, raw, sizeof(raw));
    } else {
        const uint2 raw = *reinterpret_cast<const uint2 *>(p);
        memcpy(v, &raw, sizeof(raw));
    }
}

// forward through one network: h[0] (already in shared memory) -> h[n_layers]
template <typename Real, int A>
__device__ __forceinline__ void td_forward(const TdNet &N, const Real *__restrict__ blob,
                                           Real *sm) {
    const int tid = threadIdx.x;
    for (int l = 0; l < N.n_layers; ++l) {
        const Real *W = blob + N.w_off[l];
        const Real *bb = blob + N.b_off[l];
        const int ni = N.in[l], no = N.out[l];
        const Real *hin = sm + (size_t)N.h_off[l] * A;
        Real *hout = sm + (size_t)N.h_off[l + 1] * A;
        const bool last = l == N.n_layers - 1;
        Real *dz = last ? nullptr : sm + (size_t)N.dz_off[l] * A;
        const bool res = !last && l > 0 && N.resnet && no == ni;
        for (int o = tid; o < no; o += TD_THREADS) {
            const Real b0 = (last && !N.has_out_bias) ? Real(0) : bb[o];
            Real acc[A];
#pragma unroll
            for (int a = 0; a < A; ++a) acc[a] = b0;
#pragma unroll 4
            for (int k = 0; k < ni; ++k) {
                const Real w = W[(size_t)k * no + o];
                Real hk[A];
                td_load<Real, A>(hin + (size_t)k * A, hk);
#pragma unroll
                for (int a = 0; a < A; ++a) acc[a] = fma(hk[a], w, acc[a]);
            }
#pragma unroll
            for (int a = 0; a < A; ++a) {
                if (last) {
                    hout[(size_t)o * A + a] = acc[a];
                } else {
                    Real d;
                    Real y = act_fn<Real>(N.act, acc[a], d);
                    dz[(size_t)o * A + a] = d;
                    if (res) y += hin[(size_t)o * A + a];
                    hout[(size_t)o * A + a] = y;
                }
            }
        }
        __syncthreads();
    }
}

// reverse pass: seed = dF/d(output) [n_out][A] -> dF/d(input) [n_in][A]; returns the buffer
// (one of d0 / d1) that holds the result; `tmp` is scratch.  seed may alias none of them.
template <typename Real, int A>
__device__ __forceinline__ Real *td_backward(const TdNet &N, const Real *__restrict__ blob,
                                             Real *sm, const Real *seed, Real *d0, Real *d1,
                                             Real *tmp) {
    const int tid = threadIdx.x;
    Real *cur = d0, *nxt = d1;
    for (int l = N.n_layers - 1; l >= 0; --l) {
        const Real *W = blob + N.w_off[l];
        const int ni = N.in[l], no = N.out[l];
        const bool last = l == N.n_layers - 1;
        const bool res = !last && l > 0 && N.resnet && no == ni;
        const Real *t = seed;
        if (!last) {
            // t = dF/dh_{l+1} * act'(z_l)
            const Real *dz = sm + (size_t)N.dz_off[l] * A;
            for (int q = tid; q < no * A; q += TD_THREADS) tmp[q] = cur[q] * dz[q];
            __syncthreads();
            t = tmp;
        }
        Real *dst = last ? cur : nxt;
        for (int k = tid; k < ni; k += TD_THREADS) {
            Real s[A];
#pragma unroll
            for (int a = 0; a < A; ++a) s[a] = Real(0);
            const Real *Wk = W + (size_t)k * no;
#pragma unroll 4
            for (int o = 0; o < no; ++o) {
                const Real w = Wk[o];
                Real to[A];
                td_load<Real, A>(t + (size_t)o * A, to);
#pragma unroll
                for (int a = 0; a < A; ++a) s[a] = fma(to[a], w, s[a]);
            }
#pragma unroll
            for (int a = 0; a < A; ++a)
                dst[(size_t)k * A + a] = s[a] + (res ? cur[(size_t)k * A + a] : Real(0));
        }
        __syncthreads();
        if (!last) {
            Real *sw = cur;
            cur = nxt;
            nxt = sw;
        }
    }
    return cur;
}

template <typename Real, int A>
__global__ void __launch_bounds__(TD_THREADS)
k_td_heads(int n, TdPlan P, const TdElem *__restrict__ elems, const Real *__restrict__ blob,
           const int *__restrict__ list, const int *__restrict__ cnt,
           const double *__restrict__ G, const double *__restrict__ T,
           double *__restrict__ U, double *__restrict__ S, double *__restrict__ F,
           double *__restrict__ dFdG) {
    extern __shared__ __align__(16) unsigned char td_smem[];
    Real *sm = reinterpret_cast<Real *>(td_smem);
    __shared__ int ids[A];
    __shared__ Real temp[A];
    const int tid = threadIdx.x;
    const int e = blockIdx.y;
    const int t0 = blockIdx.x * A;
    const int m = cnt[e];
    if (t0 >= m) return;
    const TdElem &E = elems[e];
    if (tid < A) {
        const int id = t0 + tid < m ? list[(size_t)e * n + t0 + tid] : -1;
        ids[tid] = id;
        temp[tid] = id >= 0 ? (Real)T[id] : Real(0);
    }
    __syncthreads();
    const int dim = P.dim, nH = P.nH;
    // inputs (atoms beyond the bucket: zeros, never stored)
    {
        Real *x = sm + (size_t)E.H.h_off[0] * A;
        for (int q = tid; q < dim * A; q += TD_THREADS) {
            const int k = q / A, a = q % A;
            const int id = ids[a];
            Real v = Real(0);
            if (id >= 0) {
                v = (Real)G[(size_t)id * dim + k];
                if (E.has_minmax) {
                    const Real lo = blob[E.xlo_off + k], hi = blob[E.xhi_off + k];
                    const Real den = hi - lo;
                    v = den != Real(0) ? (hi - v) / den : Real(0);
                }
            }
            x[q] = v;
        }
    }
    __syncthreads();
    td_forward<Real, A>(E.H, blob, sm);
    // Ht = [H, T]: the networks S and U read h_H[last] with one more row
    Real *Ht = sm + (size_t)E.H.h_off[E.H.n_layers] * A;
    if (tid < A) Ht[(size_t)nH * A + tid] = temp[tid];
    __syncthreads();
    td_forward<Real, A>(E.S, blob, sm);
    td_forward<Real, A>(E.U, blob, sm);
    Real *s_out = sm + (size_t)E.S.h_off[E.S.n_layers] * A;
    Real *u_out = sm + (size_t)E.U.h_off[E.U.n_layers] * A;
    if (tid < A) {
        const Real t = temp[tid];
        const Real s = s_out[tid], u = u_out[tid];
        Real Sv, dSds;
        if (P.special == 1) {
            // beryllium.py:23-77: fitted free-electron entropy times softplus(s)
            const Real r = fmax(Real(1) - Real(1.45) * t, Real(0));
            const Real ft = r * r;
            const Real base = Real(-0.5718444) * t * t * ft + Real(0.83744317) * t +
                              Real(-0.2110962) * (Real(1) - ft);
            Real sg;
            const Real sp = act_fn<Real>(0, s, sg);
            Sv = base * sp;
            dSds = base * sg;
        } else if (P.algo == 1) {        // Sommerfeld (finite_temperature.py:160-163)
            Sv = s * t;
            dSds = t;
        } else {
            Sv = s;
            dSds = Real(1);
        }
        const int id = ids[tid];
        if (id >= 0) {
            U[id] = (double)u;
            S[id] = (double)Sv;
            F[id] = (double)(u - t * Sv);
        }
        // seeds of the reverse pass: dF/du = 1, dF/ds = -T dS/ds
        s_out[tid] = -t * dSds;
        u_out[tid] = Real(1);
    }
    __syncthreads();
    Real *d0 = sm + (size_t)P.d_off[0] * A, *d1 = sm + (size_t)P.d_off[1] * A;
    Real *tmp = sm + (size_t)P.d_off[2] * A;
    // dF/dHt through S, kept in the input buffer of the U / S networks' first layer is not
    // possible (U still needs Ht's derivatives only): accumulate in Ht itself after both passes
    Real *gS = td_backward<Real, A>(E.S, blob, sm, s_out, d0, d1, tmp);
    // stash the S part in the seed area of H (the dz of H are still needed, Ht is not)
    for (int q = tid; q < nH * A; q += TD_THREADS) Ht[q] = gS[q];
    __syncthreads();
    Real *gU = td_backward<Real, A>(E.U, blob, sm, u_out, d0, d1, tmp);
    for (int q = tid; q < nH * A; q += TD_THREADS) Ht[q] += gU[q];
    __syncthreads();
    Real *gx = td_backward<Real, A>(E.H, blob, sm, Ht, d0, d1, tmp);
    for (int q = tid; q < dim * A; q += TD_THREADS) {
        const int k = q / A, a = q % A;
        const int id = ids[a];
        if (id < 0) continue;
        Real v = gx[q];
        if (E.has_minmax) {
            const Real den = blob[E.xhi_off + k] - blob[E.xlo_off + k];
            v = den != Real(0) ? -v / den : Real(0);
        }
        dFdG[(size_t)id * dim + k] = (double)v;
    }
}

// -- host side -------------------------------------------------------------------------
static int td_fill_net(TdNet &N, const tab_mlp_desc &q, double *host, size_t &off,
                       int &per_atom, int h0_off, int &wmax, const char *name, int in_expect,
                       int out_expect) {
    memset(&N, 0, sizeof(N));
    if (q.n_layers < 1 || q.n_layers > TD_MAX_LAYERS) {
        tab_set_error("tab_td_create: %s: 1..%d layers", name, TD_MAX_LAYERS);
        return TAB_EINVAL;
    }
    if (q.sizes[0] != in_expect || (out_expect > 0 && q.sizes[q.n_layers] != out_expect)) {
        tab_set_error("tab_td_create: %s maps %d -> %d values, expected %d -> %d", name,
                      q.sizes[0], q.sizes[q.n_layers], in_expect, out_expect);
        return TAB_EINVAL;
    }
    N.n_layers = q.n_layers;
    N.act = q.activation;
    N.resnet = q.use_resnet_dt;
    N.has_out_bias = q.output_bias;
    N.h_off[0] = h0_off;
    for (int l = 0; l < q.n_layers; ++l) {
        const int ni = q.sizes[l], no = q.sizes[l + 1];
        if (ni < 1 || no < 1 || ni > TD_MAX_WIDTH || no > TD_MAX_WIDTH || !q.weights[l]) {
            tab_set_error("tab_td_create: %s layer %d: widths 1..%d and weights required", name,
                          l, TD_MAX_WIDTH);
            return TAB_EINVAL;
        }
        N.in[l] = ni;
        N.out[l] = no;
        wmax = ni > wmax ? ni : wmax;
        wmax = no > wmax ? no : wmax;
        if (host) {
            memcpy(host + off, q.weights[l], sizeof(double) * ni * no);
            if (q.biases[l]) memcpy(host + off + (size_t)ni * no, q.biases[l], sizeof(double) * no);
            else memset(host + off + (size_t)ni * no, 0, sizeof(double) * no);
        }
        N.w_off[l] = (long long)off;
        N.b_off[l] = (long long)(off + (size_t)ni * no);
        off += (size_t)ni * no + no;
        // the output buffer of the last layer of H carries one more row (the temperature)
        N.h_off[l + 1] = per_atom;
        per_atom += no + 1;
        if (l < q.n_layers - 1) {
            N.dz_off[l] = per_atom;
            per_atom += no;
        }
    }
    return TAB_OK;
}

extern "C" int tab_td_create(tab_td **out, const tab_td_desc *d) {
    if (!out || !d || !d->H || !d->S || !d->U || d->n_elements < 1 ||
        d->n_elements > TAB_MAX_ELEMENTS || d->dim < 1 || d->dim > TD_MAX_WIDTH) {
        tab_set_error("tab_td_create: invalid descriptor");
        return TAB_EINVAL;
    }
    tab_td *m = new tab_td();
    m->n_el = d->n_elements;
    const int nH = d->H[0].sizes[d->H[0].n_layers];
    // two passes: sizes, then fill
    size_t total = 0;
    int per_atom_max = 0, wmax = d->dim;
    double *host = nullptr;
    for (int pass = 0; pass < 2; ++pass) {
        size_t off = 0;
        for (int e = 0; e < d->n_elements; ++e) {
            TdElem &E = m->elem[e];
            int per_atom = d->dim;          // h_H[0] at offset 0
            int rc = td_fill_net(E.H, d->H[e], host, off, per_atom, 0, wmax, "H", d->dim, nH);
            const int ht = E.H.h_off[E.H.n_layers];
            if (rc == TAB_OK)
                rc = td_fill_net(E.S, d->S[e], host, off, per_atom, ht, wmax, "S", nH + 1, 1);
            if (rc == TAB_OK)
                rc = td_fill_net(E.U, d->U[e], host, off, per_atom, ht, wmax, "U", nH + 1, 1);
            if (rc != TAB_OK) {
                delete[] host;
                delete m;
                return rc;
            }
            E.has_minmax = (d->H[e].xlo && d->H[e].xhi) ? 1 : 0;
            E.xlo_off = (long long)off;
            E.xhi_off = (long long)(off + d->dim);
            if (host) {
                memset(host + off, 0, sizeof(double) * 2 * d->dim);
                if (E.has_minmax) {
                    memcpy(host + off, d->H[e].xlo, sizeof(double) * d->dim);
                    memcpy(host + off + d->dim, d->H[e].xhi, sizeof(double) * d->dim);
                }
            }
            off += 2 * (size_t)d->dim;
            per_atom_max = per_atom > per_atom_max ? per_atom : per_atom_max;
        }
        total = off;
        if (pass == 0) host = new double[total + 1];
    }
    TdPlan &P = m->plan;
    P.dim = d->dim;
    P.nH = nH;
    P.algo = d->algo;
    P.special = d->special;
    wmax += 1;
    for (int k = 0; k < 3; ++k) P.d_off[k] = per_atom_max + k * wmax;
    P.per_atom = per_atom_max + 3 * wmax;
    // atoms per block: the largest of 8, 4, 2, 1 whose tile fits in shared memory
    for (int p = 0; p < 2; ++p) {
        const size_t w = p == 0 ? 8 : 4;
        int A = 8;
        while (A > 1 && (size_t)P.per_atom * A * w > TD_SMEM_MAX) A >>= 1;
        if ((size_t)P.per_atom * A * w > TD_SMEM_MAX) {
            tab_set_error("tab_td_create: networks need %zu bytes of shared memory per atom",
                          (size_t)P.per_atom * w);
            delete[] host;
            delete m;
            return TAB_EUNSUPPORTED;
        }
        m->tile[p] = A;
    }
    float *hostf = new float[total + 1];
    for (size_t k = 0; k < total; ++k) hostf[k] = (float)host[k];
    int rc = m->blob.ensure(total * 8);
    if (rc == TAB_OK) rc = m->blobf.ensure(total * 4);
    if (rc == TAB_OK) rc = m->elems_dev.ensure(sizeof(TdElem) * m->n_el);
    if (rc == TAB_OK) rc = m->cnt.ensure(sizeof(int) * TAB_MAX_ELEMENTS);
    if (rc == TAB_OK) {
        cudaError_t e = cudaMemcpy(m->blob.p, host, total * 8, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(m->blobf.p, hostf, total * 4, cudaMemcpyHostToDevice);
        if (e == cudaSuccess)
            e = cudaMemcpy(m->elems_dev.p, m->elem, sizeof(TdElem) * m->n_el,
                           cudaMemcpyHostToDevice);
        if (e != cudaSuccess) {
            tab_set_error("tab_td_create: cudaMemcpy -> %s", cudaGetErrorString(e));
            rc = TAB_ECUDA;
        }
    }
    delete[] host;
    delete[] hostf;
    if (rc != TAB_OK) {
        delete m;
        return rc;
    }
    *out = m;
    return TAB_OK;
}

extern "C" int tab_td_free(tab_td *m) {
    if (!m) return TAB_OK;
    DevBuf *bufs[] = {&m->blob, &m->blobf, &m->elems_dev, &m->list, &m->cnt};
    for (DevBuf *b : bufs) b->release();
    delete m;
    return TAB_OK;
}

template <typename Real, int A>
static int td_launch(tab_td *m, int n, const Real *blob, const double *G, const double *T,
                     double *U, double *S, double *F, double *dFdG, cudaStream_t st) {
    const size_t smem = (size_t)m->plan.per_atom * A * sizeof(Real);
    TAB_CUDA(cudaFuncSetAttribute(k_td_heads<Real, A>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)((n + A - 1) / A), (unsigned)m->n_el);
    k_td_heads<Real, A><<<grid, TD_THREADS, smem, st>>>(
        n, m->plan, m->elems_dev.as<TdElem>(), blob, m->list.as<int>(), m->cnt.as<int>(), G, T,
        U, S, F, dFdG);
    TAB_LAUNCH_CHECK();
    return TAB_OK;
}

extern "C" int tab_td_eval(tab_td *m, int32_t n, const int32_t *d_types, const double *d_G,
                           const double *d_T, int32_t precision, double *d_U, double *d_S,
                           double *d_F, double *d_dFdG, void *stream) {
    if (!m || n < 0 || !d_types || !d_G || !d_T || !d_U || !d_S || !d_F || !d_dFdG) {
        tab_set_error("tab_td_eval: null argument");
        return TAB_EINVAL;
    }
    if (n == 0) return TAB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    TAB_TRY(m->list.ensure(sizeof(int) * (size_t)n * m->n_el));
    TAB_CUDA(cudaMemsetAsync(m->cnt.p, 0, sizeof(int) * TAB_MAX_ELEMENTS, st));
    k_td_bucket<<<(n + 255) / 256, 256, 0, st>>>(n, m->n_el, d_types, m->cnt.as<int>(),
                                                 m->list.as<int>());
    TAB_LAUNCH_CHECK();
    const bool f64 = precision == TAB_PRECISION_HIGH;
    const int A = m->tile[f64 ? 0 : 1];
#define TD_GO(R, AA, B)                                                                        \
    case AA:                                                                                   \
        return td_launch<R, AA>(m, n, B, d_G, d_T, d_U, d_S, d_F, d_dFdG, st);
    if (f64) {
        switch (A) {
            TD_GO(double, 8, m->blob.as<double>())
            TD_GO(double, 4, m->blob.as<double>())
            TD_GO(double, 2, m->blob.as<double>())
            TD_GO(double, 1, m->blob.as<double>())
        }
    } else {
        switch (A) {
            TD_GO(float, 8, m->blobf.as<float>())
            TD_GO(float, 4, m->blobf.as<float>())
            TD_GO(float, 2, m->blobf.as<float>())
            TD_GO(float, 1, m->blobf.as<float>())
        }
    }
#undef TD_GO
    return TAB_ESTATE;
}
