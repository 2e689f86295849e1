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
// one block takes a tile of A atoms of ONE element through a host-built SCHEDULE of small
// GEMM operations (forward H, S, U; the entropy model; reverse S, U, H) whose operands live
// in a shared-memory pool of rows [feature][atom]:
//
//   * weights never go through registers-from-L2: every matrix is streamed in chunks of rows
//     by the TMA engine (cp.async.bulk global -> shared, completion on an mbarrier), two
//     chunks in flight (one being consumed, one landing), the first chunk of the NEXT
//     operation already under way while the current one runs its epilogue;
//   * a thread owns output column o and carries A accumulators: per k one conflict-free
//     shared-memory weight load and A / 2 (float64) 16-byte broadcast loads feed A FMAs;
//   * the first layers of S and U read the same input: they are ONE operation on the
//     column-concatenated matrix [W_S | W_U], and so is their reverse pass, which then yields
//     the SUM dF/dHt directly; the reverse pass uses transposed copies of the matrices, so it
//     is the same GEMM routine (and skips the temperature column, whose derivative nobody
//     needs);
//   * atoms are bucketed by element on the device (k_td_bucket) so tiles are element-pure;
//     the order inside a bucket is arbitrary and irrelevant (atoms are independent).
#include <cstdlib>
#include <cstring>
#include <vector>

#include "tab200.h"
#include "tab_internal.h"
#include "mlp_act.cuh"

#define TD_COLS 128                 // threads of one atom group: one per output column
#define TD_MAX_LAYERS 8
#define TD_MAX_WIDTH 256            // 2 columns per thread
#define TD_MAX_OPS 64
#define TD_STAGES 4                 // weight chunks in the ring (TD_STAGES - 1 in flight)
#ifndef TD_WBUF_BYTES
#define TD_WBUF_BYTES (16 * 1024)   // one weight chunk, float64 (float32: half)
#endif
#define TD_SMEM_MAX (220 * 1024)
#define TD_SPIN_LIMIT (1u << 24)

enum { TD_FWD_HIDDEN = 0, TD_FWD_LAST = 1, TD_BWD = 2 };

// one GEMM of the schedule: out[o][a] = sum_k in[k][a] * W[k][o]  (+ epilogue)
struct TdOp {
    long long w_off;      // W [ni][no] row-major in the blob (16-byte aligned); reverse ops: W^T
    long long b_off;      // bias [no] (forward ops)
    int ni, no;
    int in_off, out_off;  // pool rows
    int kind, act, has_bias;
    int dz_off;           // forward hidden: act'(z) goes here; reverse: act' to fold into t
    int res_off;          // rows added to the result (resnet link / accumulation) or -1
    int t_off;            // reverse: out * act' goes here (input of the next reverse op) or -1
};

struct TdElem {
    int n_ops, n_fwd;
    int has_minmax;
    int x_off, ht_off, s_off, u_off, seed_s_off, seed_u_off, dx_off;
    long long xlo_off, xhi_off;
    TdOp ops[TD_MAX_OPS];
};

struct TdPlan {
    int dim, nH, algo, special;
    int pool_rows;
};

struct tab_td {
    int n_el = 0;
    TdPlan plan;
    std::vector<TdElem> elem;
    DevBuf blob, blobf, elems_dev;     // weights as float64 and float32, TdElem table
    DevBuf list, cnt, status;          // element buckets; status: set when a TMA wait timed out
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

// -- TMA / mbarrier primitives ---------------------------------------------------------
__device__ __forceinline__ uint32_t td_smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void td_mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(td_smem_u32(bar)), "r"(count)
                 : "memory");
}
__device__ __forceinline__ void td_fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// arm the barrier with the byte count and start the bulk copy global -> shared
__device__ __forceinline__ void td_bulk_load(void *dst, const void *src, uint32_t bytes,
                                             uint64_t *bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(td_smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(td_smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(td_smem_u32(bar))
        : "memory");
}
// bounded wait: a mistake in the schedule must not hang the GPU
__device__ __forceinline__ bool td_mbar_wait(uint64_t *bar, uint32_t parity) {
    for (uint32_t spin = 0; spin < TD_SPIN_LIMIT; ++spin) {
        uint32_t ok;
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}"
            : "=r"(ok)
            : "r"(td_smem_u32(bar)), "r"(parity)
            : "memory");
        if (ok) return true;
    }
    return false;
}

// the activation as a real function: inlined into the unrolled epilogues (A / NG copies of eight
// activation kinds per operation kind) it made 380 KB of code and 9 % instruction-fetch stalls
template <typename Real>
__device__ __noinline__ Real td_act(int kind, Real z, Real &d) {
    return act_fn<Real>(kind, z, d);
}

// A consecutive values of a pool row as 16-byte shared-memory loads
template <typename Real, int A>
__device__ __forceinline__ void td_load(const Real *p, Real (&v)[A]) {
    if constexpr (sizeof(Real) * A >= 16) {
        constexpr int NV = (int)(sizeof(Real) * A / 16);
        const uint4 *q = reinterpret_cast<const uint4 *>(p);
        uint4 raw[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i) raw[i] = q[i];
        memcpy(v, raw, sizeof(raw));
    } else {
#pragma unroll
        for (int a = 0; a < A; ++a) v[a] = p[a];
    }
}

// A pool row holds the A atoms of one feature plus 16 bytes of padding: threads that own
// consecutive features then hit different banks when they read or write their rows (without
// it the row stride is a multiple of 128 bytes: a 32-way conflict in every epilogue).
template <typename Real, int A>
struct TdRow {
    static constexpr int RS = A + (int)(16 / sizeof(Real));
};

template <typename Real, int A>
__device__ __forceinline__ void td_store(Real *p, const Real (&v)[A]) {
    if constexpr (sizeof(Real) * A >= 16) {
        constexpr int NV = (int)(sizeof(Real) * A / 16);
        uint4 raw[NV];
        memcpy(raw, v, sizeof(raw));
        uint4 *q = reinterpret_cast<uint4 *>(p);
#pragma unroll
        for (int i = 0; i < NV; ++i) q[i] = raw[i];
    } else {
#pragma unroll
        for (int a = 0; a < A; ++a) p[a] = v[a];
    }
}

// bytes of one ring stage: 2048 values of either precision
template <typename Real>
__host__ __device__ constexpr size_t td_stage_bytes() {
    return TD_WBUF_BYTES / 8 * sizeof(Real);
}

template <typename Real>
__device__ __forceinline__ int td_chunk_rows(int no) {
    const int r = (int)(td_stage_bytes<Real>() / ((size_t)no * sizeof(Real))) & ~3;
    return r < 4 ? 4 : r;
}

// Weight stream of a block.  The chunks of all operations form ONE sequence; chunk c lives in
// stage c % TD_STAGES and completes phase (c / TD_STAGES) & 1 of that stage's mbarrier.  The
// producer (thread 0) runs TD_STAGES - 1 chunks ahead of the consumers: a transfer takes far
// longer than a chunk's arithmetic, so several must be in flight.
template <typename Real>
struct TdStream {
    const Real *blob;
    unsigned char *ring;
    uint64_t *bar;          // [TD_STAGES]
    const TdOp *ops;
    int n_ops;
    int ci;                 // next chunk to consume
    int pi, p_op, p_ch;     // producer: next chunk to issue = chunk p_ch of operation p_op
    bool ok;
    __device__ __forceinline__ Real *stage(int c) const {
        return reinterpret_cast<Real *>(ring + (size_t)(c % TD_STAGES) * (td_stage_bytes<Real>() + 128));
    }
};

// thread 0: start the transfer of the next chunk of the sequence (if any is left)
template <typename Real>
__device__ __forceinline__ void td_produce(TdStream<Real> &s) {
    if (s.p_op >= s.n_ops) return;
    const TdOp &op = s.ops[s.p_op];
    const int kc = td_chunk_rows<Real>(op.no);
    const int k0 = s.p_ch * kc;
    const int rows = min(kc, op.ni - k0);
    const uint32_t bytes = ((uint32_t)((size_t)rows * op.no * sizeof(Real)) + 15u) & ~15u;
    td_bulk_load(s.stage(s.pi), s.blob + op.w_off + (size_t)k0 * op.no, bytes,
                 &s.bar[s.pi % TD_STAGES]);
    ++s.pi;
    if (k0 + rows >= op.ni) {
        ++s.p_op;
        s.p_ch = 0;
    } else {
        ++s.p_ch;
    }
}

// one operation of the schedule.  Epilogue per owned column.
// The block's 128 NG threads form NG groups; group g owns the atoms [g AT, (g + 1) AT) of the
// tile (AT = A / NG) for every column, so all warps run the same loop on disjoint accumulators and
// the epilogue (activation included) is spread over all of them without any reduction.
template <typename Real, int A, int NG>
__device__ __forceinline__ void td_gemm(TdStream<Real> &st, const TdOp &op, Real *pool_tile) {
    constexpr int RS = TdRow<Real, A>::RS;
    constexpr int AT = A / NG;
    const int tid = threadIdx.x % TD_COLS;
    Real *pool = pool_tile + (threadIdx.x / TD_COLS) * AT;       // this group's atoms
    const int ni = op.ni, no = op.no;
    const int kc = td_chunk_rows<Real>(no);
    const int nch = (ni + kc - 1) / kc;
    const Real *vin = pool + (size_t)op.in_off * RS;
    const bool two = no > TD_COLS;
    const int o0 = tid, o1 = tid + TD_COLS;
    const bool own0 = o0 < no, own1 = two && o1 < no;
    Real acc0[AT], acc1[AT];
#pragma unroll
    for (int a = 0; a < AT; ++a) acc0[a] = acc1[a] = Real(0);
    for (int ch = 0; ch < nch; ++ch) {
        // the stage of chunk ci - 1 was released by the __syncthreads that ended it: refill it
        if (threadIdx.x == 0) td_produce<Real>(st);
        st.ok = td_mbar_wait(&st.bar[st.ci % TD_STAGES],
                             (uint32_t)((st.ci / TD_STAGES) & 1)) && st.ok;
        const Real *W = st.stage(st.ci);
        const int k0 = ch * kc;
        const int rows = min(kc, ni - k0);
        if (own0) {
            const Real *vk = vin + (size_t)k0 * RS;
            if (!two) {
#pragma unroll 4
                for (int kk = 0; kk < rows; ++kk) {
                    const Real w = W[(size_t)kk * no + o0];
                    Real h[AT];
                    td_load<Real, AT>(vk + (size_t)kk * RS, h);
#pragma unroll
                    for (int a = 0; a < AT; ++a) acc0[a] = fma(h[a], w, acc0[a]);
                }
            } else {
#pragma unroll 2
                for (int kk = 0; kk < rows; ++kk) {
                    const Real w0 = W[(size_t)kk * no + o0];
                    const Real w1 = own1 ? W[(size_t)kk * no + o1] : Real(0);
                    Real h[AT];
                    td_load<Real, AT>(vk + (size_t)kk * RS, h);
#pragma unroll
                    for (int a = 0; a < AT; ++a) {
                        acc0[a] = fma(h[a], w0, acc0[a]);
                        acc1[a] = fma(h[a], w1, acc1[a]);
                    }
                }
            }
        }
        ++st.ci;
        __syncthreads();
    }
    // epilogue (the inputs are no longer read by anybody: the loop ended with a barrier)
    for (int c = 0; c < 2; ++c) {
        const int o = c == 0 ? o0 : o1;
        if (!(c == 0 ? own0 : own1)) continue;
        Real v[AT];
#pragma unroll
        for (int a = 0; a < AT; ++a) v[a] = c == 0 ? acc0[a] : acc1[a];
        if (op.kind == TD_BWD) {
            if (op.res_off >= 0) {
                Real r[AT];
                td_load<Real, AT>(pool + ((size_t)op.res_off + o) * RS, r);
#pragma unroll
                for (int a = 0; a < AT; ++a) v[a] += r[a];
            }
            if (op.out_off >= 0) td_store<Real, AT>(pool + ((size_t)op.out_off + o) * RS, v);
            if (op.t_off >= 0) {
                Real dzv[AT];
                td_load<Real, AT>(pool + ((size_t)op.dz_off + o) * RS, dzv);
#pragma unroll
                for (int a = 0; a < AT; ++a) v[a] *= dzv[a];
                td_store<Real, AT>(pool + ((size_t)op.t_off + o) * RS, v);
            }
        } else {
            const Real b = op.has_bias ? st.blob[op.b_off + o] : Real(0);
            if (op.kind == TD_FWD_LAST) {
#pragma unroll
                for (int a = 0; a < AT; ++a) v[a] += b;
            } else {
                Real dzv[AT];
#pragma unroll
                for (int a = 0; a < AT; ++a) v[a] = td_act<Real>(op.act, v[a] + b, dzv[a]);
                td_store<Real, AT>(pool + ((size_t)op.dz_off + o) * RS, dzv);
                if (op.res_off >= 0) {
                    Real r[AT];
                    td_load<Real, AT>(pool + ((size_t)op.res_off + o) * RS, r);
#pragma unroll
                    for (int a = 0; a < AT; ++a) v[a] += r[a];
                }
            }
            td_store<Real, AT>(pool + ((size_t)op.out_off + o) * RS, v);
        }
    }
    __syncthreads();
}

template <typename Real, int A, int NG>
__global__ void __launch_bounds__(TD_COLS * NG)
k_td_heads(int n, TdPlan P, const TdElem *__restrict__ elems, const Real *__restrict__ blob,
           const int *__restrict__ list, const int *__restrict__ cnt,
           const double *__restrict__ G, const double *__restrict__ T,
           double *__restrict__ U, double *__restrict__ S, double *__restrict__ F,
           double *__restrict__ dFdG, int *__restrict__ status) {
    extern __shared__ __align__(128) unsigned char td_smem[];
    __shared__ __align__(8) uint64_t bar[TD_STAGES];
    __shared__ int ids[A];
    __shared__ Real temp[A];
    constexpr int RS = TdRow<Real, A>::RS;
    constexpr int NT = TD_COLS * NG;
    const int tid = threadIdx.x;
    const int e = blockIdx.y;
    const int t0 = blockIdx.x * A;
    const int m = cnt[e];
    if (t0 >= m) return;
    // the element's schedule: header in registers, operations in shared memory (the producer
    // reads an operation per chunk; from global memory that is an L2 round trip on the critical
    // path of every chunk, measured as 30 % of the kernel in barrier waits)
    __shared__ TdOp s_ops[TD_MAX_OPS];
    __shared__ int s_head[16];
    {
        const TdElem &Eg = elems[e];
        const int n_ops = Eg.n_ops;
        const int *src = reinterpret_cast<const int *>(Eg.ops);
        int *dst = reinterpret_cast<int *>(s_ops);
        for (int q = tid; q < n_ops * (int)(sizeof(TdOp) / sizeof(int)); q += NT) dst[q] = src[q];
        if (tid == 0) {
            s_head[0] = Eg.n_ops;
            s_head[1] = Eg.n_fwd;
            s_head[2] = Eg.has_minmax;
            s_head[3] = Eg.x_off;
            s_head[4] = Eg.ht_off;
            s_head[5] = Eg.s_off;
            s_head[6] = Eg.u_off;
            s_head[7] = Eg.seed_s_off;
            s_head[8] = Eg.seed_u_off;
            s_head[9] = Eg.dx_off;
        }
    }
    const long long xlo_off = elems[e].xlo_off, xhi_off = elems[e].xhi_off;
    __syncthreads();
    struct {
        int n_ops, n_fwd, has_minmax, x_off, ht_off, s_off, u_off, seed_s_off, seed_u_off, dx_off;
        long long xlo_off, xhi_off;
        const TdOp *ops;
    } E;
    E.n_ops = s_head[0];
    E.n_fwd = s_head[1];
    E.has_minmax = s_head[2];
    E.x_off = s_head[3];
    E.ht_off = s_head[4];
    E.s_off = s_head[5];
    E.u_off = s_head[6];
    E.seed_s_off = s_head[7];
    E.seed_u_off = s_head[8];
    E.dx_off = s_head[9];
    E.xlo_off = xlo_off;
    E.xhi_off = xhi_off;
    E.ops = s_ops;
    // shared memory: the ring of weight stages, then the pool
    TdStream<Real> st;
    st.blob = blob;
    st.ring = td_smem;
    st.bar = bar;
    st.ops = s_ops;
    st.n_ops = E.n_ops;
    st.ci = st.pi = st.p_op = st.p_ch = 0;
    st.ok = true;
    Real *pool = reinterpret_cast<Real *>(td_smem + TD_STAGES * (td_stage_bytes<Real>() + 128));
    if (tid == 0) {
        for (int k = 0; k < TD_STAGES; ++k) td_mbar_init(&bar[k], 1);
        td_fence_barrier_init();
    }
    if (tid < A) {
        const int id = t0 + tid < m ? list[(size_t)e * n + t0 + tid] : -1;
        ids[tid] = id;
        temp[tid] = id >= 0 ? (Real)T[id] : Real(0);
    }
    __syncthreads();
    if (tid == 0)       // the first transfers land while the inputs are read
        for (int k = 0; k < TD_STAGES - 1; ++k) td_produce<Real>(st);
    const int dim = P.dim, nH = P.nH;
    {
        // inputs (atoms beyond the bucket: zeros, never stored)
        Real *x = pool + (size_t)E.x_off * RS;
        for (int q = tid; q < dim * A; q += NT) {
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
            x[(size_t)k * RS + a] = v;
        }
        // Ht = [H, T]: the temperature row below the output of H
        if (tid < A) pool[((size_t)E.ht_off + nH) * RS + tid] = temp[tid];
    }
    __syncthreads();
    for (int i = 0; i < E.n_fwd; ++i)
        td_gemm<Real, A, NG>(st, E.ops[i], pool);
    if (tid < A) {
        const Real t = temp[tid];
        const Real s = pool[(size_t)E.s_off * RS + tid], u = pool[(size_t)E.u_off * RS + tid];
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
        pool[(size_t)E.seed_s_off * RS + tid] = -t * dSds;
        pool[(size_t)E.seed_u_off * RS + tid] = Real(1);
    }
    __syncthreads();
    for (int i = E.n_fwd; i < E.n_ops; ++i)
        td_gemm<Real, A, NG>(st, E.ops[i], pool);
    const Real *gx = pool + (size_t)E.dx_off * RS;
    const bool ok = st.ok;
    for (int q = tid; q < dim * A; q += NT) {
        const int k = q / A, a = q % A;
        const int id = ids[a];
        if (id < 0) continue;
        Real v = gx[(size_t)k * RS + a];
        if (E.has_minmax) {
            const Real den = blob[E.xhi_off + k] - blob[E.xlo_off + k];
            v = den != Real(0) ? -v / den : Real(0);
        }
        dFdG[(size_t)id * dim + k] = ok ? (double)v : __longlong_as_double(0x7ff8000000000000LL);
    }
    if (!ok && tid == 0) atomicExch(status, 1);
}

// -- host side: the schedule -------------------------------------------------------------
namespace {

struct Net {            // one network as the caller described it
    const tab_mlp_desc *q;
    int L;
    int in(int l) const { return q->sizes[l]; }
    int out(int l) const { return q->sizes[l + 1]; }
};

struct Builder {
    std::vector<double> blob;
    // pool rows of one element: persistent ones (act' of every layer, [H, T] / dF/dH, seeds,
    // dF/dx) from 0, transient ones (layer outputs of the forward pass; the reverse pass reuses
    // the region from its start) from t_base.  t_base comes from a dry run of the same build.
    int rows_p = 0, rows_t = 0, t_base = 0, t_max = 0;
    int alloc_p(int n) {
        const int at = rows_p;
        rows_p += n;
        return at;
    }
    int alloc_t(int n) {
        const int at = t_base + rows_t;
        rows_t += n;
        t_max = rows_t > t_max ? rows_t : t_max;
        return at;
    }
    void reset_t() { rows_t = 0; }
    void reset(int base) {
        rows_p = rows_t = t_max = 0;
        t_base = base;
    }
    size_t pad() {      // 16-byte alignment for float32 and float64 copies of the blob
        while (blob.size() % 4) blob.push_back(0.0);
        return blob.size();
    }
    // [ni][no] row-major, optionally the columns of a second matrix appended
    long long put_matrix(const double *w, int ni, int no, const double *w2 = nullptr,
                         int no2 = 0) {
        const size_t at = pad();
        for (int k = 0; k < ni; ++k) {
            blob.insert(blob.end(), w + (size_t)k * no, w + (size_t)(k + 1) * no);
            if (w2) blob.insert(blob.end(), w2 + (size_t)k * no2, w2 + (size_t)(k + 1) * no2);
        }
        return (long long)at;
    }
    // transpose of the first `ni_used` rows: [no][ni_used]; optionally a second matrix's
    // transpose stacked below (rows of the concatenated columns)
    long long put_transpose(const double *w, int ni_used, int no, const double *w2 = nullptr,
                            int no2 = 0) {
        const size_t at = pad();
        for (int o = 0; o < no; ++o)
            for (int k = 0; k < ni_used; ++k) blob.push_back(w[(size_t)k * no + o]);
        if (w2)
            for (int o = 0; o < no2; ++o)
                for (int k = 0; k < ni_used; ++k) blob.push_back(w2[(size_t)k * no2 + o]);
        return (long long)at;
    }
    long long put_vector(const double *b, int n, const double *b2 = nullptr, int n2 = 0) {
        const size_t at = pad();
        for (int k = 0; k < n; ++k) blob.push_back(b ? b[k] : 0.0);
        for (int k = 0; k < n2; ++k) blob.push_back(b2 ? b2[k] : 0.0);
        return (long long)at;
    }
};

TdOp make_op(int kind, long long w, long long b, int ni, int no, int in_off, int out_off) {
    TdOp op;
    memset(&op, 0, sizeof(op));
    op.kind = kind;
    op.w_off = w;
    op.b_off = b;
    op.ni = ni;
    op.no = no;
    op.in_off = in_off;
    op.out_off = out_off;
    op.dz_off = op.res_off = op.t_off = -1;
    return op;
}

int check_net(const tab_mlp_desc &q, const char *name, int in_expect, int out_expect) {
    if (q.n_layers < 1 || q.n_layers > TD_MAX_LAYERS) {
        tab_set_error("tab_td_create: %s: 1..%d layers", name, TD_MAX_LAYERS);
        return TAB_EINVAL;
    }
    if (q.sizes[0] != in_expect || q.sizes[q.n_layers] != out_expect) {
        tab_set_error("tab_td_create: %s maps %d -> %d values, expected %d -> %d", name,
                      q.sizes[0], q.sizes[q.n_layers], in_expect, out_expect);
        return TAB_EINVAL;
    }
    for (int l = 0; l < q.n_layers; ++l) {
        if (q.sizes[l] < 1 || q.sizes[l + 1] < 1 || q.sizes[l + 1] > TD_MAX_WIDTH ||
            q.sizes[l] > 4 * TD_MAX_WIDTH || !q.weights[l]) {
            tab_set_error("tab_td_create: %s layer %d: %d -> %d (at most %d outputs) and "
                          "weights required", name, l, q.sizes[l], q.sizes[l + 1], TD_MAX_WIDTH);
            return TAB_EINVAL;
        }
    }
    return TAB_OK;
}

// forward ops of layers [l0, L) of one network reading `in_off`; records where every layer's
// act' lives (dz[l]) and returns the pool offset of the output
int forward_chain(Builder &B, std::vector<TdOp> &ops, const Net &N, int l0, int in_off,
                  int *dz, int extra_out_rows) {
    for (int l = l0; l < N.L; ++l) {
        const int ni = N.in(l), no = N.out(l);
        const bool last = l == N.L - 1;
        const long long w = B.put_matrix(N.q->weights[l], ni, no);
        const long long b = B.put_vector(N.q->biases[l], no);
        // (the output of H -- with the temperature row -- doubles as dF/dH later: persistent)
        const int out = extra_out_rows && last ? B.alloc_p(no + extra_out_rows) : B.alloc_t(no);
        TdOp op = make_op(last ? TD_FWD_LAST : TD_FWD_HIDDEN, w, b, ni, no, in_off, out);
        op.act = N.q->activation;
        op.has_bias = last ? (N.q->output_bias ? 1 : 0) : 1;
        if (!last) {
            op.dz_off = dz[l] = B.alloc_p(no);
            if (l > 0 && N.q->use_resnet_dt && no == ni) op.res_off = in_off;
        }
        ops.push_back(op);
        in_off = out;
    }
    return in_off;
}

// reverse ops of layers (l_stop, L) of one network, from the seed down to the derivative with
// respect to the OUTPUT of layer l_stop (l_stop >= 0) -- multiplied by that layer's act' into
// `t_final` -- or, with l_stop = -1, down to the network's input (`out_final`).
void reverse_chain(Builder &B, std::vector<TdOp> &ops, const Net &N, int l_stop, int seed_off,
                   const int *dz, int t_final, int dz_final, int out_final, int ni_used_first) {
    int in_off = seed_off, prev_raw = -1;
    for (int l = N.L - 1; l > l_stop; --l) {
        const int ni = (l == 0 && ni_used_first > 0) ? ni_used_first : N.in(l);
        const int no = N.out(l);
        const bool last = l == N.L - 1;
        const long long wt = B.put_transpose(N.q->weights[l], ni, no);
        const bool to_input = l == 0;
        // the raw derivative is stored only where somebody reads it: the final output, or the
        // resnet link of the layer below
        const bool below_res = l >= 2 && N.q->use_resnet_dt && N.out(l - 1) == N.in(l - 1);
        const int out = to_input ? (out_final >= 0 ? out_final : B.alloc_t(ni))
                                 : (below_res ? B.alloc_t(ni) : -1);
        TdOp op = make_op(TD_BWD, wt, 0, no, ni, in_off, out);
        if (!last && l > 0 && N.q->use_resnet_dt && no == N.in(l)) op.res_off = prev_raw;
        if (l > 0) {
            // fold act' of the layer below into the input of the next reverse op
            const bool hand_over = l - 1 == l_stop;
            op.dz_off = hand_over ? dz_final : dz[l - 1];
            op.t_off = hand_over ? t_final : B.alloc_t(ni);
        }
        ops.push_back(op);
        prev_raw = out;
        in_off = op.t_off;
    }
}

}  // namespace

extern "C" int tab_td_create(tab_td **out, const tab_td_desc *d) {
    if (!out || !d || !d->H || !d->S || !d->U || d->n_elements < 1 ||
        d->n_elements > TAB_MAX_ELEMENTS || d->dim < 1 || d->dim > 4 * TD_MAX_WIDTH) {
        tab_set_error("tab_td_create: invalid descriptor");
        return TAB_EINVAL;
    }
    const int dim = d->dim;
    const int nH = d->H[0].sizes[d->H[0].n_layers];
    tab_td *m = new tab_td();
    m->n_el = d->n_elements;
    m->elem.resize(d->n_elements);
    Builder B;
    int pool_rows = 0;
    // the schedule of one element (run twice: a dry run on a copy measures the persistent
    // rows, which fixes where the transient region starts)
    auto build = [&](int e, Builder &B, TdElem &E, std::vector<TdOp> &ops) -> int {
        const Net H{&d->H[e], d->H[e].n_layers}, S{&d->S[e], d->S[e].n_layers},
            U{&d->U[e], d->U[e].n_layers};
        memset(&E, 0, sizeof(E));
        ops.clear();
        int dzH[TD_MAX_LAYERS], dzS[TD_MAX_LAYERS], dzU[TD_MAX_LAYERS];
        E.x_off = B.alloc_t(dim);
        E.ht_off = forward_chain(B, ops, H, 0, E.x_off, dzH, 1);      // + the temperature row
        // S and U: same input; with equal first-layer activation the two first layers run as
        // one operation on [W_S | W_U]
        const bool merged = S.L >= 2 && U.L >= 2 && S.q->activation == U.q->activation;
        int hS = 0, hU = 0, su_dz = -1, su_t = -1;
        if (merged) {
            hS = S.out(0);
            hU = U.out(0);
            if (hS + hU > TD_MAX_WIDTH) {
                tab_set_error("tab_td_create: first layers of S and U: %d + %d > %d columns", hS,
                              hU, TD_MAX_WIDTH);
                return TAB_EUNSUPPORTED;
            }
            const long long w = B.put_matrix(S.q->weights[0], nH + 1, hS, U.q->weights[0], hU);
            const long long b = B.put_vector(S.q->biases[0], hS, U.q->biases[0], hU);
            const int su_out = B.alloc_t(hS + hU);
            su_dz = B.alloc_p(hS + hU);
            su_t = B.alloc_p(hS + hU);
            TdOp op = make_op(TD_FWD_HIDDEN, w, b, nH + 1, hS + hU, E.ht_off, su_out);
            op.act = S.q->activation;
            op.has_bias = 1;
            op.dz_off = su_dz;
            ops.push_back(op);
            dzS[0] = su_dz;
            dzU[0] = su_dz + hS;
            E.s_off = forward_chain(B, ops, S, 1, su_out, dzS, 0);
            E.u_off = forward_chain(B, ops, U, 1, su_out + hS, dzU, 0);
        } else {
            E.s_off = forward_chain(B, ops, S, 0, E.ht_off, dzS, 0);
            E.u_off = forward_chain(B, ops, U, 0, E.ht_off, dzU, 0);
        }
        E.n_fwd = (int)ops.size();
        E.seed_s_off = B.alloc_p(1);
        E.seed_u_off = B.alloc_p(1);
        E.dx_off = B.alloc_p(dim);
        // dF/dH (sum of the S and U parts) overwrites H: nobody reads [H, T] after the forward
        // pass.  The reverse pass reuses the transient region from its start (s and u were
        // consumed by the entropy model before the first reverse operation runs).
        const int dht = E.ht_off;
        B.reset_t();
        if (merged) {
            reverse_chain(B, ops, S, 0, E.seed_s_off, dzS, su_t, su_dz, -1, 0);
            reverse_chain(B, ops, U, 0, E.seed_u_off, dzU, su_t + hS, su_dz + hS, -1, 0);
            // [t_S ; t_U] -> dF/dH through the transposed first layers (temperature column
            // skipped)
            const long long wt = B.put_transpose(S.q->weights[0], nH, hS, U.q->weights[0], hU);
            ops.push_back(make_op(TD_BWD, wt, 0, hS + hU, nH, su_t, dht));
        } else {
            reverse_chain(B, ops, S, -1, E.seed_s_off, dzS, -1, -1, dht, nH);
            reverse_chain(B, ops, U, -1, E.seed_u_off, dzU, -1, -1, dht, nH);
            ops.back().res_off = dht;       // U part + S part, in place
        }
        reverse_chain(B, ops, H, -1, dht, dzH, -1, -1, E.dx_off, 0);
        if ((int)ops.size() > TD_MAX_OPS) {
            tab_set_error("tab_td_create: %zu operations (at most %d)", ops.size(), TD_MAX_OPS);
            return TAB_EUNSUPPORTED;
        }
        E.n_ops = (int)ops.size();
        memcpy(E.ops, ops.data(), sizeof(TdOp) * ops.size());
        E.has_minmax = (d->H[e].xlo && d->H[e].xhi) ? 1 : 0;
        E.xlo_off = B.put_vector(E.has_minmax ? d->H[e].xlo : nullptr, dim);
        E.xhi_off = B.put_vector(E.has_minmax ? d->H[e].xhi : nullptr, dim);
        return TAB_OK;
    };
    for (int e = 0; e < d->n_elements; ++e) {
        int rc = check_net(d->H[e], "H", dim, nH);
        if (rc == TAB_OK) rc = check_net(d->S[e], "S", nH + 1, 1);
        if (rc == TAB_OK) rc = check_net(d->U[e], "U", nH + 1, 1);
        std::vector<TdOp> ops;
        if (rc == TAB_OK) {
            Builder dry = B;
            TdElem scratch;
            dry.reset(0);
            rc = build(e, dry, scratch, ops);
            if (rc == TAB_OK) {
                B.reset(dry.rows_p);
                rc = build(e, B, m->elem[e], ops);
            }
        }
        if (rc != TAB_OK) {
            delete m;
            return rc;
        }
        const int rows = B.t_base + B.t_max;
        pool_rows = rows > pool_rows ? rows : pool_rows;
    }
    for (int k = 0; k < 8; ++k) B.blob.push_back(0.0);      // the last chunk may be rounded up
    TdPlan &P = m->plan;
    P.dim = dim;
    P.nH = nH;
    P.algo = d->algo;
    P.special = d->special;
    P.pool_rows = pool_rows;
    // atoms per block: the largest of 16, 8, 4, 2 whose pool fits beside the weight buffers
    for (int p = 0; p < 2; ++p) {
        const size_t w = p == 0 ? 8 : 4;
        const size_t fixed = TD_STAGES * ((p == 0 ? td_stage_bytes<double>()
                                                  : td_stage_bytes<float>()) + 128);
        int A = 16;
        if (const char *env = getenv("TAB_TD_TILE")) {      // A/B measurements
            const int v = atoi(env);
            if (v == 2 || v == 4 || v == 8 || v == 16) A = v;
        }
        while (A > 2 && fixed + (size_t)pool_rows * (A * w + 16) > TD_SMEM_MAX) A >>= 1;
        if (fixed + (size_t)pool_rows * (A * w + 16) > TD_SMEM_MAX) {
            tab_set_error("tab_td_create: networks need %zu bytes of shared memory per atom",
                          (size_t)pool_rows * w);
            delete m;
            return TAB_EUNSUPPORTED;
        }
        m->tile[p] = A;
    }
    const size_t total = B.blob.size();
    std::vector<float> blobf(total);
    for (size_t k = 0; k < total; ++k) blobf[k] = (float)B.blob[k];
    int rc = m->blob.ensure(total * 8);
    if (rc == TAB_OK) rc = m->blobf.ensure(total * 4);
    if (rc == TAB_OK) rc = m->elems_dev.ensure(sizeof(TdElem) * m->n_el);
    if (rc == TAB_OK) rc = m->cnt.ensure(sizeof(int) * TAB_MAX_ELEMENTS);
    if (rc == TAB_OK) rc = m->status.ensure(sizeof(int));
    if (rc == TAB_OK) {
        cudaError_t err = cudaMemcpy(m->blob.p, B.blob.data(), total * 8, cudaMemcpyHostToDevice);
        if (err == cudaSuccess)
            err = cudaMemcpy(m->blobf.p, blobf.data(), total * 4, cudaMemcpyHostToDevice);
        if (err == cudaSuccess)
            err = cudaMemcpy(m->elems_dev.p, m->elem.data(), sizeof(TdElem) * m->n_el,
                             cudaMemcpyHostToDevice);
        if (err == cudaSuccess) err = cudaMemset(m->status.p, 0, sizeof(int));
        if (err != cudaSuccess) {
            tab_set_error("tab_td_create: cudaMemcpy -> %s", cudaGetErrorString(err));
            rc = TAB_ECUDA;
        }
    }
    if (rc != TAB_OK) {
        delete m;
        return rc;
    }
    *out = m;
    return TAB_OK;
}

extern "C" int tab_td_free(tab_td *m) {
    if (!m) return TAB_OK;
    DevBuf *bufs[] = {&m->blob, &m->blobf, &m->elems_dev, &m->list, &m->cnt, &m->status};
    for (DevBuf *b : bufs) b->release();
    delete m;
    return TAB_OK;
}

template <typename Real, int A, int NG>
static int td_launch(tab_td *m, int n, const Real *blob, const double *G, const double *T,
                     double *U, double *S, double *F, double *dFdG, cudaStream_t st) {
    const size_t smem = TD_STAGES * (td_stage_bytes<Real>() + 128) +
                        (size_t)m->plan.pool_rows * TdRow<Real, A>::RS * sizeof(Real);
    TAB_CUDA(cudaFuncSetAttribute(k_td_heads<Real, A, NG>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)((n + A - 1) / A), (unsigned)m->n_el);
    k_td_heads<Real, A, NG><<<grid, TD_COLS * NG, smem, st>>>(
        n, m->plan, m->elems_dev.as<TdElem>(), blob, m->list.as<int>(), m->cnt.as<int>(), G, T,
        U, S, F, dFdG, m->status.as<int>());
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
    // atom groups per block: as many (4, 2, 1) as leave a group 16 bytes of every pool row
    const int min_at = f64 ? 2 : 4;
    int G = A / 4 >= min_at ? 4 : (A / 2 >= min_at ? 2 : 1);
    if (!f64 && G == 4) G = 2;      // measured: float32 0.91 ms with 2 groups, 1.02 ms with 4
    if (const char *env = getenv("TAB_TD_SPLIT")) {      // A/B measurements
        const int v = atoi(env);
        if ((v == 1 || v == 2 || v == 4) && A / v >= (f64 ? 2 : 4)) G = v;
    }
#define TD_GO(R, AA, GG, B)                                                                    \
    if (A == AA && G == GG)                                                                    \
        return td_launch<R, AA, GG>(m, n, B, d_G, d_T, d_U, d_S, d_F, d_dFdG, st);
    if (f64) {
        TD_GO(double, 16, 4, m->blob.as<double>())
        TD_GO(double, 16, 2, m->blob.as<double>())
        TD_GO(double, 16, 1, m->blob.as<double>())
        TD_GO(double, 8, 4, m->blob.as<double>())
        TD_GO(double, 8, 2, m->blob.as<double>())
        TD_GO(double, 8, 1, m->blob.as<double>())
        TD_GO(double, 4, 2, m->blob.as<double>())
        TD_GO(double, 4, 1, m->blob.as<double>())
        TD_GO(double, 2, 1, m->blob.as<double>())
    } else {
        TD_GO(float, 16, 4, m->blobf.as<float>())
        TD_GO(float, 16, 2, m->blobf.as<float>())
        TD_GO(float, 16, 1, m->blobf.as<float>())
        TD_GO(float, 8, 2, m->blobf.as<float>())
        TD_GO(float, 8, 1, m->blobf.as<float>())
        TD_GO(float, 4, 1, m->blobf.as<float>())
        TD_GO(float, 2, 1, m->blobf.as<float>())
    }
#undef TD_GO
    return TAB_ESTATE;
}

/* 0 = every weight transfer of the evaluations so far completed; 1 = a transfer timed out (the
 * outputs of that call hold NaN).  Synchronises the stream. */
extern "C" int tab_td_status(tab_td *m, int32_t *out, void *stream) {
    if (!m || !out) return TAB_EINVAL;
    int h = 0;
    TAB_CUDA(cudaMemcpyAsync(&h, m->status.p, sizeof(int), cudaMemcpyDeviceToHost,
                             (cudaStream_t)stream));
    TAB_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    *out = h;
    return TAB_OK;
}
