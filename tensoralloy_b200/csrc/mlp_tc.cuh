// mlp_tc.cuh -- the per-element atomic MLP (nn/atomic/atomic.py:157-302,
// nn/convolutional.py:257-290) on the 5th-generation tensor cores: forward AND the
// backward pass to dE/dG as four chained GEMMs per tile of 128 atoms,
//
//   z1 = X  W1          [128 x D ] [D  x H1]      h1 = act(z1 + b1)
//   z2 = h1 W2          [128 x H1] [H1 x H2]      h2 = act(z2 + b2),  E = h2 . w3 (+ b3)
//   dh1 = (w3 act'(z2)) W2^T   [128 x H2] [H2 x H1]
//   dX  = (dh1 act'(z1)) W1^T  [128 x H1] [H1 x D ]
//
// issued as tcgen05.mma (cta_group::1, kind::tf32, M = 128) by one thread, operands in
// shared memory in the canonical K-major no-swizzle layout, accumulators in tensor
// memory (TMEM), read back with tcgen05.ld 32x32b (thread t = TMEM lane t = atom row t)
// for the bias / activation epilogues, which also write the next GEMM's A operand.
//
// Precision: 'medium' (float32) only -- the reference tolerance there is 1e-5 relative,
// which one TF32 product (10-bit mantissa) misses, so every GEMM is the 3-term split
//   A B ~= Ahi Bhi + Ahi Blo + Alo Bhi,   xhi = x with the low 13 mantissa bits cleared,
// accumulated in float32 in TMEM (relative error ~1e-6).  'high' (float64) stays on the
// warp-per-atom kernel k_mlp.
//
// Shape limits of this kernel (anything else falls back to k_mlp): two hidden layers,
// D <= 64, H1 <= 64, H2 <= 32 (padded to multiples of 16 with zeros), no ResNet link.
#pragma once
#include <stdint.h>

#define TC_ROWS 128          // atoms per tile = UMMA M = TMEM lanes
#define TC_KMAX 64           // widest A operand (columns)
#define TC_TMEM_COLS 256     // z1 [0,64) z2 [64,96) dh1 [96,160) dX [160,224)
#define TC_SPIN_LIMIT (1u << 26)

struct MlpTcDev {            // one element's network, padded sizes
    int dim, dp, h1, h1p, h2, h2p, act, has_out_bias, has_minmax;
    long long w1, b1, w2, b2, w3, b3, xlo, xhi;    // offsets into the double blob
};

// ---- PTX wrappers --------------------------------------------------------------------
__device__ __forceinline__ uint32_t tc_smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void tc_fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before_sync() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after_sync() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tc_smem_u32(bar)), "r"(count)
                 : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ bool tc_mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(tc_smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0u;
}
// bounded wait: a descriptor mistake must not hang the GPU
__device__ __forceinline__ bool tc_mbar_wait(uint64_t *bar, uint32_t parity) {
    for (uint32_t spin = 0; spin < TC_SPIN_LIMIT; ++spin)
        if (tc_mbar_try_wait(bar, parity)) return true;
    return false;
}
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                     tc_smem_u32(bar))
                 : "memory");
}
// D[tmem] (+)= A[smem] B[smem], one UMMA of K = 8 tf32
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                            uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}"
        :
        : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// 16 consecutive columns of this thread's TMEM lane
__device__ __forceinline__ void tc_ld16(uint32_t taddr, float *v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
          "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]),
          "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int q = 0; q < 16; ++q) v[q] = __uint_as_float(r[q]);
}

// ---- operand layout --------------------------------------------------------------------
// canonical K-major, no swizzle: 8 x 16 B core matrices (8 rows, 4 tf32 each), the 16
// (or R/8) row groups of one 4-column chunk contiguous, chunks `lbo` bytes apart.
__device__ __forceinline__ uint32_t tc_off(int row, int col, int rows) {
    return (uint32_t)((col >> 2) * (rows * 16) + (row >> 3) * 128 + (row & 7) * 16 + (col & 3) * 4);
}
__device__ __forceinline__ uint64_t tc_desc(uint32_t smem_addr, int rows) {
    const uint64_t lbo = (uint64_t)(rows * 16) >> 4, sbo = 128 >> 4;
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (lbo << 16) | (sbo << 32) | (1ull << 46);
}
__device__ __forceinline__ uint32_t tc_idesc(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) |
           ((uint32_t)(TC_ROWS >> 4) << 24);
}
__device__ __forceinline__ void tc_split(float x, float &hi, float &lo) {
    hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
    lo = x - hi;
}

// issue one GEMM: D[128 x n] = A[128 x k] B[n x k]^T, 3-term TF32 split (thread 0 only)
__device__ __forceinline__ void tc_gemm(uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo,
                                        uint32_t b_hi, uint32_t b_lo, int n, int k,
                                        uint64_t *bar) {
    const uint32_t idesc = tc_idesc(n);
    for (int ks = 0; ks < k; ks += 8) {
        // one UMMA covers two 4-column chunks
        const uint32_t a_adv = (uint32_t)(ks >> 2) * (TC_ROWS * 16);
        const uint32_t b_adv = (uint32_t)(ks >> 2) * (uint32_t)(n * 16);
        const uint64_t dah = tc_desc(a_hi + a_adv, TC_ROWS), dal = tc_desc(a_lo + a_adv, TC_ROWS);
        const uint64_t dbh = tc_desc(b_hi + b_adv, n), dbl = tc_desc(b_lo + b_adv, n);
        tc_mma_tf32(tmem_d, dah, dbh, idesc, ks > 0 ? 1u : 0u);
        tc_mma_tf32(tmem_d, dah, dbl, idesc, 1u);
        tc_mma_tf32(tmem_d, dal, dbh, idesc, 1u);
    }
    tc_commit(bar);
}

// activation value and derivative in float32 (same ids as act_fn in sf.cu)
__device__ __forceinline__ float tc_act(int kind, float z, float &d) {
    switch (kind) {
    case 0: {
        const float e = __expf(-fabsf(z));
        const float sp = fmaxf(z, 0.f) + log1pf(e);
        d = z >= 0.f ? 1.f / (1.f + e) : e / (1.f + e);
        return sp;
    }
    case 1: {
        const float t = tanhf(z);
        d = 1.f - t * t;
        return t;
    }
    case 2:
        d = z > 0.f ? 1.f : 0.f;
        return fmaxf(z, 0.f);
    case 3:
        d = z > 0.f ? 1.f : 0.2f;
        return z > 0.f ? z : 0.2f * z;
    case 4: {
        const float s = 1.f / (1.f + __expf(-z));
        d = s * (1.f - s);
        return s;
    }
    case 5: {
        const float q = 1.f + fabsf(z);
        d = 1.f / (q * q);
        return z / q;
    }
    case 6: {
        const float e = __expf(z);
        d = z > 0.f ? 1.f : e;
        return z > 0.f ? z : e - 1.f;
    }
    default: {
        const float s = sqrtf(z * z + 4.f);
        d = 0.5f * (1.f + z / s);
        return 0.5f * (z + s);
    }
    }
}

// shared-memory plan (bytes), all operand buffers 128-byte aligned
#define TC_A_BYTES (TC_ROWS * TC_KMAX * 4)        // 32 KB: one A operand
#define TC_B1_BYTES (64 * 64 * 4)                 // W1 as [H1p x Dp]
#define TC_B2_BYTES (32 * 64 * 4)                 // W2 as [H2p x H1p]
#define TC_B3_BYTES (64 * 32 * 4)                 // W2 as [H1p x H2p]
#define TC_B4_BYTES (64 * 64 * 4)                 // W1 as [Dp x H1p]
#define TC_SMEM_BYTES (2 * TC_A_BYTES + 2 * (TC_B1_BYTES + TC_B2_BYTES + TC_B3_BYTES + TC_B4_BYTES) + \
                       TC_ROWS * 64 * 4 + 1024)

// grid = (tile groups, elements); block = 128 threads; a block walks tiles
// blockIdx.x, blockIdx.x + gridDim.x, ... of 128 consecutive (sorted) atoms and
// evaluates the rows whose element is blockIdx.y (other rows are zero and ignored).
__global__ void __launch_bounds__(TC_ROWS)
k_mlp_tc(int n, int dim, const uint8_t *__restrict__ types_ext,
         const MlpTcDev *__restrict__ nets, const double *__restrict__ blob,
         const double *__restrict__ G, double *__restrict__ eat, double *__restrict__ dEdG,
         int *__restrict__ status) {
    extern __shared__ __align__(128) unsigned char tc_smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    __shared__ float b1s[64], b2s[32], w3s[32];

    const int t = threadIdx.x, warp = t >> 5;
    const int el = blockIdx.y;
    const MlpTcDev net = nets[el];
    const int dp = net.dp, h1p = net.h1p, h2p = net.h2p;

    unsigned char *p = tc_smem;
    float *a_hi = reinterpret_cast<float *>(p);               p += TC_A_BYTES;
    float *a_lo = reinterpret_cast<float *>(p);               p += TC_A_BYTES;
    float *b1_hi = reinterpret_cast<float *>(p);              p += TC_B1_BYTES;
    float *b1_lo = reinterpret_cast<float *>(p);              p += TC_B1_BYTES;
    float *b2_hi = reinterpret_cast<float *>(p);              p += TC_B2_BYTES;
    float *b2_lo = reinterpret_cast<float *>(p);              p += TC_B2_BYTES;
    float *b3_hi = reinterpret_cast<float *>(p);              p += TC_B3_BYTES;
    float *b3_lo = reinterpret_cast<float *>(p);              p += TC_B3_BYTES;
    float *b4_hi = reinterpret_cast<float *>(p);              p += TC_B4_BYTES;
    float *b4_lo = reinterpret_cast<float *>(p);              p += TC_B4_BYTES;
    float *da1 = reinterpret_cast<float *>(p);                // [h1p][128] act'(z1)

    // ---- one-time setup: barrier, TMEM, weights ----------------------------------
    if (t == 0) tc_mbar_init(&bar, 1);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         tc_smem_u32(&tmem_base_s)),
                     "r"((uint32_t)TC_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // weights -> B operands (zero padded), float32 split
    auto put = [&](float *hi, float *lo, int row, int col, int rows, double w) {
        float h, l;
        tc_split((float)w, h, l);
        const uint32_t o = tc_off(row, col, rows) >> 2;
        hi[o] = h;
        lo[o] = l;
    };
    for (int q = t; q < h1p * dp; q += TC_ROWS) {            // W1[k * h1 + j]
        const int j = q / dp, k = q - j * dp;
        const double w = (j < net.h1 && k < net.dim) ? blob[net.w1 + (long long)k * net.h1 + j] : 0.0;
        put(b1_hi, b1_lo, j, k, h1p, w);                      // B1[n = j][k]
        put(b4_hi, b4_lo, k, j, dp, w);                       // B4[n = k][j]
    }
    for (int q = t; q < h2p * h1p; q += TC_ROWS) {           // W2[j * h2 + o]
        const int o = q / h1p, j = q - o * h1p;
        const double w = (o < net.h2 && j < net.h1) ? blob[net.w2 + (long long)j * net.h2 + o] : 0.0;
        put(b2_hi, b2_lo, o, j, h2p, w);                      // B2[n = o][j]
        put(b3_hi, b3_lo, j, o, h1p, w);                      // B3[n = j][o]
    }
    if (t < 64) b1s[t] = t < net.h1 ? (float)blob[net.b1 + t] : 0.f;
    if (t < 32) {
        b2s[t] = t < net.h2 ? (float)blob[net.b2 + t] : 0.f;
        w3s[t] = t < net.h2 ? (float)blob[net.w3 + t] : 0.f;
    }
    const float b3 = net.has_out_bias ? (float)blob[net.b3] : 0.f;
    tc_fence_proxy_async();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = tmem_base_s;
    const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);   // this warp's 32 lanes
    const uint32_t sa_hi = tc_smem_u32(a_hi), sa_lo = tc_smem_u32(a_lo);
    uint32_t phase = 0;
    bool ok = true;

    const int n_tiles = (n + TC_ROWS - 1) / TC_ROWS;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int idx = tile * TC_ROWS + t;
        const bool mine = idx < n && (int)types_ext[idx] == el;
        // any row of this element in the tile?  (uniform decision for the block)
        const int any = __syncthreads_or(mine ? 1 : 0);
        if (!any) continue;

        // ---- A1 = X (min-max normalised), zero rows for other atoms -----------------
        for (int k = 0; k < dp; ++k) {
            float x = 0.f;
            if (mine && k < dim) {
                double g = G[(size_t)idx * dim + k];
                if (net.has_minmax) {
                    const double lo = blob[net.xlo + k], hi = blob[net.xhi + k];
                    const double den = hi - lo;
                    g = den != 0.0 ? (hi - g) / den : 0.0;
                }
                x = (float)g;
            }
            float h, l;
            tc_split(x, h, l);
            const uint32_t o = tc_off(t, k, TC_ROWS) >> 2;
            a_hi[o] = h;
            a_lo[o] = l;
        }
        tc_fence_proxy_async();
        tc_fence_before_sync();
        __syncthreads();
        if (t == 0) {
            tc_fence_after_sync();
            tc_gemm(tmem + 0, sa_hi, sa_lo, tc_smem_u32(b1_hi), tc_smem_u32(b1_lo), h1p, dp, &bar);
        }
        ok = tc_mbar_wait(&bar, phase) && ok;
        phase ^= 1u;
        tc_fence_after_sync();

        // ---- epilogue 1: h1 = act(z1 + b1) -> A2, act' -> da1 --------------------------
        for (int c = 0; c < h1p; c += 16) {
            float v[16];
            tc_ld16(lane_addr + (uint32_t)c, v);
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                float d;
                const float hval = tc_act(net.act, v[q] + b1s[c + q], d);
                const bool live = c + q < net.h1;
                float h, l;
                tc_split(live ? hval : 0.f, h, l);
                const uint32_t o = tc_off(t, c + q, TC_ROWS) >> 2;
                a_hi[o] = h;
                a_lo[o] = l;
                da1[(c + q) * TC_ROWS + t] = live ? d : 0.f;
            }
        }
        tc_fence_proxy_async();
        tc_fence_before_sync();
        __syncthreads();
        if (t == 0) {
            tc_fence_after_sync();
            tc_gemm(tmem + 64, sa_hi, sa_lo, tc_smem_u32(b2_hi), tc_smem_u32(b2_lo), h2p, h1p, &bar);
        }
        ok = tc_mbar_wait(&bar, phase) && ok;
        phase ^= 1u;
        tc_fence_after_sync();

        // ---- epilogue 2: E = w3 . act(z2 + b2) (+ b3); A3 = w3 act'(z2) ----------------
        float e_atom = b3;
        for (int c = 0; c < h2p; c += 16) {
            float v[16];
            tc_ld16(lane_addr + 64u + (uint32_t)c, v);
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                float d;
                const float hval = tc_act(net.act, v[q] + b2s[c + q], d);
                const bool live = c + q < net.h2;
                e_atom += live ? w3s[c + q] * hval : 0.f;
                float h, l;
                tc_split(live ? w3s[c + q] * d : 0.f, h, l);
                const uint32_t o = tc_off(t, c + q, TC_ROWS) >> 2;
                a_hi[o] = h;
                a_lo[o] = l;
            }
        }
        if (mine) eat[idx] = (double)e_atom;
        tc_fence_proxy_async();
        tc_fence_before_sync();
        __syncthreads();
        if (t == 0) {
            tc_fence_after_sync();
            tc_gemm(tmem + 96, sa_hi, sa_lo, tc_smem_u32(b3_hi), tc_smem_u32(b3_lo), h1p, h2p, &bar);
        }
        ok = tc_mbar_wait(&bar, phase) && ok;
        phase ^= 1u;
        tc_fence_after_sync();

        // ---- epilogue 3: A4 = dE/dh1 * act'(z1) ----------------------------------------
        for (int c = 0; c < h1p; c += 16) {
            float v[16];
            tc_ld16(lane_addr + 96u + (uint32_t)c, v);
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                float h, l;
                tc_split(v[q] * da1[(c + q) * TC_ROWS + t], h, l);
                const uint32_t o = tc_off(t, c + q, TC_ROWS) >> 2;
                a_hi[o] = h;
                a_lo[o] = l;
            }
        }
        tc_fence_proxy_async();
        tc_fence_before_sync();
        __syncthreads();
        if (t == 0) {
            tc_fence_after_sync();
            tc_gemm(tmem + 160, sa_hi, sa_lo, tc_smem_u32(b4_hi), tc_smem_u32(b4_lo), dp, h1p, &bar);
        }
        ok = tc_mbar_wait(&bar, phase) && ok;
        phase ^= 1u;
        tc_fence_after_sync();

        // ---- epilogue 4: dE/dG (chain rule of the min-max map) -------------------------
        for (int c = 0; c < dp; c += 16) {
            float v[16];
            tc_ld16(lane_addr + 160u + (uint32_t)c, v);
            if (mine) {
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    const int k = c + q;
                    if (k < dim) {
                        double g = (double)v[q];
                        if (net.has_minmax) {
                            const double den = blob[net.xhi + k] - blob[net.xlo + k];
                            g = den != 0.0 ? -g / den : 0.0;
                        }
                        dEdG[(size_t)idx * dim + k] = g;
                    }
                }
            }
        }
        // the next tile overwrites A and TMEM: every lane's tcgen05.ld has completed
        tc_fence_before_sync();
        __syncthreads();
        tc_fence_after_sync();
    }
    if (!ok && t == 0) atomicExch(status, 1);

    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after_sync();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem),
                     "r"((uint32_t)TC_TMEM_COLS)
                     : "memory");
    }
}
