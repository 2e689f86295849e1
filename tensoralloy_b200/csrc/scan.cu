// scan.cu -- exclusive prefix sums (uint32) and small shared utilities.
#include <stdarg.h>

#include "tab_internal.h"

// ---------------------------------------------------------------------------
// error string + launch counter + DevBuf
// ---------------------------------------------------------------------------
static thread_local char g_err[512] = "";
long long g_tab_launches = 0;

void tab_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char *tab_last_error(void) { return g_err; }
extern "C" int tab_version(void) { return 100; }
extern "C" int64_t tab_launch_count(void) { return g_tab_launches; }
extern "C" void tab_launch_count_reset(void) { g_tab_launches = 0; }

int DevBuf::ensure(size_t bytes) {
    if (bytes <= cap && p) return TAB_OK;
    if (p) {
        cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    size_t want = bytes + bytes / 8 + 256;   // headroom: MD sizes fluctuate
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
        tab_set_error("cudaMalloc(%zu) -> %s", want, cudaGetErrorString(e));
        p = nullptr;
        return TAB_ENOMEM;
    }
    cap = want;
    return TAB_OK;
}

void DevBuf::release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
}

// ---------------------------------------------------------------------------
// scan: blocks of SCAN_T threads x SCAN_E elements; block sums scanned
// recursively.
// ---------------------------------------------------------------------------
#define SCAN_T 256
#define SCAN_E 8
#define SCAN_B (SCAN_T * SCAN_E)

__global__ void __launch_bounds__(SCAN_T)
k_scan_block(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, int n,
             uint32_t *__restrict__ block_sums) {
    __shared__ uint32_t warp_sums[SCAN_T / 32];
    const int base = blockIdx.x * SCAN_B + threadIdx.x * SCAN_E;
    uint32_t v[SCAN_E];
    uint32_t run = 0;
#pragma unroll
    for (int k = 0; k < SCAN_E; ++k) {
        uint32_t x = (base + k < n) ? in[base + k] : 0u;
        v[k] = run;
        run += x;
    }
    // inclusive scan of `run` over the block
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = run;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = (lane < SCAN_T / 32) ? warp_sums[lane] : 0u;
#pragma unroll
        for (int d = 1; d < SCAN_T / 32; d <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, w, d);
            if (lane >= d) w += t;
        }
        if (lane < SCAN_T / 32) warp_sums[lane] = w;
    }
    __syncthreads();
    uint32_t offset = inc - run + (warp ? warp_sums[warp - 1] : 0u);
#pragma unroll
    for (int k = 0; k < SCAN_E; ++k)
        if (base + k < n) out[base + k] = v[k] + offset;
    if (threadIdx.x == SCAN_T - 1 && block_sums)
        block_sums[blockIdx.x] = offset + run;
}

__global__ void k_scan_add(uint32_t *__restrict__ out, int n,
                           const uint32_t *__restrict__ block_offsets) {
    const int i = blockIdx.x * SCAN_B + threadIdx.x;
    const uint32_t off = block_offsets[blockIdx.x];
    for (int k = 0; k < SCAN_E; ++k) {
        int idx = i + k * SCAN_T;
        if (idx < n) out[idx] += off;
    }
}

// total = last exclusive value + last input, written as u64 (single thread)
__global__ void k_scan_total(const uint32_t *__restrict__ in_last_excl,
                             const uint32_t *__restrict__ in_last_val,
                             unsigned long long *total) {
    *total = (unsigned long long)(*in_last_excl) + (unsigned long long)(*in_last_val);
}

static int scan_rec(const uint32_t *d_in, uint32_t *d_out, int n, uint32_t *tmp,
                    cudaStream_t st) {
    const int nblk = (n + SCAN_B - 1) / SCAN_B;
    if (nblk == 1) {
        k_scan_block<<<1, SCAN_T, 0, st>>>(d_in, d_out, n, nullptr);
        TAB_LAUNCH_CHECK();
        return TAB_OK;
    }
    uint32_t *sums = tmp;
    k_scan_block<<<nblk, SCAN_T, 0, st>>>(d_in, d_out, n, sums);
    TAB_LAUNCH_CHECK();
    TAB_TRY(scan_rec(sums, sums, nblk, tmp + ((nblk + 31) & ~31), st));
    k_scan_add<<<nblk, SCAN_T, 0, st>>>(d_out, n, sums);
    TAB_LAUNCH_CHECK();
    return TAB_OK;
}

int tab_scan_exclusive_u32(const uint32_t *d_in, uint32_t *d_out, int n,
                           unsigned long long *d_total, DevBuf &tmp,
                           cudaStream_t st) {
    if (n <= 0) {
        if (d_total) TAB_CUDA(cudaMemsetAsync(d_total, 0, 8, st));
        return TAB_OK;
    }
    // scratch: block sums of every level (+ the saved last input when in-place)
    size_t need = 64;
    for (int m = (n + SCAN_B - 1) / SCAN_B; m > 1; m = (m + SCAN_B - 1) / SCAN_B)
        need += ((m + 31) & ~31);
    need += 64;
    TAB_TRY(tmp.ensure((need + 8) * sizeof(uint32_t)));
    uint32_t *t = tmp.as<uint32_t>();
    uint32_t *saved_last = t;       // in-place scans overwrite in[n-1]
    if (d_total)
        TAB_CUDA(cudaMemcpyAsync(saved_last, d_in + (n - 1), sizeof(uint32_t),
                                 cudaMemcpyDeviceToDevice, st));
    TAB_TRY(scan_rec(d_in, d_out, n, t + 32, st));
    if (d_total) {
        k_scan_total<<<1, 1, 0, st>>>(d_out + (n - 1), saved_last, d_total);
        TAB_LAUNCH_CHECK();
    }
    return TAB_OK;
}
