// Rebuild-time kernels of the slab decomposition ("exchange migrated atoms", the
// send sets of the halo exchange).  The reference has no counterpart (single process).
//
// Both are STABLE partitions of the rank's atoms into a few classes, done as count -> scan ->
// scatter with block-local ranks from warp ballots: no atomics, so the order of the kept /
// sent atoms -- and with it every later summation order -- is reproducible.  They replace
// ~60 element-wise / select launches of the host framework and four of its read-backs; at 8
// ranks a rebuild was bound by exactly that host work.
#include <cstdint>

#include "tab200.h"
#include "tab_internal.h"

#define DD_T 256
#define DD_MAX_CLASSES 3

// x wrapped into [0, lx) with the host framework's remainder (fmod + sign fix-up; exactly lx
// -> 0), then the owner's slab.  0 keep, 1 to the left neighbour, 2 to the right, 3 lost.
__device__ __forceinline__ int dd_owner_class(double &x, double lx, double width, int world,
                                              int rank) {
    double m = fmod(x, lx);
    if (m != 0.0 && m < 0.0) m += lx;
    if (m >= lx) m = 0.0;
    x = m;
    int owner = (int)floor(m / width);
    owner = min(max(owner, 0), world - 1);
    if (owner == rank) return 0;
    if (world == 2) return 2;          // left == right: one other rank
    if (owner == (rank + world - 1) % world) return 1;
    if (owner == (rank + 1) % world) return 2;
    return 3;
}

// -- generic machinery: per-block class counts, scan over the blocks, block-local ranks -----
// class bits of one thread: bit c set = the thread's atom belongs to class c (send sets: an
// atom may be in both).
__device__ __forceinline__ void dd_block_counts(unsigned bits, int n_classes, int *blk_cnt,
                                                int nblocks) {
    for (int c = 0; c < n_classes; ++c) {
        const int cnt = __syncthreads_count((bits >> c) & 1u);
        if (threadIdx.x == 0) blk_cnt[c * nblocks + blockIdx.x] = cnt;
    }
}

// one block: exclusive scan of every class's block counts (in place), totals -> counts[c]
__global__ void __launch_bounds__(1024)
k_dd_scan(int n_classes, int nblocks, int *__restrict__ blk, int *__restrict__ counts) {
    __shared__ int warp_tot[32];
    __shared__ int carry;
    const int tid = threadIdx.x;
    for (int c = 0; c < n_classes; ++c) {
        int *a = blk + c * nblocks;
        if (tid == 0) carry = 0;
        __syncthreads();
        for (int base = 0; base < nblocks; base += 1024) {
            const int i = base + tid;
            const int v = i < nblocks ? a[i] : 0;
            int incl = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int u = __shfl_up_sync(0xffffffffu, incl, d);
                if ((tid & 31) >= d) incl += u;
            }
            if ((tid & 31) == 31) warp_tot[tid >> 5] = incl;
            __syncthreads();
            int wbase = 0, all = 0;
            for (int w = 0; w < 32; ++w) {
                wbase += w < (tid >> 5) ? warp_tot[w] : 0;
                all += warp_tot[w];
            }
            const int start = carry;
            if (i < nblocks) a[i] = start + wbase + incl - v;
            __syncthreads();
            if (tid == 0) carry = start + all;
            __syncthreads();
        }
        if (tid == 0) counts[c] = carry;
        __syncthreads();
    }
}

// rank of this thread among the block's threads of class c (stable: thread order)
__device__ __forceinline__ int dd_block_rank(bool in, int *warp_cnt /* [DD_T / 32] shared */) {
    const unsigned mask = __ballot_sync(0xffffffffu, in);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) warp_cnt[warp] = __popc(mask);
    __syncthreads();
    int before = 0;
    for (int w = 0; w < warp; ++w) before += warp_cnt[w];
    __syncthreads();          // warp_cnt is reused by the next class
    return before + __popc(mask & ((1u << lane) - 1u));
}

// -- migration -----------------------------------------------------------------------------
__global__ void __launch_bounds__(DD_T)
k_dd_partition_count(int n, int ncol, const double *__restrict__ state, double lx, double width,
                     int world, int rank, int *__restrict__ blk, int nblocks) {
    const int i = blockIdx.x * DD_T + threadIdx.x;
    unsigned bits = 0;
    if (i < n) {
        double x = state[(size_t)i * ncol];
        const int cls = dd_owner_class(x, lx, width, world, rank);
        if (cls < 3) bits = 1u << cls;
    }
    dd_block_counts(bits, 3, blk, nblocks);
}

__global__ void __launch_bounds__(DD_T)
k_dd_partition_scatter(int n, int ncol, const double *__restrict__ state, double lx, double width,
                       int world, int rank, const int *__restrict__ blk, int nblocks,
                       double *__restrict__ keep, double *__restrict__ mail_left,
                       double *__restrict__ mail_right, int mail_cap,
                       int *__restrict__ counts) {
    __shared__ int warp_cnt[DD_T / 32];
    const int i = blockIdx.x * DD_T + threadIdx.x;
    int cls = -1;
    double x = 0.0;
    if (i < n) {
        x = state[(size_t)i * ncol];
        cls = dd_owner_class(x, lx, width, world, rank);
    }
    int dest = -1;
    for (int c = 0; c < 3; ++c) {
        const int r = dd_block_rank(cls == c, warp_cnt);
        if (cls == c) dest = blk[c * nblocks + blockIdx.x] + r;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        // counts[0..2] were written by the scan; lost = the rest; the mailboxes' headers
        const int lost = n - counts[0] - counts[1] - counts[2];
        counts[3] = lost;
        counts[4] = (counts[1] > mail_cap || counts[2] > mail_cap) ? 1 : 0;
        mail_left[0] = (double)counts[1];
        mail_right[0] = (double)counts[2];
    }
    if (cls < 0 || cls > 2) return;
    double *dst;
    if (cls == 0) dst = keep + (size_t)dest * ncol;
    else {
        if (dest >= mail_cap) return;          // reported through counts[4]
        dst = (cls == 1 ? mail_left : mail_right) + 1 + (size_t)dest * ncol;
    }
    const double *src = state + (size_t)i * ncol;
    dst[0] = x;
    for (int k = 1; k < ncol; ++k) dst[k] = src[k];
}

extern "C" int tab_dd_partition(const double *d_state, int32_t n, int32_t ncol, double lx,
                                double width, int32_t world, int32_t rank, double *d_keep,
                                double *d_mail_left, double *d_mail_right, int32_t mail_cap,
                                int32_t *d_counts, int32_t *d_work, void *stream) {
    if ((n > 0 && (!d_state || !d_keep)) || n < 0 || ncol < 1 || !d_mail_left || !d_mail_right ||
        !d_counts || !d_work || world < 1 || rank < 0 || rank >= world || !(lx > 0.0) ||
        !(width > 0.0)) {
        tab_set_error("tab_dd_partition: invalid argument");
        return TAB_EINVAL;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int nblocks = n > 0 ? (n + DD_T - 1) / DD_T : 1;
    k_dd_partition_count<<<nblocks, DD_T, 0, st>>>(n, ncol, d_state, lx, width, world, rank,
                                                   d_work, nblocks);
    TAB_LAUNCH_CHECK();
    k_dd_scan<<<1, 1024, 0, st>>>(3, nblocks, d_work, d_counts);
    TAB_LAUNCH_CHECK();
    k_dd_partition_scatter<<<nblocks, DD_T, 0, st>>>(n, ncol, d_state, lx, width, world, rank,
                                                     d_work, nblocks, d_keep, d_mail_left,
                                                     d_mail_right, mail_cap, d_counts);
    TAB_LAUNCH_CHECK();
    return TAB_OK;
}

// -- send sets -------------------------------------------------------------------------------
__global__ void __launch_bounds__(DD_T)
k_dd_send_count(int n, const double *__restrict__ pos, double x_left_below, double x_right_from,
                int *__restrict__ blk, int nblocks) {
    const int i = blockIdx.x * DD_T + threadIdx.x;
    unsigned bits = 0;
    if (i < n) {
        const double x = pos[3 * (size_t)i];
        bits = (x < x_left_below ? 1u : 0u) | (x >= x_right_from ? 2u : 0u);
    }
    dd_block_counts(bits, 2, blk, nblocks);
}

__global__ void __launch_bounds__(DD_T)
k_dd_send_scatter(int n, const double *__restrict__ pos, double x_left_below,
                  double x_right_from, const int *__restrict__ blk, int nblocks,
                  long long *__restrict__ idx_left, long long *__restrict__ idx_right) {
    __shared__ int warp_cnt[DD_T / 32];
    const int i = blockIdx.x * DD_T + threadIdx.x;
    bool in_l = false, in_r = false;
    if (i < n) {
        const double x = pos[3 * (size_t)i];
        in_l = x < x_left_below;
        in_r = x >= x_right_from;
    }
    const int rl = dd_block_rank(in_l, warp_cnt);
    const int rr = dd_block_rank(in_r, warp_cnt);
    if (in_l) idx_left[blk[blockIdx.x] + rl] = i;
    if (in_r) idx_right[blk[nblocks + blockIdx.x] + rr] = i;
}

extern "C" int tab_dd_send_sets(const double *d_pos, int32_t n, double x_left_below,
                                double x_right_from, int64_t *d_idx_left, int64_t *d_idx_right,
                                int32_t *d_counts, int32_t *d_work, void *stream) {
    if (n < 0 || !d_counts || (n > 0 && (!d_pos || !d_idx_left || !d_idx_right || !d_work))) {
        tab_set_error("tab_dd_send_sets: invalid argument");
        return TAB_EINVAL;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) {
        TAB_CUDA(cudaMemsetAsync(d_counts, 0, 2 * sizeof(int32_t), st));
        return TAB_OK;
    }
    const int nblocks = n > 0 ? (n + DD_T - 1) / DD_T : 1;
    k_dd_send_count<<<nblocks, DD_T, 0, st>>>(n, d_pos, x_left_below, x_right_from, d_work,
                                              nblocks);
    TAB_LAUNCH_CHECK();
    k_dd_scan<<<1, 1024, 0, st>>>(2, nblocks, d_work, d_counts);
    TAB_LAUNCH_CHECK();
    k_dd_send_scatter<<<nblocks, DD_T, 0, st>>>(n, d_pos, x_left_below, x_right_from, d_work,
                                                nblocks, reinterpret_cast<long long *>(d_idx_left),
                                                reinterpret_cast<long long *>(d_idx_right));
    TAB_LAUNCH_CHECK();
    return TAB_OK;
}
