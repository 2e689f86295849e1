// tab_internal.h -- structures shared by the translation units of libtab200.so.
// Not part of the public ABI (include/tab200.h is).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "tab200.h"

// ---------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------
void tab_set_error(const char *fmt, ...);
extern long long g_tab_launches;

#define TAB_CUDA(expr)                                                         \
    do {                                                                       \
        cudaError_t _e = (expr);                                               \
        if (_e != cudaSuccess) {                                               \
            tab_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,        \
                          cudaGetErrorString(_e));                             \
            return TAB_ECUDA;                                                  \
        }                                                                      \
    } while (0)

#define TAB_LAUNCH_CHECK()                                                     \
    do {                                                                       \
        ++g_tab_launches;                                                      \
        cudaError_t _e = cudaGetLastError();                                   \
        if (_e != cudaSuccess) {                                               \
            tab_set_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__,    \
                          cudaGetErrorString(_e));                             \
            return TAB_ECUDA;                                                  \
        }                                                                      \
    } while (0)

#define TAB_TRY(expr)                                                          \
    do {                                                                       \
        int _r = (expr);                                                       \
        if (_r != TAB_OK) return _r;                                           \
    } while (0)

// ---------------------------------------------------------------------------
// device data layout
// ---------------------------------------------------------------------------
// One atom record = one 32-byte sector: position (cell-sorted, ghosts already
// shifted by S.h) + one model slot w (F'(rho) between the two EAM passes), so a
// neighbour gather is exactly one sector.
struct __align__(32) Atom4 {
    double x, y, z, w;
};

// Neighbour entries: low 28 bits = index into the extended (owned + ghost) atom
// array, high 4 bits = element index of the neighbour.
#define TAB_COL_IDX_MASK 0x0FFFFFFFu
#define TAB_COL_TYPE_SHIFT 28
#define TAB_COL_PAD 0xFFFFFFFFu
#define TAB_MAX_ELEMENTS 16
#define TAB_SLICE 32          // atoms per ELL slice = one warp
#define TAB_TILE_B 2          // owned cells are ordered in B x B x B tiles

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes);   // grow-only, contents NOT preserved
    void release();
    template <typename T> T *as() const { return reinterpret_cast<T *>(p); }
};

struct Grid {
    double h[9];      // lattice, rows = vectors
    double hinv[9];   // inverse: scaled = pos @ hinv
    int nb[3];        // owned bins per direction
    int ne[3];        // extended bins = nb + 2 g
    int g[3];         // ghost layers per side
    int sr[3];        // search range in bins
    int pbc[3];
    int tl[3];        // tiles per direction = ceil(nb / TAB_TILE_B)
    int n_slots;      // owned cell slots = tl0*tl1*tl2*B^3
    int n_ecells;     // extended cells
    double rc, rc2;
};

struct tab_nbr {
    int n = 0;            // owned atoms
    int n_ghost = 0;
    int n_ext = 0;
    int n_slices = 0;
    long long nij = 0;
    long long ell_rows = 0;   // total rows of 32 entries
    int nnl_max = 0;
    bool built = false;
    Grid grid;
    // per owned atom (caller order)
    DevBuf cell_of;       // int32 [n]   owned-cell rank
    DevBuf s0;            // int32 [n]   packed wrap shift
    DevBuf types_in;      // int32 [n]   copy of caller types (caller order)
    // per owned atom (sorted order)
    DevBuf perm;          // int32 [n]   sorted -> caller index
    DevBuf counts;        // int32 [n]   neighbours per atom
    // per extended atom
    DevBuf atoms;         // Atom4 [n_ext]
    DevBuf types_ext;     // uint8 [n_ext]
    DevBuf ghost_owner;   // int32 [n_ghost]  sorted owned index of the source
    DevBuf ghost_S;       // int32 [n_ghost]  packed image shift
    // cells
    DevBuf cell_count;    // uint32 [n_slots]
    DevBuf cell_start;    // uint32 [n_slots]
    DevBuf cell_fill;     // uint32 [n_slots]
    DevBuf ext_start;     // uint32 [n_ecells]
    DevBuf ext_count;     // uint32 [n_ecells]
    DevBuf gcount;        // uint32 [n_ecells]
    DevBuf gstart;        // uint32 [n_ecells]
    // ELL
    DevBuf slice_w;       // uint32 [n_slices]  width (rows) of each slice
    DevBuf slice_ptr;     // uint32 [n_slices]  first row of each slice
    DevBuf col;           // uint32 [ell_rows*32]
    // scratch
    DevBuf scan_tmp;
    DevBuf stats;         // device: [0]=nij (u64) [1]=nnl_max [2]=total (u64 scan totals)
    DevBuf row_ptr;       // uint32 [n] (export)
    // model scratch that lives with the structure (per ext atom / per block)
    DevBuf rho;           // double [n] sorted
    DevBuf partial;       // double [blocks*16] block partial sums
    DevBuf adp;           // double [n_ext*9] (ADP moments)
};

// ---------------------------------------------------------------------------
// helpers implemented in scan.cu
// ---------------------------------------------------------------------------
// out[i] = sum_{k<i} in[i]; *d_total (device, u64) = sum of all.  in == out allowed.
int tab_scan_exclusive_u32(const uint32_t *d_in, uint32_t *d_out, int n,
                           unsigned long long *d_total, DevBuf &tmp,
                           cudaStream_t st);

// packed shifts: 3 x 10-bit biased by 512
__host__ __device__ inline int tab_pack_shift(int a, int b, int c) {
    return ((a + 512) & 1023) | (((b + 512) & 1023) << 10) | (((c + 512) & 1023) << 20);
}
__host__ __device__ inline void tab_unpack_shift(int v, int &a, int &b, int &c) {
    a = (v & 1023) - 512;
    b = ((v >> 10) & 1023) - 512;
    c = ((v >> 20) & 1023) - 512;
}
