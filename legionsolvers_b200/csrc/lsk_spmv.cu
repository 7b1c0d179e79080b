// lsk_spmv.cu -- CSR and COO sparse matrix-vector products for sm_100a.
//
// Replaces cusparseSpMV and its per-call set-up in the reference's src/CSRMatrixTasks.cu,
// src/COOMatrixTasks.cu and src/CuSPARSEHelpers.hpp.  The kernels read the Legion field layout as
// it is (fp64/fp32 values, int64 column ids, rowptr = inclusive Rect<1> of GLOBAL k, x shifted to
// global column 0), so there is no indptr conversion pass, no descriptor and no workspace.
//
// HBM-bound (AI ~ 0.12 flop/B): no tensor cores.  What matters is (i) streaming the 16 B/nnz of
// values+columns with full-width 256-bit coalesced loads that bypass L1 and are marked evict-first
// in L2, (ii) keeping x on the cached path (L1 + 126 MB L2) so its traffic stays compulsory, and
// (iii) enough bytes in flight per SM.
#include <limits.h>
#include <stdlib.h>

#include <type_traits>

#include "lsk_common.cuh"
#include "lsk_spmv_tma.cuh"
#include "lsk_spmv_ws.cuh"

namespace lsk {

constexpr int kTile = 2048;  // products staged per tile: 16 KB (fp64) of shared memory per CTA

// ---- small load helpers ---------------------------------------------------------------------------
__device__ __forceinline__ void load4_stream(const double *p, double (&v)[4]) {
    const Pack32 q = ld256_stream(p);
#pragma unroll
    for (int e = 0; e < 4; ++e) v[e] = PackOf<double>::get(q, e);
}
__device__ __forceinline__ void load4_stream(const float *p, float (&v)[4]) {
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3])
                 : "l"(p));
}
__device__ __forceinline__ void store4_shared(double *s, const double (&v)[4]) {
    *reinterpret_cast<double2 *>(s) = make_double2(v[0], v[1]);
    *reinterpret_cast<double2 *>(s + 2) = make_double2(v[2], v[3]);
}
__device__ __forceinline__ void store4_shared(float *s, const float (&v)[4]) {
    *reinterpret_cast<float4 *>(s) = make_float4(v[0], v[1], v[2], v[3]);
}

// ===================================================================================================
// CSR "stream" kernel.  A CTA owns `rpb` consecutive rows (one thread per row).  The non-zeros of
// those rows are one contiguous run of k, which the whole CTA streams in tiles of kTile with
// perfectly coalesced 256-bit loads, multiplying by the gathered x and parking the ROUNDED products
// in shared memory; each row's thread then adds its products in ascending k.  That is the
// reference CPU body's order and rounding (src/CSRMatrixTasks.cpp:73-91: `+= entry * x`, product
// rounded, k ascending), so the result is bit-identical to it -- and lane utilisation does not
// depend on the row length (7-point rows would idle 25 of 32 lanes in a warp-per-row kernel).
// ===================================================================================================
template <typename T, bool VEC, int NDOT>
__global__ void __launch_bounds__(kBlock, 4)
csr_stream_kernel(int64_t rows, int rpb, int64_t n_row_blocks, const T *__restrict__ entry,
                  const long long *__restrict__ col, const lsk_rect *__restrict__ rowptr, int64_t k_base,
                  const T *__restrict__ x, T *__restrict__ y, const T *__restrict__ dot_w,
                  double *partials, unsigned int *ticket, T *out_yw, T *out_yy, const lsk_peers *peers, bool accumulate) {
    __shared__ __align__(16) T s_prod[kTile];
    __shared__ long long s_lo[kWarps], s_hi[kWarps];
    const int tid = threadIdx.x;
    double dacc[NDOT > 0 ? NDOT : 1];
#pragma unroll
    for (int j = 0; j < (NDOT > 0 ? NDOT : 1); ++j) dacc[j] = 0.0;

    for (int64_t rb = blockIdx.x; rb < n_row_blocks; rb += gridDim.x) {
        const int64_t r0 = rb * rpb;
        const int64_t left = rows - r0;
        const int nr = (int) (left < rpb ? left : rpb);
        const bool have = tid < nr;
        // my row's run [lo, hi1) in piece-local element indices
        long long lo = LLONG_MAX, hi1 = LLONG_MIN;
        if (have) {
            const longlong2 rc = __ldg(reinterpret_cast<const longlong2 *>(rowptr + r0 + tid));
            if (rc.y >= rc.x) {
                lo = rc.x - k_base;
                hi1 = rc.y + 1 - k_base;
            }
        }
        // the CTA's run [jb, je): min/max over its non-empty rows
        {
            const long long wl = warp_min_ll(lo), wh = warp_max_ll(hi1);
            if ((tid & 31) == 0) {
                s_lo[tid >> 5] = wl;
                s_hi[tid >> 5] = wh;
            }
        }
        __syncthreads();
        long long jb = s_lo[0], je = s_hi[0];
#pragma unroll
        for (int w = 1; w < kWarps; ++w) {
            jb = s_lo[w] < jb ? s_lo[w] : jb;
            je = s_hi[w] > je ? s_hi[w] : je;
        }

        T acc = (T) 0;
        if (je > jb) {
            // start tiles on a 32-byte boundary of `col` (entry is congruent when VEC)
            const int mis = VEC ? (int) ((reinterpret_cast<uintptr_t>(col + jb) >> 3) & 3) : 0;
            for (long long t0 = jb - mis; t0 < je; t0 += kTile) {
                if (VEC) {
                    constexpr int U = kTile / (4 * kBlock);
                    long long c[U][4];
                    T v[U][4];
                    bool full[U];
#pragma unroll
                    for (int u = 0; u < U; ++u) {  // issue every streamed load first
                        const long long j = t0 + (long long) (u * kBlock + tid) * 4;
                        full[u] = (j >= jb) && (j + 4 <= je);
                        if (full[u]) {
                            const Pack32 pc = ld256_stream(col + j);
#pragma unroll
                            for (int e = 0; e < 4; ++e) c[u][e] = (long long) pc.q[e];
                            load4_stream(entry + j, v[u]);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const int slot = (u * kBlock + tid) * 4;
                        const long long j = t0 + slot;
                        if (full[u]) {
                            T xv[4];
#pragma unroll
                            for (int e = 0; e < 4; ++e) xv[e] = __ldg(x + c[u][e]);
#pragma unroll
                            for (int e = 0; e < 4; ++e) v[u][e] = mul_rn(v[u][e], xv[e]);
                            store4_shared(s_prod + slot, v[u]);
                        } else {
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const long long jj = j + e;
                                if (jj >= jb && jj < je)
                                    s_prod[slot + e] = mul_rn(load1_stream(entry + jj), __ldg(x + load1_stream(col + jj)));
                            }
                        }
                    }
                } else {
#pragma unroll 4
                    for (int u = 0; u < kTile / kBlock; ++u) {
                        const int slot = u * kBlock + tid;
                        const long long j = t0 + slot;
                        if (j < je) s_prod[slot] = mul_rn(load1_stream(entry + j), __ldg(x + load1_stream(col + j)));
                    }
                }
                __syncthreads();
                if (have) {
                    const long long a = lo > t0 ? lo : t0;
                    const long long b = hi1 < t0 + kTile ? hi1 : t0 + kTile;
                    for (long long j = a; j < b; ++j) acc = add_rn(acc, s_prod[j - t0]);
                }
                __syncthreads();
            }
        } else {
            __syncthreads();  // s_lo/s_hi are rewritten by the next row block
        }
        if (have) {
            if (accumulate) acc = add_rn(y[r0 + tid], acc);  // a further block on the same rows (beta = 1)
            y[r0 + tid] = acc;
            if constexpr (NDOT >= 1) dacc[0] = fma((double) acc, dot_w != y ? (double) __ldg(dot_w + r0 + tid) : (double) acc, dacc[0]);
            if constexpr (NDOT >= 2) dacc[NDOT - 1] = fma((double) acc, (double) acc, dacc[NDOT - 1]);
        }
    }
    if constexpr (NDOT > 0) {
        T *out[NDOT];
        out[0] = out_yw;
        if constexpr (NDOT >= 2) out[NDOT - 1] = out_yy;
        grid_reduce_finish<NDOT, T>(dacc, partials, ticket, out, peers);
    }
}

// ===================================================================================================
// Software-pipelined form of the stream kernel (the one used whenever entry/col are mutually
// aligned).  Same arithmetic, same result bits.  The un-pipelined kernel pays three dependent
// memory latencies per row block (rowptr -> streamed tile -> x gather) with nothing in flight in
// between; here every thread (i) prefetches its rect of the NEXT row block one block ahead and
// (ii) issues the 256-bit streamed loads of the next tile BEFORE it adds up the current tile, so
// HBM requests are outstanding while the CTA is in its shared-memory phase.
// ===================================================================================================
template <typename T, int NDOT>
__global__ void __launch_bounds__(kBlock, 3)
csr_stream_pipe_kernel(int64_t rows, int rpb, int64_t n_row_blocks, const T *__restrict__ entry,
                       const long long *__restrict__ col, const lsk_rect *__restrict__ rowptr,
                       int64_t k_base, const T *__restrict__ x, T *__restrict__ y,
                       const T *__restrict__ dot_w, double *partials, unsigned int *ticket, T *out_yw,
                       T *out_yy, const lsk_peers *peers, bool accumulate) {
    constexpr int U = kTile / (4 * kBlock);
    __shared__ __align__(16) T s_prod[kTile];
    __shared__ long long s_lo[kWarps], s_hi[kWarps];
    const int tid = threadIdx.x;
    const int64_t G = gridDim.x;
    double dacc[NDOT > 0 ? NDOT : 1];
#pragma unroll
    for (int j = 0; j < (NDOT > 0 ? NDOT : 1); ++j) dacc[j] = 0.0;

    // my row's run [lo, hi1) of row block rb, piece-local element indices (empty: lo > hi1)
    auto load_rect = [&](int64_t rb, long long &lo, long long &hi1) {
        lo = LLONG_MAX;
        hi1 = LLONG_MIN;
        if (rb < n_row_blocks) {
            const int64_t r = rb * rpb + tid;
            if (tid < rpb && r < rows) {
                const longlong2 rc = __ldg(reinterpret_cast<const longlong2 *>(rowptr + r));
                if (rc.y >= rc.x) {
                    lo = rc.x - k_base;
                    hi1 = rc.y + 1 - k_base;
                }
            }
        }
    };
    auto publish_span = [&](long long lo, long long hi1) {
        const long long wl = warp_min_ll(lo), wh = warp_max_ll(hi1);
        if ((tid & 31) == 0) {
            s_lo[tid >> 5] = wl;
            s_hi[tid >> 5] = wh;
        }
    };
    auto read_span = [&](long long &jb, long long &je) {
        jb = s_lo[0];
        je = s_hi[0];
#pragma unroll
        for (int w = 1; w < kWarps; ++w) {
            jb = s_lo[w] < jb ? s_lo[w] : jb;
            je = s_hi[w] > je ? s_hi[w] : je;
        }
        if (je <= jb) jb = je = 0;  // no non-zeros in this row block: one empty tile
    };
    long long c[U][4];
    T v[U][4];
    bool full[U];
    auto issue_loads = [&](long long t0, long long jb, long long je) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long j = t0 + (long long) (u * kBlock + tid) * 4;
            full[u] = (j >= jb) && (j + 4 <= je);
            if (full[u]) {
                const Pack32 pc = ld256_stream(col + j);
#pragma unroll
                for (int e = 0; e < 4; ++e) c[u][e] = (long long) pc.q[e];
                load4_stream(entry + j, v[u]);
            }
        }
    };
    auto tile_start = [&](long long jb) {
        return jb - (long long) ((reinterpret_cast<uintptr_t>(col + jb) >> 3) & 3);
    };

    int64_t rb = blockIdx.x;
    long long lo, hi1, jb, je, nlo, nhi1;
    load_rect(rb, lo, hi1);
    publish_span(lo, hi1);
    __syncthreads();
    read_span(jb, je);
    __syncthreads();
    long long t0 = tile_start(jb);
    issue_loads(t0, jb, je);
    load_rect(rb + G, nlo, nhi1);  // in flight until the span exchange below
    T acc = (T) 0;

    while (rb < n_row_blocks) {
        // ---- products of the tile held in registers -> shared memory
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int slot = (u * kBlock + tid) * 4;
            const long long j = t0 + slot;
            if (full[u]) {
                T xv[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) xv[e] = __ldg(x + c[u][e]);
#pragma unroll
                for (int e = 0; e < 4; ++e) v[u][e] = mul_rn(v[u][e], xv[e]);
                store4_shared(s_prod + slot, v[u]);
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const long long jj = j + e;
                    if (jj >= jb && jj < je)
                        s_prod[slot + e] = mul_rn(load1_stream(entry + jj), __ldg(x + load1_stream(col + jj)));
                }
            }
        }
        const bool last_tile = (t0 + kTile >= je);
        if (last_tile) publish_span(nlo, nhi1);
        __syncthreads();
        // ---- put the next tile's HBM loads in flight, then add up this tile
        const long long cur_t0 = t0;
        long long njb = 0, nje = 0;
        if (last_tile) {
            read_span(njb, nje);
            if (rb + G < n_row_blocks) issue_loads(tile_start(njb), njb, nje);
        } else {
            issue_loads(t0 + kTile, jb, je);
        }
        {
            const long long a = lo > cur_t0 ? lo : cur_t0;
            const long long b = hi1 < cur_t0 + kTile ? hi1 : cur_t0 + kTile;
            for (long long j = a; j < b; ++j) acc = add_rn(acc, s_prod[j - cur_t0]);
        }
        if (last_tile) {
            const int64_t r = rb * rpb + tid;
            if (tid < rpb && r < rows) {
                if (accumulate) acc = add_rn(y[r], acc);
                y[r] = acc;
                if constexpr (NDOT >= 1) dacc[0] = fma((double) acc, dot_w != y ? (double) __ldg(dot_w + r) : (double) acc, dacc[0]);
                if constexpr (NDOT >= 2) dacc[NDOT - 1] = fma((double) acc, (double) acc, dacc[NDOT - 1]);
            }
            acc = (T) 0;
            rb += G;
            lo = nlo;
            hi1 = nhi1;
            jb = njb;
            je = nje;
            t0 = tile_start(jb);
            load_rect(rb + G, nlo, nhi1);
        } else {
            t0 += kTile;
        }
        __syncthreads();
    }
    if constexpr (NDOT > 0) {
        T *out[NDOT];
        out[0] = out_yw;
        if constexpr (NDOT >= 2) out[NDOT - 1] = out_yy;
        grid_reduce_finish<NDOT, T>(dacc, partials, ticket, out, peers);
    }
}

// ===================================================================================================
// TMA-staged stream kernel (fp64; the default for short and medium rows).
//
// ncu on the register-staged kernel above showed DRAM traffic equal to the algorithmic bytes but the
// L1 data pipe as the busiest unit: with lanes striding the non-zero stream, one 32-lane x-gather
// touches ~18 rows x 7 stencil legs = ~11 cache lines, and every product crosses shared memory twice.
// This kernel flips the mapping.  The CTA's run of (col, entry) is copied global->shared by the TMA
// engine (cp.async.bulk, 16 KB per array per tile, L2 evict-first, completion on an mbarrier, two
// stages so the next tile is in flight while this one is consumed); threads then own ROWS: lane l
// of a warp walks row r0+l, so gather e of a warp reads leg e of 32 consecutive rows -- one or two
// lines for banded matrices -- and shared-memory reads are stride-(row length) 8-byte accesses
// (conflict-free for odd lengths).  Each thread adds its rounded products in ascending k in a
// register: bit-identical to the reference CPU body, no product staging, one barrier per tile.
// ===================================================================================================
template <int NDOT, int LPR = 1>
__global__ void __launch_bounds__(kBlock, LSK_TMA_MINB)
csr_tma_kernel(TmaSpmvArgs a, double *partials, unsigned int *ticket, double *out_yw, double *out_yy, const lsk_peers *peers) {
    extern __shared__ __align__(128) unsigned char s_dyn[];
    __shared__ __align__(8) uint64_t s_full[kTmaStages];
    __shared__ long long s_lo[2][kWarps], s_hi[2][kWarps];
    double dacc[NDOT > 0 ? NDOT : 1];
#pragma unroll
    for (int j = 0; j < (NDOT > 0 ? NDOT : 1); ++j) dacc[j] = 0.0;
    pdl_launch_dependents();
    TmaSpmvState st;
    csr_tma_init(st, s_full);
    // prologue (rowptr rects of the first row blocks, TMA copy of the first matrix tile) does not depend on the kernel
    // before this one; x, the fused dot's w and y do
    TmaCursor cur;
    csr_tma_run<NDOT, 1, LPR>(a, st, s_dyn, s_full, s_lo, s_hi, dacc, &cur);
    pdl_wait();
    csr_tma_run<NDOT, 2, LPR>(a, st, s_dyn, s_full, s_lo, s_hi, dacc, &cur);
    if constexpr (NDOT > 0) {
        double *out[NDOT];
        out[0] = out_yw;
        if constexpr (NDOT >= 2) out[NDOT - 1] = out_yy;
        grid_reduce_finish<NDOT, double>(dacc, partials, ticket, out, peers);
    }
}

// ===================================================================================================
// CSR "vector" kernel: V lanes per row (V = 32 is warp-per-row).  Lanes stride the row, partial
// sums are combined with warp shuffles.  For long rows (dense blocks, power-law heads) where a
// single thread adding the whole row would serialise.  Tree order + fma: <= 1e-12 relative.
// ===================================================================================================
template <typename T, int V, int NDOT>
__global__ void __launch_bounds__(kBlock)
csr_vector_kernel(int64_t rows, const T *__restrict__ entry, const long long *__restrict__ col,
                  const lsk_rect *__restrict__ rowptr, int64_t k_base, const T *__restrict__ x,
                  T *__restrict__ y, const T *__restrict__ dot_w, double *partials, unsigned int *ticket,
                  T *out_yw, T *out_yy, const lsk_peers *peers, bool accumulate) {
    constexpr int RPC = kBlock / V;  // rows per CTA pass
    const int sub = threadIdx.x % V;
    double dacc[NDOT > 0 ? NDOT : 1];
#pragma unroll
    for (int j = 0; j < (NDOT > 0 ? NDOT : 1); ++j) dacc[j] = 0.0;

    for (int64_t base = (int64_t) blockIdx.x * RPC; base < rows; base += (int64_t) gridDim.x * RPC) {
        const int64_t row = base + threadIdx.x / V;
        const bool active = row < rows;
        T sum = (T) 0;
        if (active) {
            const longlong2 rc = __ldg(reinterpret_cast<const longlong2 *>(rowptr + row));
            const long long hi = rc.y - k_base;
            for (long long j = rc.x - k_base + sub; j <= hi; j += V)
                sum = fma_rn(load1_stream(entry + j), __ldg(x + load1_stream(col + j)), sum);
        }
#pragma unroll
        for (int o = V / 2; o > 0; o >>= 1) sum += __shfl_down_sync(0xffffffffu, sum, o, V);
        if (active && sub == 0) {
            if (accumulate) sum = add_rn(y[row], sum);
            y[row] = sum;
            if constexpr (NDOT >= 1) dacc[0] = fma((double) sum, dot_w != y ? (double) __ldg(dot_w + row) : (double) sum, dacc[0]);
            if constexpr (NDOT >= 2) dacc[NDOT - 1] = fma((double) sum, (double) sum, dacc[NDOT - 1]);
        }
    }
    if constexpr (NDOT > 0) {
        T *out[NDOT];
        out[0] = out_yw;
        if constexpr (NDOT >= 2) out[NDOT - 1] = out_yy;
        grid_reduce_finish<NDOT, T>(dacc, partials, ticket, out, peers);
    }
}

// ===================================================================================================
// COO kernel: segmented warp-shuffle reduction.  Each warp takes runs of 32 consecutive non-zeros,
// forms the rounded products, does a segmented inclusive scan keyed on "row differs from the
// previous lane" (correct for ANY ordering -- unsorted input just makes shorter segments), and the
// last lane of every segment adds its total into y with one fp64 atomic (segments of one row may
// continue in the next warp).  y is accumulated into (beta = 1), as the reference calls cuSPARSE.
// ===================================================================================================
template <typename T>
__global__ void __launch_bounds__(kBlock)
coo_segreduce_kernel(int64_t nnz, const T *__restrict__ entry, const long long *__restrict__ row,
                     const long long *__restrict__ col, const T *__restrict__ x, T *__restrict__ y,
                     long long row_lo, long long row_hi, long long col_lo, long long col_hi) {
    constexpr int ITEMS = 4;  // independent 32-wide chunks in flight per warp
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t) blockIdx.x * kBlock + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t) gridDim.x * kBlock) >> 5;
    for (int64_t base = warp * (32 * ITEMS); base < nnz; base += nwarps * (32 * ITEMS)) {
        long long r[ITEMS], c[ITEMS];
        T v[ITEMS];
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            const int64_t k = base + i * 32 + lane;
            r[i] = -1;
            c[i] = -1;
            v[i] = (T) 0;
            if (k < nnz) {
                r[i] = load1_stream(row + k);
                c[i] = load1_stream(col + k);
                v[i] = load1_stream(entry + k);
            }
        }
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            const bool ok = (r[i] >= row_lo) && (r[i] <= row_hi) && (c[i] >= col_lo) && (c[i] <= col_hi);
            const long long key = ok ? r[i] : -1;  // rejected entries form their own (discarded) segments
            T p = ok ? mul_rn(v[i], __ldg(x + c[i])) : (T) 0;
            const long long prev = __shfl_up_sync(0xffffffffu, key, 1);
            const long long next = __shfl_down_sync(0xffffffffu, key, 1);
            bool f = (lane == 0) || (prev != key);  // segment head
            const bool tail = (lane == 31) || (next != key);
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const T pv = __shfl_up_sync(0xffffffffu, p, o);
                const int pf = __shfl_up_sync(0xffffffffu, (int) f, o);
                if (lane >= o && !f) {
                    p += pv;
                    f = pf != 0;
                }
            }
            if (ok && tail) atomicAdd(y + key, p);
        }
    }
}

// ---- host dispatch -----------------------------------------------------------------------------------
static int pick_variant(int64_t rows, int64_t nnz) {
    if (rows <= 0) return LSK_SPMV_STREAM;
    const double mean = (double) nnz / (double) rows;
    // row-length statistics pick the mapping: up to ~a dozen non-zeros per row a thread owns a row (a 2048-element
    // tile then feeds all 256 threads); up to ~a hundred, 2-8 lanes share a row of the same TMA-staged tile; beyond
    // that a warp per row
    return mean <= 12.0 ? LSK_SPMV_STREAM : mean <= 96.0 ? LSK_SPMV_LANES : LSK_SPMV_WARP;
}

template <typename T, bool VEC>
static void launch_stream_kernel(int ndot, int grid, cudaStream_t st, int64_t rows, int rpb, int64_t nrb,
                                 const T *entry, const long long *col, const lsk_rect *rowptr,
                                 int64_t k_base, const T *x, T *y, const T *dot_w, RedScratch rs, T *o0,
                                 T *o1, bool accumulate) {
    if (ndot == 0)
        csr_stream_kernel<T, VEC, 0><<<grid, kBlock, 0, st>>>(rows, rpb, nrb, entry, col, rowptr, k_base, x, y, dot_w, rs.partials, rs.ticket, o0, o1, rs.peers, accumulate);
    else if (ndot == 1)
        csr_stream_kernel<T, VEC, 1><<<grid, kBlock, 0, st>>>(rows, rpb, nrb, entry, col, rowptr, k_base, x, y, dot_w, rs.partials, rs.ticket, o0, o1, rs.peers, accumulate);
    else
        csr_stream_kernel<T, VEC, 2><<<grid, kBlock, 0, st>>>(rows, rpb, nrb, entry, col, rowptr, k_base, x, y, dot_w, rs.partials, rs.ticket, o0, o1, rs.peers, accumulate);
}

template <typename T>
static void launch_pipe_kernel(int ndot, int grid, cudaStream_t st, int64_t rows, int rpb, int64_t nrb,
                               const T *entry, const long long *col, const lsk_rect *rowptr, int64_t k_base,
                               const T *x, T *y, const T *dot_w, RedScratch rs, T *o0, T *o1, bool accumulate) {
    if (ndot == 0)
        csr_stream_pipe_kernel<T, 0><<<grid, kBlock, 0, st>>>(rows, rpb, nrb, entry, col, rowptr, k_base, x, y, dot_w, rs.partials, rs.ticket, o0, o1, rs.peers, accumulate);
    else if (ndot == 1)
        csr_stream_pipe_kernel<T, 1><<<grid, kBlock, 0, st>>>(rows, rpb, nrb, entry, col, rowptr, k_base, x, y, dot_w, rs.partials, rs.ticket, o0, o1, rs.peers, accumulate);
    else
        csr_stream_pipe_kernel<T, 2><<<grid, kBlock, 0, st>>>(rows, rpb, nrb, entry, col, rowptr, k_base, x, y, dot_w, rs.partials, rs.ticket, o0, o1, rs.peers, accumulate);
}

template <int LPR>
static int launch_tma_kernel_lpr(lsk_ctx *ctx, int ndot, int grid, cudaStream_t st, const TmaSpmvArgs &a, RedScratch rs, double *o0,
                                 double *o1) {
    static const int family = configure_family_index();
    const int rc = configure_once(ctx, family, [] {
        cudaError_t e = cudaFuncSetAttribute(csr_tma_kernel<0, LPR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kTmaSmem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(csr_tma_kernel<1, LPR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kTmaSmem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(csr_tma_kernel<2, LPR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kTmaSmem);
        return e;
    });
    if (rc != 0) return rc;
    cudaError_t e;
    if (ndot == 0)
        e = launch_pdl(kPdlSpmv, csr_tma_kernel<0, LPR>, grid, kBlock, kTmaSmem, st, a, rs.partials, rs.ticket, o0, o1, rs.peers);
    else if (ndot == 1)
        e = launch_pdl(kPdlSpmv, csr_tma_kernel<1, LPR>, grid, kBlock, kTmaSmem, st, a, rs.partials, rs.ticket, o0, o1, rs.peers);
    else
        e = launch_pdl(kPdlSpmv, csr_tma_kernel<2, LPR>, grid, kBlock, kTmaSmem, st, a, rs.partials, rs.ticket, o0, o1, rs.peers);
    return (int) e;
}

// lanes per row 1 (thread per row) .. 8; the row block is sized so that its non-zeros fill about one tile
static int launch_tma_kernel(lsk_ctx *ctx, int lpr, int ndot, cudaStream_t st, int64_t rows, int64_t nnz, const double *entry,
                             const long long *col, const lsk_rect *rowptr, int64_t k_base, const double *x, double *y,
                             const double *dot_w, RedScratch rs, double *o0, double *o1, bool accumulate) {
    const double mean = rows > 0 ? (double) nnz / (double) rows : 1.0;
    const int unit = 32 / lpr;  // rows per warp
    int rpb = (int) ((double) kTmaTile / (mean < 1.0 ? 1.0 : mean));
    rpb = (rpb / unit) * unit;
    if (rpb < unit) rpb = unit;
    if (rpb > kBlock / lpr) rpb = kBlock / lpr;
    const int64_t nrb = rows > 0 ? (rows + rpb - 1) / rpb : 1;
    const int64_t cap = (int64_t) ctx->sm_count * LSK_TMA_MINB;
    const int grid = (int) (nrb < cap ? nrb : cap);
    TmaSpmvArgs a;
    a.rows = rows; a.nnz = nnz; a.rpb = rpb; a.n_row_blocks = nrb; a.entry = entry; a.col = col; a.rowptr = rowptr;
    a.k_base = k_base; a.x = x; a.y = y; a.dot_w = dot_w; a.accumulate = accumulate ? 1 : 0;
    switch (lpr) {
    case 1: return launch_tma_kernel_lpr<1>(ctx, ndot, grid, st, a, rs, o0, o1);
    case 2: return launch_tma_kernel_lpr<2>(ctx, ndot, grid, st, a, rs, o0, o1);
    case 4: return launch_tma_kernel_lpr<4>(ctx, ndot, grid, st, a, rs, o0, o1);
    default: return launch_tma_kernel_lpr<8>(ctx, ndot, grid, st, a, rs, o0, o1);
    }
}

// ---- the warp-specialised kernel (lsk_spmv_ws.cuh): lanes per row 1 .. 8 ---------------------------------------------
template <int LPR>
static int launch_ws_kernel_lpr(lsk_ctx *ctx, int ndot, int grid, cudaStream_t st, const TmaSpmvArgs &a, RedScratch rs, double *o0, double *o1) {
    static const int family = configure_family_index();
    const int rc = configure_once(ctx, family, [] {
        cudaError_t e = cudaFuncSetAttribute(csr_ws_kernel<0, LPR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kWsSmem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(csr_ws_kernel<1, LPR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kWsSmem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(csr_ws_kernel<2, LPR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kWsSmem);
        return e;
    });
    if (rc != 0) return rc;
    cudaError_t e;
    if (ndot == 0) e = launch_pdl(kPdlSpmv, csr_ws_kernel<0, LPR>, grid, kWsThreads, kWsSmem, st, a, rs, o0, o1);
    else if (ndot == 1) e = launch_pdl(kPdlSpmv, csr_ws_kernel<1, LPR>, grid, kWsThreads, kWsSmem, st, a, rs, o0, o1);
    else e = launch_pdl(kPdlSpmv, csr_ws_kernel<2, LPR>, grid, kWsThreads, kWsSmem, st, a, rs, o0, o1);
    return (int) e;
}

static int ws_lanes_per_row(int64_t rows, int64_t nnz, int variant) {
    if (variant != LSK_SPMV_LANES) return 1;
    static const char *lpr_env = getenv("LSK_WS_LPR");  // developer knob: 2, 4 or 8 lanes per row for the LANES variant
    if (lpr_env) {
        const int v = atoi(lpr_env);
        if (v == 2 || v == 4 || v == 8) return v;
    }
    const double mean = rows > 0 ? (double) nnz / (double) rows : 1.0;
    return mean <= 16.0 ? 2 : mean <= 40.0 ? 4 : 8;
}

static int launch_ws_kernel(lsk_ctx *ctx, int lpr, int ndot, cudaStream_t st, int64_t rows, int64_t nnz, const double *entry,
                            const long long *col, const lsk_rect *rowptr, int64_t k_base, const double *x, double *y,
                            const double *dot_w, double *o0, double *o1, bool accumulate) {
    const int rpb = ws_rows_per_block(rows, nnz, lpr);
    const int64_t nrb = rows > 0 ? (rows + rpb - 1) / rpb : 0;  // no rows: one CTA that only finishes the fused dots (0)
    static const char *cta_env = getenv("LSK_WS_CTAS");  // developer knob: CTAs per SM
    const int64_t cap = (int64_t) ctx->sm_count * (cta_env ? atoi(cta_env) : LSK_WS_MINB);
    const int grid = (int) (nrb < 1 ? 1 : nrb < cap ? nrb : cap);
    TmaSpmvArgs a;
    a.rows = rows; a.nnz = nnz; a.rpb = rpb; a.n_row_blocks = nrb; a.entry = entry; a.col = col; a.rowptr = rowptr;
    a.k_base = k_base; a.x = x; a.y = y; a.dot_w = dot_w; a.accumulate = accumulate ? 1 : 0;
    // a reduction whose consumer is the next cg_update (lsk_ctx_defer_next_allreduce): only the y.w form
    const bool defer = take_defer(ctx) && ndot == 1 && o0 != nullptr;
    RedScratch rs = {nullptr, nullptr, nullptr, nullptr, 0};
    if (ndot > 0) rs = next_scratch(ctx);
    if (defer) {
        rs.defer = 1;
        ctx->pending_slot = o0;
    }
    switch (lpr) {
    case 1: return launch_ws_kernel_lpr<1>(ctx, ndot, grid, st, a, rs, o0, o1);
    case 2: return launch_ws_kernel_lpr<2>(ctx, ndot, grid, st, a, rs, o0, o1);
    case 4: return launch_ws_kernel_lpr<4>(ctx, ndot, grid, st, a, rs, o0, o1);
    default: return launch_ws_kernel_lpr<8>(ctx, ndot, grid, st, a, rs, o0, o1);
    }
}

template <typename T, int V>
static void launch_vector_kernel(int ndot, int grid, cudaStream_t st, int64_t rows, const T *entry,
                                 const long long *col, const lsk_rect *rowptr, int64_t k_base, const T *x,
                                 T *y, const T *dot_w, RedScratch rs, T *o0, T *o1, bool accumulate) {
    if (ndot == 0)
        csr_vector_kernel<T, V, 0><<<grid, kBlock, 0, st>>>(rows, entry, col, rowptr, k_base, x, y, dot_w, rs.partials, rs.ticket, o0, o1, rs.peers, accumulate);
    else if (ndot == 1)
        csr_vector_kernel<T, V, 1><<<grid, kBlock, 0, st>>>(rows, entry, col, rowptr, k_base, x, y, dot_w, rs.partials, rs.ticket, o0, o1, rs.peers, accumulate);
    else
        csr_vector_kernel<T, V, 2><<<grid, kBlock, 0, st>>>(rows, entry, col, rowptr, k_base, x, y, dot_w, rs.partials, rs.ticket, o0, o1, rs.peers, accumulate);
}

template <typename T>
static int csr_spmv(lsk_ctx *ctx, lsk_stream s, int64_t rows, int64_t nnz, const T *entry,
                    const int64_t *col, const lsk_rect *rowptr, int64_t k_base, const T *x_shifted, T *y,
                    const T *dot_w, T *dot_out, T *dot_yy_out, int variant) {
    if (!ctx || rows < 0 || nnz < 0) return LSK_E_INVALID;
    if (rows > 0 && (!rowptr || !y || !x_shifted)) return LSK_E_INVALID;
    if (nnz > 0 && (!entry || !col)) return LSK_E_INVALID;
    if ((dot_w == nullptr) != (dot_out == nullptr)) return LSK_E_INVALID;
    const bool accumulate = (variant & LSK_SPMV_ACCUMULATE) != 0;
    variant &= ~LSK_SPMV_ACCUMULATE;
    if (variant < LSK_SPMV_AUTO || variant > LSK_SPMV_LANES) return LSK_E_INVALID;
    // the fused reductions share one kernel shape: {} | {y.w} | {y.w, y.y}; y.y alone rides on a
    // y.w slot pointed at y itself (the kernels then use the row's own result for w)
    int ndot = 0;
    const T *w = dot_w;
    T *o0 = dot_out, *o1 = dot_yy_out;
    if (dot_out && dot_yy_out) ndot = 2;
    else if (dot_out) ndot = 1;
    else if (dot_yy_out) { ndot = 1; w = y; o0 = dot_yy_out; o1 = nullptr; }
    if (rows == 0 && ndot == 0) return 0;

    if (variant == LSK_SPMV_AUTO) variant = pick_variant(rows, nnz);
    const cudaStream_t st = (cudaStream_t) s;
    const long long *colp = reinterpret_cast<const long long *>(col);
    {
        const int rc = settle_pending(ctx, st);
        if (rc != 0) return rc;
    }

    // the TMA kernels need fp64 and col / entry 16-byte aligned at the same elements
    const bool tma_ok = std::is_same<T, double>::value && reinterpret_cast<uintptr_t>(entry) % 8 == 0 &&
                        reinterpret_cast<uintptr_t>(col) % 8 == 0 &&
                        (((reinterpret_cast<uintptr_t>(entry) >> 3) & 1) == ((reinterpret_cast<uintptr_t>(col) >> 3) & 1));
    static const char *impl_env = getenv("LSK_SPMV_IMPL");  // developer A/B switch: ws (default) | tma | pipe | regs
    const int impl = !impl_env ? 0 : (impl_env[0] == 't' ? 3 : impl_env[0] == 'p' ? 1 : impl_env[0] == 'r' ? 2 : 0);
    // ... and the warp-specialised one copies the rects with TMA as well: rowptr 16-byte aligned
    const bool ws_ok = tma_ok && reinterpret_cast<uintptr_t>(rowptr) % 16 == 0 && (variant == LSK_SPMV_STREAM || variant == LSK_SPMV_LANES);
    if (ws_ok && impl == 0) {
        // the warp-specialised TMA pipeline: thread per row (bit-exact) or 2-8 lanes per row
        const int rc = launch_ws_kernel(ctx, ws_lanes_per_row(rows, nnz, variant), ndot, st, rows, nnz, reinterpret_cast<const double *>(entry),
                                        colp, rowptr, k_base, reinterpret_cast<const double *>(x_shifted), reinterpret_cast<double *>(y),
                                        reinterpret_cast<const double *>(w), reinterpret_cast<double *>(o0), reinterpret_cast<double *>(o1), accumulate);
        if (rc != 0) return rc;
        return after_launch(ctx);
    }
    RedScratch rs = {nullptr, nullptr, nullptr, nullptr, 0};
    if (ndot > 0) rs = next_scratch(ctx);
    if (variant == LSK_SPMV_LANES) {
        if (tma_ok) {  // round-1 kernel (LSK_SPMV_IMPL=tma)
            const double mean = rows > 0 ? (double) nnz / (double) rows : 1.0;
            const int lpr = mean <= 16.0 ? 2 : mean <= 40.0 ? 4 : 8;
            const int rc = launch_tma_kernel(ctx, lpr, ndot, st, rows, nnz, reinterpret_cast<const double *>(entry), colp, rowptr, k_base,
                                             reinterpret_cast<const double *>(x_shifted), reinterpret_cast<double *>(y),
                                             reinterpret_cast<const double *>(w), rs, reinterpret_cast<double *>(o0),
                                             reinterpret_cast<double *>(o1), accumulate);
            if (rc != 0) return rc;
            return after_launch(ctx);
        }
        variant = LSK_SPMV_VECTOR;  // same idea without the TMA staging (fp32, odd alignments)
    }
    if (variant == LSK_SPMV_STREAM) {
        const double mean = rows > 0 ? (double) nnz / (double) rows : 1.0;
        int rpb = (int) ((double) kTile / (mean < 1.0 ? 1.0 : mean));
        rpb = (rpb / 32) * 32;
        if (rpb < 32) rpb = 32;
        if (rpb > kBlock) rpb = kBlock;
        const int64_t nrb = rows > 0 ? (rows + rpb - 1) / rpb : 1;
        static const char *cta_env = getenv("LSK_SPMV_CTAS");  // developer knob: CTAs per SM
        int64_t cap = (int64_t) ctx->sm_count * (cta_env ? atoi(cta_env) : 4);
        if (cap > kMaxPartials) cap = kMaxPartials;
        const int grid = (int) (nrb < cap ? nrb : cap);
        // 256-bit path needs entry and col to hit their vector alignment at the same elements
        const bool vec = (((reinterpret_cast<uintptr_t>(entry) / sizeof(T)) & 3) ==
                          ((reinterpret_cast<uintptr_t>(col) >> 3) & 3)) &&
                         (reinterpret_cast<uintptr_t>(entry) % sizeof(T) == 0) &&
                         (reinterpret_cast<uintptr_t>(col) % 8 == 0);
        if (tma_ok && impl == 3) {
            const int rc = launch_tma_kernel(ctx, 1, ndot, st, rows, nnz, reinterpret_cast<const double *>(entry), colp, rowptr, k_base,
                                             reinterpret_cast<const double *>(x_shifted), reinterpret_cast<double *>(y),
                                             reinterpret_cast<const double *>(w), rs, reinterpret_cast<double *>(o0),
                                             reinterpret_cast<double *>(o1), accumulate);
            if (rc != 0) return rc;
        } else if (vec && impl != 2)
            launch_pipe_kernel<T>(ndot, grid, st, rows, rpb, nrb, entry, colp, rowptr, k_base, x_shifted, y, w, rs, o0, o1, accumulate);
        else if (vec)
            launch_stream_kernel<T, true>(ndot, grid, st, rows, rpb, nrb, entry, colp, rowptr, k_base, x_shifted, y, w, rs, o0, o1, accumulate);
        else
            launch_stream_kernel<T, false>(ndot, grid, st, rows, rpb, nrb, entry, colp, rowptr, k_base, x_shifted, y, w, rs, o0, o1, accumulate);
    } else {
        int V = 32;
        if (variant == LSK_SPMV_VECTOR) {
            const double mean = rows > 0 ? (double) nnz / (double) rows : 1.0;
            V = mean <= 4.0 ? 2 : mean <= 8.0 ? 4 : mean <= 16.0 ? 8 : 16;
        }
        const int64_t passes = (rows * V + kBlock - 1) / kBlock;
        const int grid = stream_grid(ctx, (passes > 0 ? passes : 1) * kBlock, 8);
        switch (V) {
        case 2: launch_vector_kernel<T, 2>(ndot, grid, st, rows, entry, colp, rowptr, k_base, x_shifted, y, w, rs, o0, o1, accumulate); break;
        case 4: launch_vector_kernel<T, 4>(ndot, grid, st, rows, entry, colp, rowptr, k_base, x_shifted, y, w, rs, o0, o1, accumulate); break;
        case 8: launch_vector_kernel<T, 8>(ndot, grid, st, rows, entry, colp, rowptr, k_base, x_shifted, y, w, rs, o0, o1, accumulate); break;
        case 16: launch_vector_kernel<T, 16>(ndot, grid, st, rows, entry, colp, rowptr, k_base, x_shifted, y, w, rs, o0, o1, accumulate); break;
        default: launch_vector_kernel<T, 32>(ndot, grid, st, rows, entry, colp, rowptr, k_base, x_shifted, y, w, rs, o0, o1, accumulate); break;
        }
    }
    return after_launch(ctx);
}

template <typename T>
static int coo_spmv(lsk_ctx *ctx, lsk_stream s, int64_t nnz, const T *entry, const int64_t *row,
                    const int64_t *col, const T *x_shifted, T *y_shifted, int64_t row_lo, int64_t row_hi,
                    int64_t col_lo, int64_t col_hi) {
    if (!ctx || nnz < 0) return LSK_E_INVALID;
    if (nnz == 0) return 0;
    if (!entry || !row || !col || !x_shifted || !y_shifted) return LSK_E_INVALID;
    const int grid = stream_grid(ctx, (nnz + 3) / 4, 8);
    coo_segreduce_kernel<T><<<grid, kBlock, 0, (cudaStream_t) s>>>(
        nnz, entry, reinterpret_cast<const long long *>(row), reinterpret_cast<const long long *>(col),
        x_shifted, y_shifted, row_lo, row_hi, col_lo, col_hi);
    return after_launch(ctx);
}

}  // namespace lsk

using namespace lsk;

extern "C" {

int lsk_csr_spmv_pick(int64_t rows, int64_t nnz) { return pick_variant(rows, nnz); }

int lsk_csr_spmv_f64(lsk_ctx *ctx, lsk_stream s, int64_t rows, int64_t nnz, const double *entry,
                     const int64_t *col, const lsk_rect *rowptr, int64_t k_base, const double *x_shifted,
                     double *y, const double *dot_w, double *dot_out, double *dot_yy_out, int variant) {
    return csr_spmv<double>(ctx, s, rows, nnz, entry, col, rowptr, k_base, x_shifted, y, dot_w, dot_out,
                            dot_yy_out, variant);
}
int lsk_csr_spmv_f32(lsk_ctx *ctx, lsk_stream s, int64_t rows, int64_t nnz, const float *entry,
                     const int64_t *col, const lsk_rect *rowptr, int64_t k_base, const float *x_shifted,
                     float *y, const float *dot_w, float *dot_out, float *dot_yy_out, int variant) {
    return csr_spmv<float>(ctx, s, rows, nnz, entry, col, rowptr, k_base, x_shifted, y, dot_w, dot_out,
                           dot_yy_out, variant);
}
int lsk_coo_spmv_f64(lsk_ctx *ctx, lsk_stream s, int64_t nnz, const double *entry, const int64_t *row,
                     const int64_t *col, const double *x_shifted, double *y_shifted, int64_t row_lo,
                     int64_t row_hi, int64_t col_lo, int64_t col_hi) {
    return coo_spmv<double>(ctx, s, nnz, entry, row, col, x_shifted, y_shifted, row_lo, row_hi, col_lo, col_hi);
}
int lsk_coo_spmv_f32(lsk_ctx *ctx, lsk_stream s, int64_t nnz, const float *entry, const int64_t *row,
                     const int64_t *col, const float *x_shifted, float *y_shifted, int64_t row_lo,
                     int64_t row_hi, int64_t col_lo, int64_t col_hi) {
    return coo_spmv<float>(ctx, s, nnz, entry, row, col, x_shifted, y_shifted, row_lo, row_hi, col_lo, col_hi);
}

}  // extern "C"
