// lsk_spmv_tma.cuh -- the TMA-staged, thread-per-row CSR mat-vec as a device-side building block.
//
// Used by csr_tma_kernel (lsk_spmv.cu: one mat-vec per launch) -- round 1's single-role kernel, kept for pieces whose rowptr
// is not 16-byte aligned (the warp-specialised kernel copies the rects with TMA) and as LSK_SPMV_IMPL=tma for A/B runs.  See lsk_spmv.cu for the design notes: the
// CTA's contiguous run of (col, entry) is copied global->shared by the TMA engine
// (cp.async.bulk, mbarrier completion, L2 evict-first, two stages), threads own ROWS and add their
// rounded products in ascending k in a register -- bit-identical to the reference CPU body
// (src/CSRMatrixTasks.cpp:73-91).
#pragma once

#include <limits.h>

#include "lsk_common.cuh"

namespace lsk {

#ifndef LSK_TMA_TILE
#define LSK_TMA_TILE 2048
#endif
#ifndef LSK_TMA_MINB
#define LSK_TMA_MINB 3
#endif
constexpr int kTmaTile = LSK_TMA_TILE;  // non-zeros per TMA stage (col + entry: 16 B each)
constexpr int kTmaStages = 2;           // 2 x (16 KB col + 16 KB entry) = 64 KB dynamic shared memory
constexpr size_t kTmaSmem = (size_t) kTmaStages * kTmaTile * (sizeof(long long) + sizeof(double));

#ifdef __CUDACC__

__device__ __forceinline__ double load1_stream(const double *p) {
    return __longlong_as_double((long long) ld64_stream(p));
}
__device__ __forceinline__ float load1_stream(const float *p) { return __uint_as_float(ld32_stream(p)); }
__device__ __forceinline__ long long load1_stream(const long long *p) { return (long long) ld64_stream(p); }

__device__ __forceinline__ long long warp_min_ll(long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const long long t = __shfl_xor_sync(0xffffffffu, v, o);
        v = t < v ? t : v;
    }
    return v;
}
__device__ __forceinline__ long long warp_max_ll(long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const long long t = __shfl_xor_sync(0xffffffffu, v, o);
        v = t > v ? t : v;
    }
    return v;
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t) __cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
// try_wait sleeps in hardware between polls.  A wait that has not completed after kSpinLimitNs can only be a protocol
// bug or a lost copy: trap (the launch fails loudly) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    unsigned int polls = 0;
    unsigned long long t_start = 0;
    for (;;) {
        uint32_t done;
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (done) return;
        if (spin_expired(polls, t_start)) __trap();
    }
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar,
                                             uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
        ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}

// same, default L2 policy (vectors, which should stay L2-resident)
__device__ __forceinline__ void tma_bulk_g2s_plain(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// one rowptr rect (16 bytes), streamed once per mat-vec: no L1 allocation, evict-first in L2, so that the
// solver's vectors keep the L2
__device__ __forceinline__ longlong2 ld_rect_stream(const lsk_rect *p) {
    longlong2 r;
    uint64_t policy;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.b64 {%0, %1}, [%2], %3;"
                 : "=l"(r.x), "=l"(r.y)
                 : "l"(p), "l"(policy));
    return r;
}

struct TmaSpmvArgs {
    int64_t rows, nnz;
    int rpb;                 // rows per row block (<= kBlock, one thread per row)
    int64_t n_row_blocks;
    const double *entry;     // element 0 <-> k = k_base
    const long long *col;
    const lsk_rect *rowptr;  // inclusive rects of GLOBAL k
    int64_t k_base;
    const double *x;         // shifted to global column 0
    double *y;
    const double *dot_w;     // NDOT >= 1
    int accumulate;          // non-zero: y = y + A x (a further block of a multi-operator system on the same rows)
};

struct TmaSpmvState {
    uint64_t policy;   // thread 0 only
    uint32_t phases;   // bit s = parity to wait for on stage s (persists across runs)
};

// Call once per kernel, by all threads, before the first csr_tma_run (which starts with a CTA barrier).
__device__ __forceinline__ void csr_tma_init(TmaSpmvState &st, uint64_t *s_full) {
    st.policy = 0;
    st.phases = 0;
    if (threadIdx.x == 0) {
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(st.policy));
#pragma unroll
        for (int s = 0; s < kTmaStages; ++s) mbar_init(&s_full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
}

// One mat-vec over the row blocks blockIdx.x, blockIdx.x + gridDim.x, ...
//   NDOT      0: y only; 1: dacc[0] += y.w; 2: also dacc[1] += y.y
//   PART      0: the whole mat-vec.  1: only the prologue -- the rects of this CTA's first two row blocks are
//             loaded into `cur` and the TMA copy of its first tile is started (the matrix is constant, so a kernel
//             launched programmatically does this BEFORE it waits for its predecessor).  2: the rest.
struct TmaCursor {
    long long lo, hi1, nlo, nhi1;
};
//   LPR       lanes per row (1, 2, 4, 8).  1: a thread owns a row and adds its products in ascending k (bit-identical to
//             the reference CPU body).  > 1, for rows of a few dozen non-zeros where one thread per row would leave
//             most lanes of a tile idle: LPR adjacent lanes stride one row and their partial sums are combined with
//             shuffles (fixed tree order: deterministic, <= 1e-12 from the sequential sum).
template <int NDOT, int PART = 0, int LPR = 1>
__device__ __forceinline__ void csr_tma_run(const TmaSpmvArgs &a, TmaSpmvState &st, unsigned char *s_dyn, uint64_t *s_full,
                                            long long (*s_lo)[kWarps], long long (*s_hi)[kWarps],
                                            double (&dacc)[NDOT > 0 ? NDOT : 1], TmaCursor *cur = nullptr) {
    constexpr int S = kTmaStages;
    long long (*s_col)[kTmaTile] = reinterpret_cast<long long (*)[kTmaTile]>(s_dyn);
    double (*s_ent)[kTmaTile] = reinterpret_cast<double (*)[kTmaTile]>(s_dyn + (size_t) S * kTmaTile * sizeof(long long));
    const int tid = threadIdx.x;
    // row (within the row block) and sub-lane of this thread.  With LPR lanes per row a warp covers RPW = 32 / LPR rows
    // and the lanes of one row sit RPW apart: lanes 0 .. RPW-1 read element s of RPW CONSECUTIVE rows -- the same stencil
    // leg, one or two cache lines -- exactly like the thread-per-row mapping, instead of LPR scattered runs.
    constexpr int RPW = 32 / LPR;
    const int trow = (tid >> 5) * RPW + (tid & 31) % RPW;
    const int tsub = (tid & 31) / RPW;
    const int64_t G = gridDim.x;
    const int64_t rows = a.rows, n_row_blocks = a.n_row_blocks;
    const int rpb = a.rpb;
    const double *__restrict__ entry = a.entry;
    const long long *__restrict__ col = a.col;
    const lsk_rect *__restrict__ rowptr = a.rowptr;
    const int64_t k_base = a.k_base;
    const double *x = a.x;
    const uint64_t policy = st.policy;

    auto load_rect = [&](int64_t rb, long long &lo, long long &hi1) {
        lo = LLONG_MAX;
        hi1 = LLONG_MIN;
        if (rb < n_row_blocks) {
            const int64_t r = rb * rpb + trow;
            if (trow < rpb && r < rows) {
                const longlong2 rc = ld_rect_stream(rowptr + r);
                if (rc.y >= rc.x) {
                    lo = rc.x - k_base;
                    hi1 = rc.y + 1 - k_base;
                }
            }
        }
    };
    auto publish_span = [&](int pb, long long lo, long long hi1) {
        const long long wl = warp_min_ll(lo), wh = warp_max_ll(hi1);
        if ((tid & 31) == 0) {
            s_lo[pb][tid >> 5] = wl;
            s_hi[pb][tid >> 5] = wh;
        }
    };
    auto read_span = [&](int pb, long long &jb, long long &je) {
        jb = s_lo[pb][0];
        je = s_hi[pb][0];
#pragma unroll
        for (int w = 1; w < kWarps; ++w) {
            jb = s_lo[pb][w] < jb ? s_lo[pb][w] : jb;
            je = s_hi[pb][w] > je ? s_hi[pb][w] : je;
        }
        if (je <= jb) jb = je = 0;
    };
    // tiles start where `col` (and the congruent `entry`) are 16-byte aligned
    auto tile_start = [&](long long jb) { return jb - (long long) ((reinterpret_cast<uintptr_t>(col + jb) >> 3) & 1); };
    // thread 0: copy elements [t0, t0 + kTmaTile) /\ [jb, je) of both arrays into stage s
    auto issue_tile = [&](int s, long long t0, long long jb, long long je) {
        long long a0 = t0 > jb ? t0 : jb;                             // first needed element
        long long b0 = (t0 + kTmaTile) < je ? (t0 + kTmaTile) : je;   // one past the last
        if (b0 < a0) b0 = a0;
        // bulk part: 16-byte aligned on both ends, never outside [0, nnz)
        long long A = a0 + ((reinterpret_cast<uintptr_t>(col + a0) >> 3) & 1);
        long long B = b0 - ((reinterpret_cast<uintptr_t>(col + b0) >> 3) & 1);
        if (B < A) B = A;
        const uint32_t bytes = (uint32_t) (B - A) * 8u;
        mbar_expect_tx(&s_full[s], 2u * bytes);
        if (bytes) {
            tma_bulk_g2s(&s_col[s][A - t0], col + A, bytes, &s_full[s], policy);
            tma_bulk_g2s(&s_ent[s][A - t0], entry + A, bytes, &s_full[s], policy);
        }
        // ragged single elements at either end (generic proxy; visible after the next CTA barrier)
        if (a0 < A && a0 < b0) {
            s_col[s][a0 - t0] = load1_stream(col + a0);
            s_ent[s][a0 - t0] = load1_stream(entry + a0);
        }
        if (B < b0 && B >= A && !(a0 < A && B == a0)) {
            s_col[s][B - t0] = load1_stream(col + B);
            s_ent[s][B - t0] = load1_stream(entry + B);
        }
    };

    uint32_t phases = st.phases;
    // This CTA's row blocks are blockIdx.x + kk * G, kk = 0, 1, ...
    auto next_rb = [&](int64_t r) -> int64_t { return r + G; };  // successor; >= n_row_blocks = none
    int64_t rb = (int64_t) blockIdx.x < n_row_blocks ? (int64_t) blockIdx.x : n_row_blocks;
    int64_t rb_next = rb < n_row_blocks ? next_rb(rb) : n_row_blocks;
    if (rb_next > n_row_blocks) rb_next = n_row_blocks;
    long long lo, hi1, jb, je, nlo, nhi1;
    long long t0;
    if constexpr (PART != 2) {
        load_rect(rb, lo, hi1);
        publish_span(0, lo, hi1);
        __syncthreads();  // also publishes the mbarrier inits (first run) / retires the previous run's tiles
        read_span(0, jb, je);
        t0 = tile_start(jb);
        if (tid == 0 && rb < n_row_blocks) issue_tile(0, t0, jb, je);
        load_rect(rb_next, nlo, nhi1);
        if constexpr (PART == 1) {
            cur->lo = lo; cur->hi1 = hi1; cur->nlo = nlo; cur->nhi1 = nhi1;
            return;
        }
    } else {
        lo = cur->lo; hi1 = cur->hi1; nlo = cur->nlo; nhi1 = cur->nhi1;
        read_span(0, jb, je);  // still published: nothing has touched s_lo / s_hi since the prologue
        t0 = tile_start(jb);
    }
    int stage = 0, pb = 1;
    double acc = 0.0;

    while (rb < n_row_blocks) {
        const bool last_tile = (t0 + kTmaTile >= je);
        if (last_tile) publish_span(pb, nlo, nhi1);
        // one barrier per tile: (i) stage^1, consumed last iteration, may now be overwritten;
        // (ii) the next block's span is published; (iii) ragged elements stored by thread 0 are visible
        __syncthreads();
        long long njb = 0, nje = 0;
        if (last_tile) {
            read_span(pb, njb, nje);
            pb ^= 1;
        }
        if (tid == 0) {
            if (!last_tile) issue_tile(stage ^ 1, t0 + kTmaTile, jb, je);
            else if (rb_next < n_row_blocks) issue_tile(stage ^ 1, tile_start(njb), njb, nje);
        }
        // ---- consume this tile: thread-per-row, products added in ascending k
        // the fused dot's w[r] is requested now, so that its latency hides behind the gathers below
        double wv = 0.0;
        if constexpr (NDOT >= 1) {
            const int64_t r = rb * rpb + trow;
            if (last_tile && tsub == 0 && trow < rpb && r < rows && a.dot_w != a.y) wv = __ldg(a.dot_w + r);
        }
        mbar_wait(&s_full[stage], (phases >> stage) & 1u);
        phases ^= (1u << stage);
        {
            const long long ka = lo > t0 ? lo : t0;
            const long long kb = hi1 < t0 + kTmaTile ? hi1 : t0 + kTmaTile;
            const long long *sc = s_col[stage];
            const double *se = s_ent[stage];
            // up to kChunk gathers in flight per thread; the adds stay in ascending k
            constexpr int kChunk = 8;
            long long j = ka < kb ? ka : kb;  // (rows without elements in this tile carry sentinel bounds)
            if constexpr (LPR > 1) {
                // LPR lanes stride the row: lane `sub` takes elements ka + sub, ka + sub + LPR, ...
                for (j += tsub; j < kb; j += (long long) kChunk * LPR) {
                    const int o = (int) (j - t0);
                    double xv[kChunk];
#pragma unroll
                    for (int e = 0; e < kChunk; ++e)
                        xv[e] = (j + e * LPR < kb) ? __ldg(x + sc[o + e * LPR]) : 0.0;
#pragma unroll
                    for (int e = 0; e < kChunk; ++e)
                        if (j + e * LPR < kb) acc = add_rn(acc, mul_rn(se[o + e * LPR], xv[e]));
                }
                j = kb;  // nothing left for the one-lane loops below
            }
            for (; j + kChunk <= kb; j += kChunk) {  // full chunks: no predication
                const int o = (int) (j - t0);
                double xv[kChunk];
#pragma unroll
                for (int e = 0; e < kChunk; ++e) xv[e] = __ldg(x + sc[o + e]);
#pragma unroll
                for (int e = 0; e < kChunk; ++e) acc = add_rn(acc, mul_rn(se[o + e], xv[e]));
            }
            if (j < kb) {  // 1..kChunk-1 left
                const int o = (int) (j - t0);
                const int rem = (int) (kb - j);
                double xv[kChunk - 1];
#pragma unroll
                for (int e = 0; e < kChunk - 1; ++e)
                    xv[e] = (e < rem) ? __ldg(x + sc[o + e]) : 0.0;
#pragma unroll
                for (int e = 0; e < kChunk - 1; ++e)
                    if (e < rem) acc = add_rn(acc, mul_rn(se[o + e], xv[e]));
            }
        }
        if (last_tile) {
            if constexpr (LPR > 1) {
#pragma unroll
                for (int o = 16; o >= RPW; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            }
            const int64_t r = rb * rpb + trow;
            if (tsub == 0 && trow < rpb && r < rows) {
                if (a.accumulate) acc = add_rn(a.y[r], acc);
                a.y[r] = acc;
                if constexpr (NDOT >= 1) dacc[0] = fma(acc, a.dot_w != a.y ? wv : acc, dacc[0]);  // w == y: the y.y-only form
                if constexpr (NDOT >= 2) dacc[NDOT - 1] = fma(acc, acc, dacc[NDOT - 1]);
            }
            acc = 0.0;
            rb = rb_next;
            rb_next = rb < n_row_blocks ? next_rb(rb) : n_row_blocks;
            if (rb_next > n_row_blocks) rb_next = n_row_blocks;
            lo = nlo;
            hi1 = nhi1;
            jb = njb;
            je = nje;
            t0 = tile_start(jb);
            load_rect(rb_next, nlo, nhi1);
        } else {
            t0 += kTmaTile;
        }
        stage ^= 1;
    }
    st.phases = phases;
    // stage parity: a run may end after an odd number of tiles; the next run starts at stage 0 again, which
    // is safe because every issued tile has been waited for (no copy in flight) and `phases` is per stage
}

// rows per row block so that one block's non-zeros fill about one tile
inline int tma_rows_per_block(int64_t rows, int64_t nnz) {
    const double mean = rows > 0 ? (double) nnz / (double) rows : 1.0;
    int rpb = (int) ((double) kTmaTile / (mean < 1.0 ? 1.0 : mean));
    rpb = (rpb / 32) * 32;
    if (rpb < 32) rpb = 32;
    if (rpb > kBlock) rpb = kBlock;
    return rpb;
}

#endif  // __CUDACC__

}  // namespace lsk
