// lsk_spmv_ws.cuh -- the warp-specialised, TMA-pipelined CSR mat-vec (fp64; the default CSR kernel).
//
// What the round-1 kernel (lsk_spmv_tma.cuh) left on the table, per its own ncu source view: the CTA that consumes a
// tile is also the one that requests the next, one CTA barrier per tile, so HBM requests are only in flight while the
// consumers compute; and the rowptr rect of the next row block is needed the moment it has been requested (20 % of all
// stall samples sat on that one compare).  This kernel splits the two roles:
//
//   producer warp   walks this CTA's row blocks AHEAD of the consumers and has the TMA engine copy, per row block, the
//                   block's rowptr rects and the (col, entry) tile(s) of its run of non-zeros global->shared
//                   (cp.async.bulk, mbarrier transaction count, L2 evict-first).  It only ever blocks on an EMPTY
//                   barrier, i.e. when all kWsStages stages are full: the memory pipeline is kWsStages deep and never
//                   drains while there is work.
//   consumer warps  (8 x 32 threads) wait on a stage's FULL barrier, read their row's rect from the stage, gather x,
//                   and add the rounded products in ascending k in a register -- thread-per-row, bit-identical to the
//                   reference CPU body (src/CSRMatrixTasks.cpp:73-91) -- or LPR lanes per row with a fixed shuffle
//                   tree.  Each warp releases the stage with one mbarrier arrive; there is no CTA-wide barrier in the
//                   steady state.
//
// The producer must stay cheap: measured with idle consumers, ONE warp that loads, normalises and min/max-reduces all
// 256 rects of a 7-point row block (64-bit arithmetic, ~600 dependent instructions) sustains 0.56 of the HBM peak per SM,
// two of them 1.0 -- and lose issue slots to the consumers.  So the producer does not look at every rect.  It SAMPLES
// the block's first and last 16 rows (one rect per lane, requested three blocks ahead) and takes the bounding interval
// of their rects as the block's run [jb, je) -- exact for any matrix whose rows are stored in row order, i.e. every CSR
// matrix in practice.  The reference's format allows arbitrary rects, so each consumer thread checks its own row
// against [jb, je); a row that is not covered (never, for ordered storage) is summed straight from global memory in
// the same ascending-k order.  The result is identical either way; only speed depends on the ordering.
//
// Tiles over-fetch to 16-byte boundaries inside [0, nnz) instead of patching single elements with scalar loads (the
// neighbours' elements land in shared memory and are never read); scalar patching remains only at the two ends of the
// piece.  On several GPUs the ghost values of x are in place when the kernel starts (the kernel that produced x
// exchanged its halo itself, lsk_blas1.cu), so the mat-vec knows nothing about ranks.  Programmatic dependent launch:
// the producer needs nothing from the preceding kernel (the matrix is constant), so the pipeline fills while the
// predecessor drains; consumers wait.
#pragma once

#include <limits.h>

#include "lsk_common.cuh"
#include "lsk_spmv_tma.cuh"

namespace lsk {

// Measured on B200 (fused SpMV + dot, fraction of the 6.54 TB/s copy peak; 7-pt 256^3 / 5-pt 8192^2 / 27-pt 192^3 / the
// 2 M-row slab of an 8-GPU run): 2 stages x 3 CTAs per SM 1.05 / 1.03 / 0.93 / 0.93; 3 stages x 2 CTAs 1.00 / 0.89 / 0.86 /
// 0.86.  More CTAs (24 consumer warps, three producers per SM) beat a deeper ring.
#ifndef LSK_WS_STAGES
#define LSK_WS_STAGES 2
#endif
#ifndef LSK_WS_MINB
#define LSK_WS_MINB 3
#endif
#ifndef LSK_WS_DEBUG_MODE
#define LSK_WS_DEBUG_MODE 0  // developer builds: 1 = consumers release every stage untouched (the producer's ceiling)
#endif
constexpr int kWsTile = 2048;                 // non-zeros per stage (16 KB col + 16 KB entry)
constexpr int kWsStages = LSK_WS_STAGES;
constexpr int kWsConsumerWarps = kBlock / 32;  // 8
constexpr int kWsThreads = kBlock + 32;        // consumers + the producer warp
constexpr int kWsMaxRows = kBlock;             // rows per row block (thread per row)
constexpr size_t kWsStageBytes = (size_t) kWsTile * 16 + (size_t) kWsMaxRows * 16;
constexpr size_t kWsSmem = (size_t) kWsStages * kWsStageBytes;
constexpr int kWsMaxSpanTiles = 64;            // longest run of one row block that is streamed through the tiles
constexpr int kWsAhead = 3;                    // row blocks between the request of a block's sample rects and their use

#ifdef __CUDACC__

struct WsMeta {
    long long rb;      // row block held by the stage; < 0: no more work
    long long t0;      // element index (piece-local) of shared-memory slot 0 of the tile
    long long jb, je;  // the block's run of non-zeros as the producer sees it (piece-local)
    int first;         // first tile of the row block: the stage carries the block's rects
    int last;          // last tile of the row block: rows are finished
};

// 64-bit warp min / max with redux.sync (one instruction per 32-bit half instead of five shuffle steps): first the
// signed high words, then the unsigned low words of the lanes that hold the winning high word
__device__ __forceinline__ long long warp_min_ll_redux(long long v) {
    const int hi = (int) (v >> 32);
    const int mhi = __reduce_min_sync(0xffffffffu, hi);
    const unsigned lo = hi == mhi ? (unsigned) v : 0xffffffffu;
    const unsigned mlo = __reduce_min_sync(0xffffffffu, lo);
    return ((long long) mhi << 32) | (long long) mlo;
}
__device__ __forceinline__ long long warp_max_ll_redux(long long v) {
    const int hi = (int) (v >> 32);
    const int mhi = __reduce_max_sync(0xffffffffu, hi);
    const unsigned lo = hi == mhi ? (unsigned) v : 0u;
    const unsigned mlo = __reduce_max_sync(0xffffffffu, lo);
    return ((long long) mhi << 32) | (long long) mlo;
}

__device__ __forceinline__ void mbar_arrive_cta(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ longlong2 ld_rect_policy(const lsk_rect *p, uint64_t policy) {
    longlong2 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.b64 {%0, %1}, [%2], %3;" : "=l"(r.x), "=l"(r.y) : "l"(p), "l"(policy));
    return r;
}

//   NDOT   0: y only; 1: + y.w; 2: + y.w and y.y
//   LPR    lanes per row (1 = thread per row, bit-exact; 2, 4, 8 = tree-combined partial sums)
template <int NDOT, int LPR>
__global__ void __launch_bounds__(kWsThreads, LSK_WS_MINB)
csr_ws_kernel(TmaSpmvArgs a, RedScratch rs, double *out_yw, double *out_yy) {
    constexpr int S = kWsStages;
    constexpr int RPW = 32 / LPR;  // rows per consumer warp
    extern __shared__ __align__(128) unsigned char s_dyn[];
    __shared__ __align__(8) uint64_t s_full[S], s_empty[S];
    __shared__ __align__(16) WsMeta s_meta[S];
    long long (*s_col)[kWsTile] = reinterpret_cast<long long (*)[kWsTile]>(s_dyn);
    double (*s_ent)[kWsTile] = reinterpret_cast<double (*)[kWsTile]>(s_dyn + (size_t) S * kWsTile * 8);
    longlong2 (*s_rect)[kWsMaxRows] = reinterpret_cast<longlong2 (*)[kWsMaxRows]>(s_dyn + (size_t) S * kWsTile * 16);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < S; ++s) {
            mbar_init(&s_full[s], 1);
            mbar_init(&s_empty[s], kWsConsumerWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    pdl_launch_dependents();

    const int64_t rows = a.rows, nrb = a.n_row_blocks;
    const int rpb = a.rpb;
    const int64_t k_base = a.k_base;
    double dacc[NDOT > 0 ? NDOT : 1];
#pragma unroll
    for (int j = 0; j < (NDOT > 0 ? NDOT : 1); ++j) dacc[j] = 0.0;

    if (warp == kWsConsumerWarps) {
        // =========================================== producer ===========================================
        const long long *__restrict__ col = a.col;
        const double *__restrict__ entry = a.entry;
        const lsk_rect *__restrict__ rowptr = a.rowptr;
        const int64_t nnz = a.nnz;
        const int64_t G = gridDim.x;
        // (keeping part of a small slab's matrix in the L2 between solver iterations -- evict-normal instead of evict-first
        // for the first 20-90 MB of an 8-GPU slab -- was measured: 70.6 -> 71.5-77.4 us per CG iteration; the L2 is better
        // spent on the solver's vectors)
        uint64_t policy;
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
        auto par = [&](long long j) { return (long long) ((reinterpret_cast<uintptr_t>(col + j) >> 3) & 1); };

        // ---- the walk: position w = 0, 1, ... of this CTA -> row block blockIdx.x + w * G (nrb = none left)
        const int64_t mine = nrb > (int64_t) blockIdx.x ? (nrb - 1 - blockIdx.x) / G + 1 : 0;
        auto block_at = [&](int64_t k) -> int64_t { return k >= mine ? nrb : (int64_t) blockIdx.x + k * G; };
        // sample rect of this lane for row block rb: lanes 0-15 the block's first rows, lanes 16-31 its last ones
        auto sample = [&](int64_t rb) -> longlong2 {
            longlong2 rc = make_longlong2(0, -1);
            if (rb < nrb) {
                const int64_t r0 = rb * rpb;
                const int64_t left = rows - r0;
                const int nr = (int) (left < rpb ? left : rpb);
                int idx = lane < 16 ? lane : nr - 32 + lane;
                idx = idx < 0 ? 0 : (idx >= nr ? nr - 1 : idx);
                if (nr > 0) rc = ld_rect_policy(rowptr + r0 + idx, policy);
            }
            return rc;
        };
        longlong2 q[kWsAhead];  // samples of the blocks at positions w, w + 1, ..
#pragma unroll
        for (int i = 0; i < kWsAhead; ++i) q[i] = sample(block_at(i));
        int64_t w = 0;
        int64_t rb = block_at(0);
        int stage = 0;
        uint32_t empty_phase = (1u << S) - 1u;  // a fresh barrier passes a wait on the preceding phase

        while (rb < nrb) {
            const longlong2 rc = q[0];
#pragma unroll
            for (int i = 0; i + 1 < kWsAhead; ++i) q[i] = q[i + 1];
            q[kWsAhead - 1] = sample(block_at(w + kWsAhead));  // in flight for kWsAhead row blocks
            // ---- the block's run of non-zeros: bounding interval of the sampled rects
            long long mn = LLONG_MAX, mx = LLONG_MIN;
            if (rc.y >= rc.x) {
                mn = rc.x - k_base;
                mx = rc.y + 1 - k_base;
            }
            long long jb = warp_min_ll_redux(mn), je = warp_max_ll_redux(mx);
            if (je <= jb) jb = je = 0;
            if (jb < 0) jb = 0;      // (a rect that reaches outside the piece is a contract violation; stay in bounds)
            if (je > nnz) je = nnz;
            if (je < jb) je = jb;
            // rows stored in scattered order can make the sampled interval arbitrarily long: beyond kWsMaxSpanTiles the
            // block gets an empty run and its rows take the direct path (bounded work for any input)
            if (je - jb > (long long) kWsMaxSpanTiles * kWsTile) jb = je = 0;
            const int64_t r0 = rb * rpb;
            const int64_t left = rows - r0;
            const uint32_t rect_bytes = (uint32_t) (left < rpb ? left : rpb) * 16u;
            long long t0 = jb - par(jb);
            bool first = true;
            do {
                mbar_wait(&s_empty[stage], (empty_phase >> stage) & 1u);
                empty_phase ^= (1u << stage);
                if (lane == 0) {
                    const bool last = (t0 + kWsTile >= je);
                    WsMeta m;
                    m.rb = rb; m.t0 = t0; m.jb = jb; m.je = je; m.first = first ? 1 : 0; m.last = last ? 1 : 0;
                    s_meta[stage] = m;
                    const long long a0 = t0 > jb ? t0 : jb;
                    long long b0 = (t0 + kWsTile) < je ? (t0 + kWsTile) : je;
                    if (b0 < a0) b0 = a0;
                    // bulk part: the 16-byte aligned superset of [a0, b0) where it stays inside [0, nnz), else subset
                    long long A = a0 - par(a0), B = b0 + par(b0);
                    const bool head = A < 0, tail = B > nnz;
                    if (head) A = a0 + 1;
                    if (tail) B = b0 - 1;
                    if (B < A) B = A;
                    const uint32_t bytes = (uint32_t) (B - A) * 8u;
                    if (head && a0 < b0) {  // single elements at the two ends of the piece (generic proxy)
                        s_col[stage][a0 - t0] = load1_stream(col + a0);
                        s_ent[stage][a0 - t0] = load1_stream(entry + a0);
                    }
                    if (tail && a0 < b0) {
                        s_col[stage][b0 - 1 - t0] = load1_stream(col + b0 - 1);
                        s_ent[stage][b0 - 1 - t0] = load1_stream(entry + b0 - 1);
                    }
                    // the arrive that completes the phase once the bytes have landed
                    mbar_expect_tx(&s_full[stage], 2u * bytes + (first ? rect_bytes : 0u));
                    if (first) tma_bulk_g2s(&s_rect[stage][0], rowptr + r0, rect_bytes, &s_full[stage], policy);
                    if (bytes) {
                        tma_bulk_g2s(&s_col[stage][A - t0], col + A, bytes, &s_full[stage], policy);
                        tma_bulk_g2s(&s_ent[stage][A - t0], entry + A, bytes, &s_full[stage], policy);
                    }
                }
                stage = stage + 1 == S ? 0 : stage + 1;
                t0 += kWsTile;
                first = false;
            } while (t0 < je);
            ++w;
            rb = block_at(w);
        }
        // terminator stage
        mbar_wait(&s_empty[stage], (empty_phase >> stage) & 1u);
        if (lane == 0) {
            WsMeta m;
            m.rb = -1; m.t0 = 0; m.jb = 0; m.je = 0; m.first = 0; m.last = 0;
            s_meta[stage] = m;
            mbar_arrive_cta(&s_full[stage]);
        }
    } else {
        // =========================================== consumers ===========================================
        const double *x = a.x;
        const int trow = warp * RPW + lane % RPW;   // row within the row block
        const int tsub = lane / RPW;                // lane of the row (LPR > 1)
        pdl_wait();  // x, w and y belong to the kernels before this one
        int stage = 0;
        uint32_t full_phase = 0;
        long long lo = 0, hi1 = 0;   // this row's run (piece-local), [0, 0) if it has none
        bool direct = false;         // the row is not covered by the producer's [jb, je): summed from global memory
        double acc = 0.0, wv = 0.0;
        for (;;) {
            mbar_wait(&s_full[stage], (full_phase >> stage) & 1u);
            full_phase ^= (1u << stage);
            const long long rb = s_meta[stage].rb;
            if (rb < 0) break;
#if LSK_WS_DEBUG_MODE == 1
            __syncwarp();
            if (lane == 0) mbar_arrive_cta(&s_empty[stage]);
            stage = stage + 1 == S ? 0 : stage + 1;
            continue;
#endif
            const long long t0 = s_meta[stage].t0;
            const bool last = s_meta[stage].last != 0;
            const int64_t r = rb * rpb + trow;
            const bool have = trow < rpb && r < rows;
            if (s_meta[stage].first != 0) {
                const longlong2 rc = s_rect[stage][trow];  // the row's rect as stored: inclusive, global k
                lo = hi1 = 0;
                direct = false;
                if (have && rc.y >= rc.x) {
                    lo = rc.x - k_base;
                    hi1 = rc.y + 1 - k_base;
                    direct = lo < s_meta[stage].jb || hi1 > s_meta[stage].je;
                }
                // beta = 0; or the row's current value when a further block of a multi-operator system accumulates onto the
                // same rows: the products are then added one by one onto it, exactly as the reference's CPU body adds them
                // onto the zero-filled (or already partly summed) destination
                acc = (a.accumulate && tsub == 0 && have) ? a.y[r] : 0.0;
                if constexpr (NDOT >= 1) {  // requested now: its latency hides behind the gathers
                    // (w == y: the y.y-only form of the C ABI; the row's own result is used instead, below)
                    if (tsub == 0 && have && a.dot_w != a.y) wv = __ldg(a.dot_w + r);
                }
            }
            if (!direct) {
                const long long ka_l = lo > t0 ? lo : t0;
                const long long kb_l = hi1 < t0 + kWsTile ? hi1 : t0 + kWsTile;
                int j = 0, kb = 0;  // this row's elements inside the tile, as shared-memory slots [j, kb)
                if (ka_l < kb_l) {
                    j = (int) (ka_l - t0);
                    kb = (int) (kb_l - t0);
                }
                const long long *sc = s_col[stage];
                const double *se = s_ent[stage];
                constexpr int kChunk = 8;  // gathers in flight per thread; the adds stay in ascending k
                if constexpr (LPR > 1) {
                    for (j += tsub; j < kb; j += kChunk * LPR) {
                        double xv[kChunk];
#pragma unroll
                        for (int e = 0; e < kChunk; ++e)
                            xv[e] = (j + e * LPR < kb) ? __ldg(x + sc[j + e * LPR]) : 0.0;
#pragma unroll
                        for (int e = 0; e < kChunk; ++e)
                            if (j + e * LPR < kb) acc = add_rn(acc, mul_rn(se[j + e * LPR], xv[e]));
                    }
                } else {
                    for (; j + kChunk <= kb; j += kChunk) {
                        double xv[kChunk];
#pragma unroll
                        for (int e = 0; e < kChunk; ++e) xv[e] = __ldg(x + sc[j + e]);
#pragma unroll
                        for (int e = 0; e < kChunk; ++e) acc = add_rn(acc, mul_rn(se[j + e], xv[e]));
                    }
                    if (j < kb) {
                        const int rem = kb - j;
                        double xv[kChunk - 1];
#pragma unroll
                        for (int e = 0; e < kChunk - 1; ++e)
                            xv[e] = (e < rem) ? (__ldg(x + sc[j + e])) : 0.0;
#pragma unroll
                        for (int e = 0; e < kChunk - 1; ++e)
                            if (e < rem) acc = add_rn(acc, mul_rn(se[j + e], xv[e]));
                    }
                }
            }
            // this warp is done with the stage (its rects were copied to registers above)
            __syncwarp();
            if (lane == 0) mbar_arrive_cta(&s_empty[stage]);
            if (last) {
                if (direct && tsub == 0) {
                    // rows stored out of order: the tiles did not cover this one.  Same sum, same order, from global memory.
                    for (long long k = lo; k < hi1; ++k) {
                        const long long c = load1_stream(a.col + k);
                        acc = add_rn(acc, mul_rn(load1_stream(a.entry + k), __ldg(x + c)));
                    }
                }
                if constexpr (LPR > 1) {
#pragma unroll
                    for (int o = 16; o >= RPW; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                }
                if (tsub == 0 && have) {
                    a.y[r] = acc;
                    if constexpr (NDOT >= 1) dacc[0] = fma(acc, a.dot_w != a.y ? wv : acc, dacc[0]);
                    if constexpr (NDOT >= 2) dacc[NDOT - 1] = fma(acc, acc, dacc[NDOT - 1]);
                }
            }
            stage = stage + 1 == S ? 0 : stage + 1;
        }
    }
    if constexpr (NDOT > 0) {
        double *out[NDOT];
        out[0] = out_yw;
        if constexpr (NDOT >= 2) out[NDOT - 1] = out_yy;
        grid_reduce_finish<NDOT, double, kWsThreads>(dacc, rs.partials, rs.ticket, out, rs.peers, nullptr, rs.defer != 0);
    }
}

// rows per row block: the block's non-zeros should fill about one tile; at most one consumer thread group per row
inline int ws_rows_per_block(int64_t rows, int64_t nnz, int lpr) {
    const double mean = rows > 0 ? (double) nnz / (double) rows : 1.0;
    const int unit = 32 / lpr;  // rows per warp
    int rpb = (int) ((double) kWsTile / (mean < 1.0 ? 1.0 : mean));
    rpb = (rpb / unit) * unit;
    if (rpb < unit) rpb = unit;
    if (rpb > kWsMaxRows / lpr) rpb = kWsMaxRows / lpr;
    return rpb;
}

#endif  // __CUDACC__

}  // namespace lsk
