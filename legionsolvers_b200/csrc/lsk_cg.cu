// lsk_cg.cu -- the whole CG step (src/CGSolver.hpp:46-55) as ONE persistent kernel.
//
// The three HBM passes of the fused CG iteration are launch- and latency-bound once the problem is
// strong-scaled over 8 GPUs (a 256^3 slab is ~80 us of memory traffic per iteration, but three kernel
// boundaries, two dependent all-reduces and a halo exchange cost ~40 us more).  This kernel keeps one
// co-resident wave of CTAs alive for `niter` complete iterations:
//
//   phase A   q = A p          (the TMA-staged thread-per-row mat-vec of lsk_spmv_tma.cuh) + p.q partial
//   sync 1    grid barrier (gather-broadcast through CTA 0 with LL packets); CTA 0 sums the partials in a
//             fixed order and, on several GPUs, exchanges the rank sums over NVLink peer memory, then
//             releases the grid
//   phase B   x = fma(rr/pq, p, x);  r = fma((-1*rr)/pq, q, r);  r.r partial      (:50-52)
//   sync 2    same as sync 1 for r.r
//   phase C   p = fma(rr_new/rr, p, r)                                             (:54)
//             boundary elements are also stored into the neighbours' ghost regions
//   sync 3    grid barrier; CTA 0 publishes "my halo of this iteration has landed" to the neighbours.
//             Nobody waits here: in the next phase A a CTA waits for the neighbours' flag only before it
//             consumes a row block that references ghost columns (lsk_spmv_tma.cuh, GhostGate) -- and those
//             blocks are walked mid-phase, when the halo has long arrived.
//
// Element-wise arithmetic is the reference's (same fma / rounded multiply per element as the
// 3-kernel fused path and the oracle); scalars never leave the device; the residual history is
// appended by the kernel.  Matrix tiles come through the async proxy (TMA) and are read-only.  The
// vectors are written with generic-proxy stores and re-read -- by the mat-vec's gathers on the coherent
// path, by the vector phases through cp.async.bulk (async proxy) -- after a grid barrier: every thread
// issues fence.proxy.async before it arrives, thread 0 fences at gpu scope around the barrier's packets.
#include <stdlib.h>

#include "lsk_common.cuh"
#include "lsk_spmv_tma.cuh"
#include "lsk_vec_stream.cuh"

namespace lsk {

struct GridSync {
    unsigned int gen;    // generation of the last completed barrier (read at kernel start, written back at its end)
    int error;
    unsigned long long release[2];   // root -> everybody: LL packets {32 data bits | generation << 32}
    unsigned long long work[4];      // dynamic work counters of the phases (A, B, C); reset by CTA 0 in the barrier that ends the phase
    unsigned long long phase_ns[8];  // accounting (CTA 0): ns in phase A, its barrier, phase B, its barrier, phase C, its barrier; [6] = iterations
    unsigned long long slots[kMaxPartials][2];  // CTA i -> root: its partial sum as two LL packets
};

struct CgArgs {
    TmaSpmvArgs mv;          // x = P shifted to global column 0, y = Q, dot_w = P (owned piece)
    double *p, *q, *x, *r;   // owned pieces
    int64_t n, head, npacks; // 256-bit body shared by the four vectors
    double *rr_cur, *rr_new, *p_norm;
    double *hist;
    long long hist_cap;
    long long *hist_count;
    int niter;
    GridSync *gs;
    const lsk_peers *peers;  // device copy; null on one rank
    HaloSpec halo;           // what to push (send side) and whom to expect data from
    long long own_lo;
    const unsigned char *ghost_blocks;  // per row block: references a ghost column (null = unknown)
};

// one thread per row: flags[row / rpb] = 1 if the row references a column outside [own_lo, own_lo + own_n)
__global__ void __launch_bounds__(kBlock) ghost_blocks_kernel(int64_t rows, int rpb, const lsk_rect *__restrict__ rowptr,
                                                              int64_t k_base, const long long *__restrict__ col, long long own_lo,
                                                              unsigned long long own_n, unsigned char *flags) {
    for (int64_t r = (int64_t) blockIdx.x * kBlock + threadIdx.x; r < rows; r += (int64_t) gridDim.x * kBlock) {
        const longlong2 rc = __ldg(reinterpret_cast<const longlong2 *>(rowptr + r));
        bool ghost = false;
        for (long long k = rc.x - k_base; k <= rc.y - k_base; ++k)
            ghost |= ((unsigned long long) (__ldg(col + k) - own_lo) >= own_n);
        if (ghost) flags[r / rpb] = 1;
    }
}

__device__ __forceinline__ unsigned int ld_volatile_u32(const unsigned int *p) {
    return *reinterpret_cast<const volatile unsigned int *>(p);
}

// ---- grid barrier with the reduction riding on it ---------------------------------------------------------------
// All CTAs are co-resident (cooperative launch).  Gather-broadcast through CTA 0 with LL packets -- an aligned
// 8-byte word carries 32 data bits and the barrier generation, so data and arrival are ONE store and one poll:
//   every CTA:  block sum -> thread 0: fence (release), 2 packet stores into its slot, then polls `release`
//   CTA 0:      its 256 threads poll the slots of all other CTAs (values arrive in registers), fixed-order sum,
//               [cross-rank sum over NVLink], tail(), fence, 2 packet stores to `release`
// No atomics, no second pass over partials, no separate flag: ~3 us of critical path instead of ~7 for the
// counter-based barrier.  Results are bitwise reproducible for a given grid size.
__device__ __forceinline__ void ll_store(unsigned long long *pk, double v, unsigned int tag) {
    const unsigned long long bits = (unsigned long long) __double_as_longlong(v);
    const unsigned long long t = (unsigned long long) tag << 32;
    *reinterpret_cast<volatile unsigned long long *>(pk) = t | (bits & 0xffffffffull);
    *reinterpret_cast<volatile unsigned long long *>(pk + 1) = t | (bits >> 32);
}
// false = gave up
__device__ __forceinline__ bool ll_poll(const unsigned long long *pk, unsigned int tag, double &v) {
    const volatile unsigned long long *p = pk;
    unsigned long long a, b;
    unsigned int polls = 0;
    unsigned long long t_start = 0;
    for (;;) {
        a = p[0];
        b = p[1];
        if ((unsigned int) (a >> 32) == tag && (unsigned int) (b >> 32) == tag) break;
        if (spin_expired(polls, t_start)) return false;
    }
    v = __longlong_as_double((long long) ((a & 0xffffffffull) | (b << 32)));
    return true;
}

// REDUCE: returns the sum of `v` over all threads of all CTAs -- and over all ranks when `peers` is set -- identical
// bits in every thread.  `tail()` is executed by every thread of CTA 0 after all CTAs have arrived and before anybody
// is released.  `gen` is the caller's running generation counter (same value in every thread of the grid).
template <bool REDUCE, class Tail>
__device__ __forceinline__ double grid_sync(GridSync *gs, unsigned int &gen, const lsk_peers *peers, double v, bool &err, Tail tail) {
    __shared__ double s_red[kWarps];
    __shared__ double s_val;
    __shared__ int s_err;
    const unsigned int G = gridDim.x;
    const unsigned int tag = ++gen;
    double b = 0.0;
    if constexpr (REDUCE) b = block_sum(v, s_red);  // valid in warp 0
    if (threadIdx.x == 0) s_err = 0;
    // q, x, r, p are written with ordinary (generic-proxy) stores and re-read in the next phase by cp.async.bulk
    // (async proxy), possibly from another CTA: order them across the proxies before this thread arrives
    asm volatile("fence.proxy.async;" ::: "memory");
    __syncthreads();  // every thread's writes of the phase are ordered before thread 0's fence below
    if (blockIdx.x != 0) {
        if (threadIdx.x == 0) {
            __threadfence();
            ll_store(gs->slots[blockIdx.x], b, tag);
            double r = 0.0;
            if (!ll_poll(gs->release, tag, r)) {
                gs->error = 1;
                s_err = 1;
            }
            __threadfence();
            s_val = r;
        }
    } else {
        double acc = (threadIdx.x == 0) ? b : 0.0;  // CTA 0's own partial takes the place of slot 0
        bool ok = true;
        for (unsigned int i = threadIdx.x; i < G; i += kBlock) {
            if (i == 0) continue;
            double x = 0.0;
            ok = ll_poll(gs->slots[i], tag, x) && ok;
            if constexpr (REDUCE) acc += x;
        }
        if (!ok) {
            gs->error = 1;
            s_err = 1;
        }
        double tot = block_sum(acc, s_red);  // its barriers also mean: every CTA has arrived
        if constexpr (REDUCE) {
            if (peers != nullptr && peers->nranks > 1 && threadIdx.x < 32) {
                double w[kMaxRed];
                w[0] = tot;
                allreduce_warp(*peers, w, 1);
                tot = w[0];
            }
        }
        __threadfence();  // acquire side of the arrivals observed by this thread, before the tail acts on them
        tail();
        if (threadIdx.x == 0) {
            __threadfence();
            ll_store(gs->release, tot, tag);
            s_val = tot;
        }
    }
    __syncthreads();
    err = (s_err != 0);
    return REDUCE ? s_val : 0.0;
}

// The mat-vec phase.  Measured: as a real (noinline) call it needs no spills but runs 9 % slower -- behind a call the
// shared-memory ring is reached through generic pointers -- so it is inlined and the kernel around it keeps its
// live state small instead (rarely used state lives in shared memory).
struct MatvecPhaseShared {
    uint64_t full[kTmaStages];
    long long lo[2][kWarps], hi[2][kWarps];
    long long rbq[2];
};
template <bool GATED, int PART>
__device__ __forceinline__ double matvec_phase(const TmaSpmvArgs &mv, TmaSpmvState &st, unsigned char *s_dyn, MatvecPhaseShared *sh,
                                               const GhostGate *gate, TmaCursor &cur, unsigned long long *counter) {
    double dacc[1] = {0.0};
    TmaDynamic dyn;
    dyn.counter = counter;
    dyn.s_rbq = sh->rbq;
    // several ranks: start from the middle of the slab, so that the row blocks at its two ends -- the ones that read
    // ghost columns -- come up when the neighbours' halo has long arrived
    dyn.rot = GATED ? mv.n_row_blocks / 2 : 0;
    // Measured on the 256^3 system: dynamic hand-out shortens the tail of the phase (10 -> 5 us) but the walk itself
    // gets 12 us slower; the static strided walk wins for the mat-vec (it is not limited by per-SM memory share
    // the way the vector phases are), so DYN stays off here.
    constexpr bool kDynamicMatvec = false;
    csr_tma_run<1, true, GATED, PART, kDynamicMatvec>(mv, st, s_dyn, sh->full, sh->lo, sh->hi, dacc, gate, &cur, &dyn);
    return dacc[0];
}

__global__ void __launch_bounds__(kBlock, LSK_TMA_MINB) cg_persistent_kernel(CgArgs a) {
    extern __shared__ __align__(128) unsigned char s_dyn[];
    __shared__ __align__(8) MatvecPhaseShared s_mv;
    __shared__ __align__(8) uint64_t s_vbar[kVecStages];
    __shared__ long long s_vchunk[kVecStages];
    TmaSpmvState st;
    csr_tma_init(st, s_mv.full);
    VecRing ring;
    ring.smem = s_dyn;
    ring.bar = s_vbar;
    ring.chunk = s_vchunk;
    ring.phases = 0;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < kVecStages; ++s) mbar_init(&s_vbar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }

    GridSync *gs = a.gs;
    const lsk_peers *peers = a.peers;
    const bool multi = (peers != nullptr && peers->nranks > 1);
    CommWindow *me = multi ? static_cast<CommWindow *>(peers->window[peers->rank]) : nullptr;
    // exchanges are numbered per pair of ranks (CommWindow::halo_sent); s_hbase[q] = the pair counter of move q at entry
    __shared__ unsigned long long s_hbase[4];
    // state that only one thread (or only the rare ghost path) needs lives in shared memory, not in registers:
    // the mat-vec phase runs at the register limit of 3 CTAs per SM
    __shared__ GhostGate gate;
    __shared__ long long s_hcount;
    __shared__ unsigned long long s_tmark;
    const bool scribe = (blockIdx.x == 0 && threadIdx.x == 0);
    if (threadIdx.x == 0) {
        gate.blocks = a.ghost_blocks;
        gate.nflags = 0;
        gate.error = &gs->error;
        if (multi) {
            for (int q = 0; q < a.halo.nmoves; ++q) {
                s_hbase[q] = *reinterpret_cast<volatile unsigned long long *>(&me->halo_sent[a.halo.m[q].peer]);
                if (a.halo.m[q].expect) {
                    gate.want[gate.nflags] = s_hbase[q];
                    gate.flag[gate.nflags++] = &me->halo_done[a.halo.m[q].peer];
                }
            }
        }
        s_hcount = scribe ? *reinterpret_cast<volatile long long *>(a.hist_count) : 0;
        s_tmark = global_ns();
    }
    __syncthreads();

    double rr = *reinterpret_cast<volatile double *>(a.rr_cur);
    double pq = 0.0;
    unsigned int gen = *reinterpret_cast<volatile unsigned int *>(&gs->gen);  // written back by the scribe at the end
    bool err = false;
    int done = 0;

    auto lap = [&](int which) {  // accounting by one thread: time since the previous lap goes to phase `which`
        if (scribe) {
            const unsigned long long now = global_ns();
            gs->phase_ns[which] += now - s_tmark;
            s_tmark = now;
        }
    };
    // ragged edges of the vectors (outside the 32-byte aligned body): elements [0, head) and [tail0, n)
    auto for_each_edge = [&](auto &&body) {
        const int64_t tail0 = a.head + a.npacks * 4;
        const int64_t nedge = a.head + (a.n - tail0);
        for (int64_t e = (int64_t) blockIdx.x * kBlock + threadIdx.x; e < nedge; e += (int64_t) gridDim.x * kBlock)
            body(e < a.head ? e : tail0 + (e - a.head));
    };
    TmaCursor cur;
    if (multi) matvec_phase<true, 1>(a.mv, st, s_dyn, &s_mv, &gate, cur, &gs->work[0]);
    else matvec_phase<false, 1>(a.mv, st, s_dyn, &s_mv, nullptr, cur, &gs->work[0]);
    for (int it = 0; it < a.niter; ++it) {
        // ---- phase A: q = A p, partial p.q -------------------------------------------------------------
        if (threadIdx.x == 0 && multi) {  // published by the mat-vec's first CTA barrier
            int f = 0;
            for (int q = 0; q < a.halo.nmoves; ++q)
                if (a.halo.m[q].expect) gate.want[f++] = s_hbase[q] + (unsigned long long) it;
        }
        // (its prologue -- first rects, first matrix tile in flight -- ran before the previous grid barrier)
        const double pq_part = multi ? matvec_phase<true, 2>(a.mv, st, s_dyn, &s_mv, &gate, cur, &gs->work[0])
                                     : matvec_phase<false, 2>(a.mv, st, s_dyn, &s_mv, nullptr, cur, &gs->work[0]);
        lap(0);
        pq = grid_sync<true>(gs, gen, peers, pq_part, err, [&] {
            if (threadIdx.x == 0) *reinterpret_cast<volatile unsigned long long *>(&gs->work[0]) = 0ull;  // everybody is done with phase A
        });
        lap(1);
        if (err) break;

        // ---- phase B: x += (rr/pq) p;  r += ((-1*rr)/pq) q;  partial r.r -------------------------------------
        const double a1 = div_rn(rr, pq);
        const double a2 = div_rn(mul_rn(-1.0, rr), pq);
        double racc = 0.0;
        {
            const double *const in[4] = {a.p, a.q, a.x, a.r};
            vec_stream<4>(ring, in, a.head, a.npacks * 4, &gs->work[1], 0, [](int64_t, int) {}, [&](int64_t i, const double (&v)[4][2]) {
                const double x0 = fma_rn(a1, v[0][0], v[2][0]), x1 = fma_rn(a1, v[0][1], v[2][1]);
                const double r0 = fma_rn(a2, v[1][0], v[3][0]), r1 = fma_rn(a2, v[1][1], v[3][1]);
                *reinterpret_cast<double2 *>(a.x + i) = make_double2(x0, x1);
                *reinterpret_cast<double2 *>(a.r + i) = make_double2(r0, r1);
                racc = fma(r0, r0, racc);
                racc = fma(r1, r1, racc);
            });
        }
        for_each_edge([&](int64_t i) {
            a.x[i] = fma_rn(a1, ld_f64(a.p + i), a.x[i]);
            const double rn = fma_rn(a2, ld_f64(a.q + i), a.r[i]);
            a.r[i] = rn;
            racc = fma(rn, rn, racc);
        });
        lap(2);
        const double rr_new = grid_sync<true>(gs, gen, peers, racc, err, [&] {
            if (threadIdx.x == 0) *reinterpret_cast<volatile unsigned long long *>(&gs->work[1]) = 0ull;  // everybody is done with phase B
        });
        lap(3);
        if (err) break;

        // ---- phase C: p = fma(rr_new/rr, p, r), boundary mirrored into the neighbours' ghosts -------------
        const double beta = div_rn(rr_new, rr);
        bool remote = false;
        {
            const double *const in[2] = {a.p, a.r};
            bool chunk_halo = false;  // this chunk overlaps a range that is mirrored into a neighbour (uniform per chunk)
            auto begin = [&](int64_t i0, int cnt) { chunk_halo = multi && halo_chunk_overlaps(a.halo, i0, cnt); };
            const int64_t rot = multi ? halo_first_chunk(a.halo, a.head, kVecStageBytes / 16) : 0;
            vec_stream<2>(ring, in, a.head, a.npacks * 4, &gs->work[2], rot, begin, [&](int64_t i, const double (&v)[2][2]) {
                const double p0 = fma_rn(beta, v[0][0], v[1][0]), p1 = fma_rn(beta, v[0][1], v[1][1]);
                *reinterpret_cast<double2 *>(a.p + i) = make_double2(p0, p1);
                if (chunk_halo) remote |= halo_mirror_pair(a.halo, i, p0, p1);
            });
        }
        for_each_edge([&](int64_t i) {
            const double v = fma_rn(beta, ld_f64(a.p + i), ld_f64(a.r + i));
            a.p[i] = v;
            if (multi) remote |= halo_mirror_one(a.halo, i, v);
        });
        if (remote) __threadfence_system();  // my stores into the neighbours' memory are visible there
        const bool final_it = (it + 1 == a.niter);
        // the ring is free again: start the next mat-vec's first matrix tile now, so that it lands during the barrier
        if (!final_it) {
            if (multi) matvec_phase<true, 1>(a.mv, st, s_dyn, &s_mv, &gate, cur, &gs->work[0]);
            else matvec_phase<false, 1>(a.mv, st, s_dyn, &s_mv, nullptr, cur, &gs->work[0]);
        }
        lap(4);
        grid_sync<false>(gs, gen, peers, 0.0, err, [&] {
            if (threadIdx.x == 0) *reinterpret_cast<volatile unsigned long long *>(&gs->work[2]) = 0ull;  // everybody is done with phase C
            if (!multi) return;
            // every CTA's remote stores are fenced and ordered before its arrival: publish the epoch
            const unsigned long long t0 = global_ns();
            __threadfence_system();
            if (threadIdx.x < a.halo.nmoves) {
                const lsk_halo_move &mv = a.halo.m[threadIdx.x];
                const unsigned long long e_now = s_hbase[threadIdx.x] + (unsigned long long) it + 1ull;
                if (mv.n > 0) {
                    CommWindow *dst = static_cast<CommWindow *>(peers->window[mv.peer]);
                    *reinterpret_cast<volatile unsigned long long *>(&dst->halo_done[peers->rank]) = e_now;
                }
                // leaving the kernel: the ghosts must be current for whatever runs next on this stream
                if (final_it && mv.expect) spin_until(&me->halo_done[mv.peer], e_now, &gs->error);
                if (final_it) me->halo_sent[mv.peer] = e_now;
            }
            __syncthreads();
            if (final_it) {
                __threadfence_system();
                if (threadIdx.x == 0) {
                    me->halo_calls += (unsigned long long) a.niter;
                    me->halo_wait_ns += global_ns() - t0;
                }
            }
        });
        lap(5);
        if (scribe) {
            a.hist[s_hcount % a.hist_cap] = rr_new;  // residual_norm_squared.push_back (src/CGSolver.hpp:53)
            ++s_hcount;
            gs->phase_ns[6] += 1ull;
        }
        rr = rr_new;
        ++done;
        if (err) break;
    }
    if (scribe) {
        gs->gen = gen;  // every CTA has left the last barrier's arrival side; nobody reads gen again in this launch
        *a.hist_count = s_hcount;
        if (done > 0) {
            *a.rr_cur = rr;
            *a.rr_new = rr;
            *a.p_norm = pq;
        }
        if (multi && gs->error) me->error = 1;
    }
    if (err && multi && done < a.niter && blockIdx.x == 0 && threadIdx.x < a.halo.nmoves) {
        // a wait gave up: keep the epoch arithmetic of later launches consistent anyway
        me->halo_sent[a.halo.m[threadIdx.x].peer] = s_hbase[threadIdx.x] + (unsigned long long) a.niter;
    }
}

}  // namespace lsk

using namespace lsk;

extern "C" {

int lsk_cg_steps_supported(const lsk_cg_problem *pb) {
    if (!pb || pb->rows <= 0 || pb->nnz <= 0 || !pb->entry || !pb->col || !pb->rowptr) return 0;
    const uintptr_t e = reinterpret_cast<uintptr_t>(pb->entry), c = reinterpret_cast<uintptr_t>(pb->col);
    if (e % 8 != 0 || c % 8 != 0 || ((e >> 3) & 1) != ((c >> 3) & 1)) return 0;  // TMA tiles: col/entry 16-byte aligned together
    if (pb->nmoves < 0 || pb->nmoves > 4) return 0;
    return 1;
}

int lsk_cg_steps_f64(lsk_ctx *ctx, lsk_stream s, const lsk_cg_problem *pb, int niter) {
    if (!ctx || !pb || niter < 0) return LSK_E_INVALID;
    if (niter == 0) return 0;
    if (!lsk_cg_steps_supported(pb)) return LSK_E_INVALID;
    if (!pb->p_shifted || !pb->q || !pb->x || !pb->r || !pb->rr_cur || !pb->rr_new || !pb->p_norm || !pb->history ||
        !pb->history_count || pb->history_capacity <= 0)
        return LSK_E_INVALID;
    if (pb->nmoves > 0 && (!pb->moves || !ctx->d_peers)) return LSK_E_INVALID;
    LSK_RETURN_IF_CUDA(cudaSetDevice(ctx->device));
    if (ctx->cg_blocks_per_sm == 0) {
        LSK_RETURN_IF_CUDA(cudaFuncSetAttribute(cg_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kTmaSmem));
        int nb = 0;
        LSK_RETURN_IF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, cg_persistent_kernel, kBlock, kTmaSmem));
        if (nb < 1) return LSK_E_CAPACITY;
        if (nb > LSK_TMA_MINB) nb = LSK_TMA_MINB;
        ctx->cg_blocks_per_sm = nb;
    }
    CgArgs a;
    a.mv.rows = pb->rows;
    a.mv.nnz = pb->nnz;
    a.mv.rpb = tma_rows_per_block(pb->rows, pb->nnz);
    a.mv.n_row_blocks = (pb->rows + a.mv.rpb - 1) / a.mv.rpb;
    a.mv.entry = pb->entry;
    a.mv.col = reinterpret_cast<const long long *>(pb->col);
    a.mv.rowptr = pb->rowptr;
    a.mv.k_base = pb->k_base;
    a.mv.x = pb->p_shifted;
    a.mv.y = pb->q;
    a.p = pb->p_shifted + pb->own_lo;
    a.mv.dot_w = a.p;
    a.mv.accumulate = 0;
    a.q = pb->q;
    a.x = pb->x;
    a.r = pb->r;
    a.n = pb->rows;
    // 256-bit body shared by the four vectors (scalar edges otherwise)
    const uintptr_t m0 = mod32(a.p);
    const bool congruent = m0 % 8 == 0 && mod32(a.q) == m0 && mod32(a.x) == m0 && mod32(a.r) == m0;
    if (congruent) {
        int64_t head = (int64_t) ((32 - m0) % 32) / 8;
        if (head > a.n) head = a.n;
        a.head = head;
        a.npacks = (a.n - head) / 4;
    } else {
        a.head = a.n;
        a.npacks = 0;
    }
    a.rr_cur = pb->rr_cur;
    a.rr_new = pb->rr_new;
    a.p_norm = pb->p_norm;
    a.hist = pb->history;
    a.hist_cap = pb->history_capacity;
    a.hist_count = reinterpret_cast<long long *>(pb->history_count);
    a.niter = niter;
    a.gs = static_cast<GridSync *>(ctx->gridsync);
    a.peers = ctx->d_peers;  // non-null = several ranks: the dot products are summed across them inside the grid barrier
    a.own_lo = pb->own_lo;
    a.ghost_blocks = pb->ghost_blocks;
    a.halo.nmoves = pb->nmoves;
    for (int i = 0; i < 4; ++i) {
        a.halo.lo[i] = 0;
        a.halo.m[i].peer = 0; a.halo.m[i].expect = 0; a.halo.m[i].src = nullptr; a.halo.m[i].dst = nullptr; a.halo.m[i].n = 0;
    }
    for (int i = 0; i < pb->nmoves; ++i) {
        const lsk_halo_move &m = pb->moves[i];
        if (m.n < 0 || (m.n > 0 && (!m.dst || m.src < a.p || m.src + m.n > a.p + a.n))) return LSK_E_INVALID;
        a.halo.m[i] = m;
        a.halo.lo[i] = m.n > 0 ? (int64_t) (m.src - a.p) : 0;
    }
    int64_t cap = (int64_t) ctx->sm_count * ctx->cg_blocks_per_sm;
    if (cap > kMaxPartials) cap = kMaxPartials;
    const int grid = (int) (a.mv.n_row_blocks < cap ? a.mv.n_row_blocks : cap);

    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned) grid);
    cfg.blockDim = dim3(kBlock);
    cfg.dynamicSmemBytes = kTmaSmem;
    cfg.stream = (cudaStream_t) s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    static const char *coop = getenv("LSK_CG_COOPERATIVE");  // developer switch: "0" = plain launch (grid <= one wave anyway)
    cfg.attrs = attr;
    cfg.numAttrs = (coop && coop[0] == '0') ? 0 : 1;
    LSK_RETURN_IF_CUDA(cudaLaunchKernelEx(&cfg, cg_persistent_kernel, a));
    return after_launch(ctx);
}

int64_t lsk_cg_row_blocks(int64_t rows, int64_t nnz) {
    if (rows <= 0) return 0;
    const int rpb = tma_rows_per_block(rows, nnz);
    return (rows + rpb - 1) / rpb;
}

int lsk_cg_ghost_blocks(lsk_ctx *ctx, lsk_stream s, const lsk_cg_problem *pb, uint8_t *flags) {
    if (!ctx || !pb || !flags || pb->rows <= 0 || !pb->rowptr || !pb->col) return LSK_E_INVALID;
    const int rpb = tma_rows_per_block(pb->rows, pb->nnz);
    const int64_t nrb = (pb->rows + rpb - 1) / rpb;
    LSK_RETURN_IF_CUDA(cudaMemsetAsync(flags, 0, (size_t) nrb, (cudaStream_t) s));
    const int grid = stream_grid(ctx, pb->rows, 8);
    ghost_blocks_kernel<<<grid, kBlock, 0, (cudaStream_t) s>>>(pb->rows, rpb, pb->rowptr, pb->k_base,
                                                              reinterpret_cast<const long long *>(pb->col), pb->own_lo,
                                                              (unsigned long long) pb->rows, flags);
    return after_launch(ctx);
}

int lsk_ctx_error(lsk_ctx *ctx, lsk_stream s, int *host_out) {
    if (!ctx || !host_out) return LSK_E_INVALID;
    const GridSync *gs = static_cast<const GridSync *>(ctx->gridsync);
    LSK_RETURN_IF_CUDA(cudaMemcpyAsync(host_out, &gs->error, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t) s));
    LSK_RETURN_IF_CUDA(cudaStreamSynchronize((cudaStream_t) s));
    return 0;
}

int lsk_cg_phase_stats(lsk_ctx *ctx, lsk_stream s, uint64_t *host_out7) {
    if (!ctx || !host_out7) return LSK_E_INVALID;
    const GridSync *gs = static_cast<const GridSync *>(ctx->gridsync);
    LSK_RETURN_IF_CUDA(cudaMemcpyAsync(host_out7, gs->phase_ns, 7 * sizeof(uint64_t), cudaMemcpyDeviceToHost, (cudaStream_t) s));
    LSK_RETURN_IF_CUDA(cudaStreamSynchronize((cudaStream_t) s));
    return 0;
}

size_t lsk_gridsync_bytes(void) { return sizeof(GridSync); }

}  // extern "C"
