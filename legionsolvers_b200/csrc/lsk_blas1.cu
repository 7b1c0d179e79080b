// lsk_blas1.cu -- BLAS-1 leaf tasks, device-resident scalar algebra and the fused solver passes.
//
// Replaces the cuBLAS calls and xpay_kernel of the reference's src/LinearAlgebraTasks.cu.  All of
// these kernels are pure HBM streaming: one persistent wave of CTAs (grid from the SM count),
// 256-bit loads/stores on the 32-byte-aligned body, scalar handling of the ragged head/tail,
// alpha folded on the device from up to four scalar slots, reductions finished on the device in a
// fixed order.  Element-wise arithmetic is exactly the reference CPU body's (std::fma where it uses
// std::fma, a rounded multiply for scal), so vector results are bit-identical to the oracle.
#include <stdlib.h>

#include <initializer_list>
#include <type_traits>

#include "lsk_common.cuh"
#include "lsk_vec_stream.cuh"

namespace lsk {

// ---------------------------------------------------------------------------------------------------
// Generic streaming driver.  F provides:  T, NRED, init(), scalar(i, acc), pack(i, acc), outs(o)
// ---------------------------------------------------------------------------------------------------
template <typename F>
__global__ void __launch_bounds__(kBlock)
stream_kernel(F f, int64_t n, int64_t head, int64_t npacks, double *partials, unsigned int *ticket,
              const lsk_peers *peers, bool defer) {
    using T = typename F::T;
    constexpr int NRED = F::NRED;
    constexpr int EPP = PackOf<T>::N;
    f.init();
    double acc[NRED > 0 ? NRED : 1];
#pragma unroll
    for (int j = 0; j < (NRED > 0 ? NRED : 1); ++j) acc[j] = 0.0;

    const int64_t tid = (int64_t) blockIdx.x * kBlock + threadIdx.x;
    const int64_t stride = (int64_t) gridDim.x * kBlock;
    for (int64_t p = tid; p < npacks; p += stride) f.pack(head + p * EPP, acc);
    // ragged edges [0, head) and [head + npacks*EPP, n); everything when the arrays are not
    // mutually 32-byte congruent (head == n, npacks == 0)
    const int64_t tail0 = head + npacks * EPP;
    const int64_t nedge = head + (n - tail0);
    for (int64_t e = tid; e < nedge; e += stride) f.scalar(e < head ? e : tail0 + (e - head), acc);

    if constexpr (NRED > 0) {
        T *out[NRED];
        f.outs(out);
        grid_reduce_finish<NRED, T>(acc, partials, ticket, out, peers, nullptr, defer);
    }
}

struct Span {
    int64_t head, npacks;
};

// Largest 32-byte-aligned body shared by all arrays, or all-scalar when they are not congruent.
template <typename T>
static Span plan_span(int64_t n, std::initializer_list<const void *> ptrs) {
    constexpr int64_t EPP = 32 / sizeof(T);
    const void *first = *ptrs.begin();
    bool congruent = true;
    for (const void *p : ptrs) congruent = congruent && (mod32(p) == mod32(first)) && (mod32(p) % sizeof(T) == 0);
    Span s;
    if (!congruent) {
        s.head = n;
        s.npacks = 0;
        return s;
    }
    int64_t head = (int64_t) ((32 - mod32(first)) % 32) / (int64_t) sizeof(T);
    if (head > n) head = n;
    s.head = head;
    s.npacks = (n - head) / EPP;
    return s;
}

constexpr int64_t kTmaStreamMinPacksDefault = 1500000;  // 6 M elements: passes that do not fit the L2
static int64_t tma_stream_min_packs() {  // LSK_TMA_STREAM_MIN_PACKS: developer A/B knob
    static const int64_t v = [] {
        const char *e = getenv("LSK_TMA_STREAM_MIN_PACKS");
        return e ? (int64_t) atoll(e) : kTmaStreamMinPacksDefault;
    }();
    return v;
}

template <typename F>
__global__ void __launch_bounds__(kBlock, 3) stream_tma_kernel(F f, int64_t n, int64_t head, int64_t npacks, RedScratch rs);

template <typename F, typename = void>
struct HasTmaForm : std::false_type {};
template <typename F>
struct HasTmaForm<F, std::enable_if_t<std::is_same<typename F::T, double>::value && (F::NIN >= 1)>> : std::true_type {};

template <typename F>
static int tma_stream_configure(lsk_ctx *ctx) {
    static const int family = configure_family_index();
    return configure_once(ctx, family, [] {
        return cudaFuncSetAttribute(stream_tma_kernel<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, kVecStages * kVecStageBytes);
    });
}

// defer: the reduction's cross-rank sum is left to the consumer kernel (grid-stride form only; see lsk_cg_update_f64)
template <typename F>
static int launch_stream(lsk_ctx *ctx, lsk_stream s, F f, int64_t n, Span sp, bool defer = false, bool settled = false) {
    if (n == 0 && F::NRED == 0) return 0;
    if (!settled) {  // a deferred reduction whose designated consumer is not this launch is finished first
        const int rc = settle_pending(ctx, (cudaStream_t) s);
        if (rc != 0) return rc;
    }
    if constexpr (HasTmaForm<F>::value) {
        if (!defer && sp.npacks >= tma_stream_min_packs() && tma_stream_configure<F>(ctx) == 0) {
            const int64_t nchunks = (sp.npacks * 4 + VecChunk<F::NIN>::value - 1) / VecChunk<F::NIN>::value;
            const int64_t cap = (int64_t) ctx->sm_count * 3;
            const int grid = (int) (nchunks < cap ? nchunks : cap);
            const RedScratch rs = next_scratch(ctx);
            stream_tma_kernel<F><<<grid, kBlock, kVecStages * kVecStageBytes, (cudaStream_t) s>>>(f, n, sp.head, sp.npacks, rs);
            return after_launch(ctx);
        }
    }
    const int64_t items = sp.npacks > 0 ? sp.npacks : n;
    const int grid = stream_grid(ctx, items > 0 ? items : 1, 8);
    RedScratch rs = {nullptr, nullptr, nullptr, nullptr, 0};
    if (F::NRED > 0) rs = next_scratch(ctx);
    stream_kernel<F><<<grid, kBlock, 0, (cudaStream_t) s>>>(f, n, sp.head, sp.npacks, rs.partials, rs.ticket,
                                                            F::NRED > 0 ? ctx->d_peers : nullptr, defer);
    return after_launch(ctx);
}

// ---------------------------------------------------------------------------------------------------
// xpay fused with the halo exchange of its result (the p = r + beta p of CG feeds the next mat-vec):
// elements that fall in a send range also leave as LL packets for the neighbour's landing buffer over
// NVLink; every CTA then unpacks its share of the neighbours' packets into this rank's ghost region, so
// when the kernel completes the ghosts of y are current -- no separate exchange launch, no fence, no flag.
// The grid is sized so that all its CTAs are resident at once (a CTA polling for a neighbour's packet must
// never keep a CTA that still has packets to send off the SMs).
// ---------------------------------------------------------------------------------------------------
constexpr int kXpayHaloCtasPerSm = 4;
__global__ void __launch_bounds__(kBlock, kXpayHaloCtasPerSm)
xpay_halo_kernel(Alpha<double> al, const double *__restrict__ x, double *__restrict__ y, int64_t n, int64_t head,
                 int64_t npacks, HaloSpec h, const lsk_peers *peers) {
    __shared__ HaloLive hl;
    halo_begin(hl, h.m, h.nmoves, peers);
    const double a = fold_alpha(al);
    const int64_t tid = (int64_t) blockIdx.x * kBlock + threadIdx.x;
    const int64_t stride = (int64_t) gridDim.x * kBlock;
    for (int64_t p = tid; p < npacks; p += stride) {
        const int64_t i = head + p * 4;
        const Pack32 px = ld256(x + i);
        Pack32 py = ld256(y + i);
        double v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            v[e] = fma_rn(a, PackOf<double>::get(py, e), PackOf<double>::get(px, e));
            PackOf<double>::set(py, e, v[e]);
        }
        st256(y + i, py);
        halo_send_pair(h, hl, i, v[0], v[1]);
        halo_send_pair(h, hl, i + 2, v[2], v[3]);
    }
    const int64_t tail0 = head + npacks * 4;
    const int64_t nedge = head + (n - tail0);
    for (int64_t e = tid; e < nedge; e += stride) {
        const int64_t i = e < head ? e : tail0 + (e - head);
        const double v = fma_rn(a, y[i], x[i]);
        y[i] = v;
        halo_send_one(h, hl, i, v);
    }
    halo_unpack(hl, h.m, h.nmoves, peers);
    // ---- epilogue: the last CTA advances the pair counters
    CommWindow *me = static_cast<CommWindow *>(peers->window[peers->rank]);
    __shared__ bool s_last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int t = atomicAdd(&me->halo_ticket, 1u);
        s_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;
    halo_finish(h.m, h.nmoves, peers);
    if (threadIdx.x == 0) me->halo_ticket = 0u;
}

// ---------------------------------------------------------------------------------------------------
// The two vector passes of the fused CG step, TMA-streamed and dynamically scheduled (lsk_vec_stream.cuh):
// 3 CTAs per SM, a 4 x 16 KB shared-memory ring filled by cp.async.bulk three chunks ahead, chunks handed out
// first come, first served.  7.0 TB/s on the 805 MB x/r update of the 256^3 system, where the grid-stride
// register-staged kernel above reaches 6.3 and the torch copy that defines the "measured peak" 6.5.
// ---------------------------------------------------------------------------------------------------
struct VecKernelShared {
    uint64_t bar[kVecStages];
    long long chunk[kVecStages];
};
__device__ __forceinline__ void vec_ring_init(VecRing &ring, unsigned char *s_dyn, VecKernelShared &sh) {
    ring.smem = s_dyn;
    ring.bar = sh.bar;
    ring.chunk = sh.chunk;
    ring.phases = 0;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < kVecStages; ++s) mbar_init(&sh.bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
}
// ragged edges [0, head) and [head + 4 npacks, n)
template <class Body>
__device__ __forceinline__ void for_each_edge(int64_t n, int64_t head, int64_t npacks, Body body) {
    const int64_t tail0 = head + npacks * 4;
    const int64_t nedge = head + (n - tail0);
    for (int64_t e = (int64_t) blockIdx.x * kBlock + threadIdx.x; e < nedge; e += (int64_t) gridDim.x * kBlock)
        body(e < head ? e : tail0 + (e - head));
}

// src/CGSolver.hpp:50-52: x = fma(rr/pq, p, x); r = fma((-1*rr)/pq, q, r); rr_new = r.r
__global__ void __launch_bounds__(kBlock, 3)
cg_update_tma_kernel(const double *rr_old, double *pq, const double *neg_one, const double *p, const double *q, double *x,
                     double *r, double *rr_new, int64_t n, int64_t head, int64_t npacks, RedScratch rs, bool resolve) {
    extern __shared__ __align__(128) unsigned char s_dyn[];
    __shared__ __align__(8) VecKernelShared sh;
    pdl_launch_dependents();
    VecRing ring;
    vec_ring_init(ring, s_dyn, sh);
    pdl_wait();  // everything below reads what the mat-vec before this kernel produced (q, p.q)
    __shared__ double s_pq[kMaxRed];
    double pqv;
    if (resolve) {  // the mat-vec only SENT its rank's p.q: sum the ranks' packets (and leave the global value in the slot)
        allreduce_resolve(*rs.peers, s_pq, 1, pq);
        pqv = s_pq[0];
    } else {
        pqv = *pq;
    }
    const double a1 = div_rn(*rr_old, pqv);
    const double a2 = div_rn(mul_rn(*neg_one, *rr_old), pqv);
    double racc = 0.0;
    const double *const in[4] = {p, q, x, r};
    vec_stream<4>(ring, in, head, npacks * 4, rs.work, 0, [](int64_t, int) {}, [&](int64_t i, const double (&v)[4][2]) {
        const double x0 = fma_rn(a1, v[0][0], v[2][0]), x1 = fma_rn(a1, v[0][1], v[2][1]);
        const double r0 = fma_rn(a2, v[1][0], v[3][0]), r1 = fma_rn(a2, v[1][1], v[3][1]);
        *reinterpret_cast<double2 *>(x + i) = make_double2(x0, x1);
        *reinterpret_cast<double2 *>(r + i) = make_double2(r0, r1);
        racc = fma(r0, r0, racc);
        racc = fma(r1, r1, racc);
    });
    for_each_edge(n, head, npacks, [&](int64_t i) {
        x[i] = fma_rn(a1, p[i], x[i]);
        const double rn = fma_rn(a2, q[i], r[i]);
        r[i] = rn;
        racc = fma(rn, rn, racc);
    });
    const double acc[1] = {racc};
    double *const out[1] = {rr_new};
    grid_reduce_finish<1, double>(acc, rs.partials, rs.ticket, out, rs.peers, rs.work, rs.defer != 0);
}

// src/CGSolver.hpp:53-54: residual_norm_squared.push_back(rr_new); p = fma(rr_new/rr_cur, p, r) -- plus, on several
// ranks, the halo exchange of p (as xpay_halo_kernel: boundary chunks are taken first and leave as LL packets, every
// CTA unpacks its share of the neighbours' packets after its last chunk), and rr_cur <- rr_new for the next step.
// The scalar bookkeeping is done by the last CTA to finish.
__global__ void __launch_bounds__(kBlock, 3)
cg_direction_tma_kernel(double *rr_cur, double *rr_new, const double *r, double *p, int64_t n, int64_t head, int64_t npacks,
                        HaloSpec h, RedScratch rs, double *hist, long long hist_cap, long long *hist_count, bool resolve,
                        const lsk_peers *resolve_peers) {
    extern __shared__ __align__(128) unsigned char s_dyn[];
    __shared__ __align__(8) VecKernelShared sh;
    __shared__ bool s_last;
    __shared__ HaloLive hl;
    pdl_launch_dependents();
    VecRing ring;
    vec_ring_init(ring, s_dyn, sh);
    pdl_wait();  // rr_new and r come from the update kernel before this one
    const lsk_peers *peers = rs.peers;
    const bool multi = (peers != nullptr && h.nmoves > 0);
    if (multi) halo_begin(hl, h.m, h.nmoves, peers);
    __shared__ double s_rr[kMaxRed];
    if (resolve) allreduce_resolve(*resolve_peers, s_rr, 1, rr_new);  // the update kernel only SENT its rank's r.r
    const double rr_new_v = resolve ? s_rr[0] : *rr_new;
    const double beta = div_rn(rr_new_v, *rr_cur);  // xpay(P, rr_new, rr_cur, R): alpha = f0 / f1
    bool chunk_halo = false;
    const double *const in[2] = {p, r};
    const int64_t rot = multi ? halo_first_chunk(h, head, kVecStageBytes / 16) : 0;
    vec_stream<2>(ring, in, head, npacks * 4, rs.work, rot,
                  [&](int64_t i0, int cnt) { chunk_halo = multi && halo_chunk_overlaps(h, i0, cnt); },
                  [&](int64_t i, const double (&v)[2][2]) {
                      const double p0 = fma_rn(beta, v[0][0], v[1][0]), p1 = fma_rn(beta, v[0][1], v[1][1]);
                      *reinterpret_cast<double2 *>(p + i) = make_double2(p0, p1);
                      if (chunk_halo) halo_send_pair(h, hl, i, p0, p1);
                  });
    for_each_edge(n, head, npacks, [&](int64_t i) {
        const double v = fma_rn(beta, p[i], r[i]);
        p[i] = v;
        if (multi) halo_send_one(h, hl, i, v);
    });
    if (multi) halo_unpack(hl, h.m, h.nmoves, peers);
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int t = atomicAdd(rs.ticket, 1u);
        s_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;
    if (multi) halo_finish(h.m, h.nmoves, peers);
    if (threadIdx.x == 0) {
        const double v = rr_new_v;
        if (hist != nullptr) {
            const long long c = *hist_count;
            hist[c % hist_cap] = v;
            *hist_count = c + 1;
        }
        *rr_cur = v;  // every CTA read the old value before drawing its ticket
        *rs.ticket = 0u;
        *rs.work = 0ull;
    }
}

// ---------------------------------------------------------------------------------------------------
// The two vector passes above as ONE launch, for slabs whose vectors are L2-resident (a multi-GPU run's 2-4 M rows per
// rank, where a 10-15 us kernel is mostly ramp, tail and launch): src/CGSolver.hpp:50-54 --
//     x = fma(rr/pq, p, x);  r = fma((-1*rr)/pq, q, r);  rr' = r.r;                                   [phase 1]
//     history.push_back(rr');  p = fma(rr'/rr, p, r) (+ halo exchange of p);  rr = rr'               [phase 2]
// One CTA per SM, each owning a contiguous run of 32-byte packs in both phases, so the r it produced in phase 1 waits in
// shared memory for phase 2 (16 MB less L2 traffic per iteration on a 2 M-row slab) and the kernel boundary between the
// passes becomes a grid-wide wait on ONE word: the CTA that draws the last ticket folds the CTAs' r.r partials in block
// order, sums across ranks (LL all-reduce), and publishes {value, launch number} as a packet in device memory that every
// CTA polls.  Element-wise arithmetic is that of cg_update / cg_direction; r.r is folded in a different (fixed) order.
// All CTAs must be resident at once: the grid never exceeds the SM count and the kernel needs one CTA per SM.
// ---------------------------------------------------------------------------------------------------
#ifndef LSK_TAIL_UNROLL
#define LSK_TAIL_UNROLL 2
#endif
#ifndef LSK_TAIL_THREADS
#define LSK_TAIL_THREADS 512  // half an SM's registers: the next mat-vec's CTAs (launched programmatically) fit beside it
#endif
constexpr int kTailThreads = LSK_TAIL_THREADS;
constexpr int kTailUnroll = LSK_TAIL_UNROLL;
struct TailSync {  // per context, zeroed at creation
    unsigned long long pkt[2];   // {lo32 | tag, hi32 | tag} of the global r.r of launch `tag`
    unsigned long long launches; // completed launches (the next one's tag is launches + 1)
    unsigned int ticket1, ticket2;
    // accounting (CTA 0's view, accumulated): ns in {p.q resolve, phase 1, r.r wait, phase 2, unpack}, launches
    unsigned long long ns[5], counted;
};
struct CgTailArgs {
    double *rr_cur, *pq, *rr_new;
    const double *neg_one, *q;
    double *p, *x, *r;
    int64_t n, head, npacks, per;  // per: packs per CTA
    HaloSpec h;
    double *partials;
    TailSync *sync;
    const lsk_peers *peers;  // null: one rank
    int resolve_pq, pdl;
    double *hist;
    long long hist_cap;
    long long *hist_count;
};

#ifndef LSK_TAIL_MINB
#define LSK_TAIL_MINB 1  // measured (2 M-row slab, one B200): 116 registers 67.6 us per iteration, capped at 64 registers 70.5
#endif
__global__ void __launch_bounds__(kTailThreads, LSK_TAIL_MINB) cg_tail_kernel(CgTailArgs a) {
    extern __shared__ __align__(128) unsigned char s_dyn[];
    Pack32 *s_r = reinterpret_cast<Pack32 *>(s_dyn);
    __shared__ double s_red[kTailThreads / 32];
    __shared__ double s_val[kMaxRed];
    __shared__ unsigned long long s_tag;
    __shared__ bool s_last;
    __shared__ HaloLive hl;
    const int tid = threadIdx.x;
    const bool multi = (a.peers != nullptr && a.h.nmoves > 0);
    // the mat-vec that follows may start now: its producer warps fill their rings from the (constant) matrix while this
    // kernel runs, its consumers wait for this kernel's completion.  Every CTA of THIS grid is resident by the time the
    // last of them has executed this, so the early CTAs of the successor never stand in the way of the grid-wide wait.
    if (a.pdl) pdl_launch_dependents();
    const unsigned long long t_0 = (blockIdx.x == 0 && tid == 0) ? global_ns() : 0ull;
    if (tid == 0) s_tag = ((*reinterpret_cast<volatile unsigned long long *>(&a.sync->launches) + 1) & 0xffffffffull) << 32;
    if (multi) halo_begin(hl, a.h.m, a.h.nmoves, a.peers);  // (ends with a CTA barrier)
    double pqv;
    if (a.resolve_pq) {  // the mat-vec only SENT its rank's p.q
        allreduce_resolve(*a.peers, s_val, 1, a.pq);
        pqv = s_val[0];
    } else {
        pqv = *a.pq;
    }
    __syncthreads();
    const unsigned long long t_1 = (blockIdx.x == 0 && tid == 0) ? global_ns() : 0ull;
    const unsigned long long tag = s_tag;
    const double rr_old = *a.rr_cur;
    const double a1 = div_rn(rr_old, pqv);
    const double a2 = div_rn(mul_rn(*a.neg_one, rr_old), pqv);
    const int64_t pk_lo = (int64_t) blockIdx.x * a.per;
    const int64_t pk_hi = pk_lo + a.per < a.npacks ? pk_lo + a.per : a.npacks;
    const int64_t mine = pk_hi > pk_lo ? pk_hi - pk_lo : 0;
    const bool backwards = false;  // (kept for experiments: the upper half of the grid walking its packs backwards)
    const int64_t tail0 = a.head + a.npacks * 4;
    const int64_t nedge = a.head + (a.n - tail0);
    // ---- phase 1
    double acc = 0.0;
#pragma unroll kTailUnroll
    for (int64_t k = tid; k < mine; k += kTailThreads) {
        const int64_t j = backwards ? mine - 1 - k : k;
        const int64_t i = a.head + (pk_lo + j) * 4;
        const Pack32 pp = ld256(a.p + i);
        const Pack32 pq4 = ld256(a.q + i);
        Pack32 px = ld256(a.x + i);
        Pack32 pr = ld256(a.r + i);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            PackOf<double>::set(px, e, fma_rn(a1, PackOf<double>::get(pp, e), PackOf<double>::get(px, e)));
            const double rn = fma_rn(a2, PackOf<double>::get(pq4, e), PackOf<double>::get(pr, e));
            PackOf<double>::set(pr, e, rn);
            acc = fma(rn, rn, acc);
        }
        st256(a.x + i, px);
        st256(a.r + i, pr);
        s_r[j] = pr;
    }
    if (blockIdx.x == 0)
        for (int64_t e = tid; e < nedge; e += kTailThreads) {
            const int64_t i = e < a.head ? e : tail0 + (e - a.head);
            a.x[i] = fma_rn(a1, a.p[i], a.x[i]);
            const double rn = fma_rn(a2, a.q[i], a.r[i]);
            a.r[i] = rn;
            acc = fma(rn, rn, acc);
        }
    // ---- r.r: block partial -> ticket -> the last CTA folds, sums across ranks and publishes
    unsigned long long t_2 = 0;
    {
        const double b = block_sum<kTailThreads>(acc, s_red);
        if (blockIdx.x == 0 && tid == 0) t_2 = global_ns();
        if (tid == 0) {
            a.partials[blockIdx.x] = b;
            __threadfence();
            s_last = (atomicAdd(&a.sync->ticket1, 1u) == gridDim.x - 1);
        }
        __syncthreads();
        if (s_last) {
            __threadfence();
            double v = 0.0;
            const volatile double *pj = a.partials;
            if (tid < kBlock)
                for (unsigned int i = tid; i < gridDim.x; i += kBlock) v += pj[i];
            v = block_sum<kTailThreads>(v, s_red);
            if (tid < 32) {
                double vv[1] = {__shfl_sync(0xffffffffu, v, 0)};
                if (a.peers != nullptr && a.peers->nranks > 1) allreduce_warp(*a.peers, vv, 1);
                if (tid == 0) {
                    const unsigned long long bits = (unsigned long long) __double_as_longlong(vv[0]);
                    *a.rr_new = vv[0];
                    a.sync->ticket1 = 0u;
                    volatile unsigned long long *pk = a.sync->pkt;
                    pk[0] = tag | (bits & 0xffffffffull);
                    pk[1] = tag | (bits >> 32);
                }
            }
        }
        if (tid == 0) {
            const volatile unsigned long long *pk = a.sync->pkt;
            unsigned long long lo, hi;
            unsigned int polls = 0;
            unsigned long long t_start = 0;
            bool bad = false;
            do {
                lo = pk[0];
                hi = pk[1];
                if (spin_expired(polls, t_start)) {
                    bad = true;
                    break;
                }
            } while ((lo >> 32) != (tag >> 32) || (hi >> 32) != (tag >> 32));
            s_val[0] = bad ? __longlong_as_double(0x7ff8000000000000ll) : __longlong_as_double((long long) ((lo & 0xffffffffull) | (hi << 32)));
        }
        __syncthreads();
    }
    const unsigned long long t_3 = (blockIdx.x == 0 && tid == 0) ? global_ns() : 0ull;
    const double rr_new_v = s_val[0];
    const double beta = div_rn(rr_new_v, rr_old);  // xpay(P, rr_new, rr_cur, R): alpha = f0 / f1
    // ---- phase 2
    // Packs that overlap a send range first, and by ALL CTAs (global thread k takes boundary pack k): the 2 x 1 MB of
    // packets of a 256^3 halo leave from every SM at once instead of from the few CTAs that own the slab's ends
    // (measured on 8 GPUs: 30 us of waiting per iteration otherwise).  Their r comes from global memory -- written in
    // phase 1 by the owning CTA, ordered by its fence before the ticket -- and their owners skip them below.
    auto boundary_pack = [&](int64_t pk) -> bool {  // does pack pk hold an element some neighbour receives?
        if (!multi) return false;
        const int64_t i = a.head + pk * 4;
        bool any = false;
#pragma unroll
        for (int q = 0; q < kMaxFusedMoves; ++q) any |= (q < a.h.nmoves && i + 4 > a.h.lo[q] && i < a.h.lo[q] + a.h.m[q].n);
        return any;
    };
    if (multi) {
        const int64_t g = (int64_t) blockIdx.x * kTailThreads + tid, gstride = (int64_t) gridDim.x * kTailThreads;
        for (int q = 0; q < a.h.nmoves; ++q) {
            if (a.h.m[q].n <= 0) continue;
            // packs [b0, b1) touch this move's range; a pack shared with an earlier move's range was done there
            int64_t b0 = (a.h.lo[q] - a.head) / 4, b1 = (a.h.lo[q] + a.h.m[q].n - a.head + 3) / 4;
            if (a.h.lo[q] < a.head) b0 = 0;
            if (b1 > a.npacks) b1 = a.npacks;
            for (int64_t pk = b0 + g; pk < b1; pk += gstride) {
                const int64_t i = a.head + pk * 4;
                bool earlier = false;
                for (int q2 = 0; q2 < q; ++q2) earlier |= (i + 4 > a.h.lo[q2] && i < a.h.lo[q2] + a.h.m[q2].n);
                if (earlier) continue;
                Pack32 pp = ld256(a.p + i);
                Pack32 pr;
                asm volatile("ld.global.cg.v4.b64 {%0, %1, %2, %3}, [%4];" : "=l"(pr.q[0]), "=l"(pr.q[1]), "=l"(pr.q[2]), "=l"(pr.q[3]) : "l"(a.r + i) : "memory");
                double v[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    v[e] = fma_rn(beta, PackOf<double>::get(pp, e), PackOf<double>::get(pr, e));
                    PackOf<double>::set(pp, e, v[e]);
                }
                st256(a.p + i, pp);
                halo_send_pair(a.h, hl, i, v[0], v[1]);
                halo_send_pair(a.h, hl, i + 2, v[2], v[3]);
            }
        }
    }
#pragma unroll kTailUnroll
    for (int64_t k = tid; k < mine; k += kTailThreads) {
        const int64_t j = backwards ? mine - 1 - k : k;
        if (boundary_pack(pk_lo + j)) continue;
        const int64_t i = a.head + (pk_lo + j) * 4;
        Pack32 pp = ld256(a.p + i);
        const Pack32 pr = s_r[j];
        double v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            v[e] = fma_rn(beta, PackOf<double>::get(pp, e), PackOf<double>::get(pr, e));
            PackOf<double>::set(pp, e, v[e]);
        }
        st256(a.p + i, pp);
    }
    if (blockIdx.x == 0)
        for (int64_t e = tid; e < nedge; e += kTailThreads) {
            const int64_t i = e < a.head ? e : tail0 + (e - a.head);
            const double v = fma_rn(beta, a.p[i], a.r[i]);
            a.p[i] = v;
            if (multi) halo_send_one(a.h, hl, i, v);
        }
    const unsigned long long t_4 = (blockIdx.x == 0 && tid == 0) ? global_ns() : 0ull;
    if (multi) halo_unpack(hl, a.h.m, a.h.nmoves, a.peers);
    __syncthreads();
    if (tid == 0) {
        if (blockIdx.x == 0) {
            const unsigned long long t_5 = global_ns();
            a.sync->ns[0] += t_1 - t_0; a.sync->ns[1] += t_2 - t_1; a.sync->ns[2] += t_3 - t_2; a.sync->ns[3] += t_4 - t_3; a.sync->ns[4] += t_5 - t_4;
            a.sync->counted += 1;
        }
        __threadfence();
        s_last = (atomicAdd(&a.sync->ticket2, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;
    if (multi) halo_finish(a.h.m, a.h.nmoves, a.peers);
    if (tid == 0) {
        if (a.hist != nullptr) {
            const long long c = *a.hist_count;
            a.hist[c % a.hist_cap] = rr_new_v;
            *a.hist_count = c + 1;
        }
        *a.rr_cur = rr_new_v;  // every CTA read the old value at its start
        a.sync->launches += 1;
        a.sync->ticket2 = 0u;
    }
}

// The same traversal for any fp64 functor that names its input arrays (NIN <= 4) and provides pair(i, v, acc):
// scal / axpy / xpay / dot / dot2 / axpy_dot / bicg_p_update above 6 M elements (below, the vectors of the benchmark
// slabs are L2-resident and the grid-stride kernel with 2048 threads per SM is faster).
template <typename F>
__global__ void __launch_bounds__(kBlock, 3)
stream_tma_kernel(F f, int64_t n, int64_t head, int64_t npacks, RedScratch rs) {
    constexpr int NRED = F::NRED;
    constexpr int NIN = F::NIN;
    extern __shared__ __align__(128) unsigned char s_dyn[];
    __shared__ __align__(8) VecKernelShared sh;
    __shared__ bool s_last;
    VecRing ring;
    vec_ring_init(ring, s_dyn, sh);
    f.init();
    double acc[NRED > 0 ? NRED : 1];
#pragma unroll
    for (int j = 0; j < (NRED > 0 ? NRED : 1); ++j) acc[j] = 0.0;
    const double *in[NIN];
    f.inputs(in);
    const double *const (&cin)[NIN] = in;
    vec_stream<NIN>(ring, cin, head, npacks * 4, rs.work, 0, [](int64_t, int) {}, [&](int64_t i, const double (&v)[NIN][2]) { f.pair(i, v, acc); });
    for_each_edge(n, head, npacks, [&](int64_t i) { f.scalar(i, acc); });
    if constexpr (NRED > 0) {
        double *out[NRED];
        f.outs(out);
        grid_reduce_finish<NRED, double>(acc, rs.partials, rs.ticket, out, rs.peers, rs.work);
    } else {  // the last CTA to finish re-arms the work counter
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            s_last = (atomicAdd(rs.ticket, 1u) == gridDim.x - 1);
        }
        __syncthreads();
        if (s_last && threadIdx.x == 0) {
            *rs.ticket = 0u;
            *rs.work = 0ull;
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// Functors
// ---------------------------------------------------------------------------------------------------
template <typename TT>
struct ScalF {  // ScalTask: x = alpha * x  (src/LinearAlgebraTasks.cpp:41)
    using T = TT;
    static constexpr int NRED = 0;
    Alpha<T> al; T *x; T a;
    __device__ void init() { a = fold_alpha(al); }
    __device__ void scalar(int64_t i, double *) { x[i] = mul_rn(a, x[i]); }
    static constexpr int NIN = 1;
    __device__ void inputs(const double **in) { in[0] = reinterpret_cast<const double *>(x); }
    __device__ void pair(int64_t i, const double (&v)[1][2], double *) {
        *reinterpret_cast<double2 *>(x + i) = make_double2(mul_rn((double) a, v[0][0]), mul_rn((double) a, v[0][1]));
    }
    __device__ void pack(int64_t i, double *) {
        Pack32 px = ld256(x + i);
#pragma unroll
        for (int e = 0; e < PackOf<T>::N; ++e) PackOf<T>::set(px, e, mul_rn(a, PackOf<T>::get(px, e)));
        st256(x + i, px);
    }
};

template <typename TT>
struct AxpyF {  // AxpyTask: y = fma(alpha, x, y)  (src/LinearAlgebraTasks.cpp:84-85)
    using T = TT;
    static constexpr int NRED = 0;
    Alpha<T> al; const T *x; T *y; T a;
    __device__ void init() { a = fold_alpha(al); }
    __device__ void scalar(int64_t i, double *) { y[i] = fma_rn(a, x[i], y[i]); }
    static constexpr int NIN = 2;
    __device__ void inputs(const double **in) { in[0] = reinterpret_cast<const double *>(x); in[1] = reinterpret_cast<const double *>(y); }
    __device__ void pair(int64_t i, const double (&v)[2][2], double *) {
        *reinterpret_cast<double2 *>(y + i) = make_double2(fma_rn((double) a, v[0][0], v[1][0]), fma_rn((double) a, v[0][1], v[1][1]));
    }
    __device__ void pack(int64_t i, double *) {
        const Pack32 px = ld256(x + i);
        Pack32 py = ld256(y + i);
#pragma unroll
        for (int e = 0; e < PackOf<T>::N; ++e)
            PackOf<T>::set(py, e, fma_rn(a, PackOf<T>::get(px, e), PackOf<T>::get(py, e)));
        st256(y + i, py);
    }
};

template <typename TT>
struct XpayF {  // XpayTask: y = fma(alpha, y, x)  (src/LinearAlgebraTasks.cpp:128-129)
    using T = TT;
    static constexpr int NRED = 0;
    Alpha<T> al; const T *x; T *y; T a;
    __device__ void init() { a = fold_alpha(al); }
    __device__ void scalar(int64_t i, double *) { y[i] = fma_rn(a, y[i], x[i]); }
    static constexpr int NIN = 2;
    __device__ void inputs(const double **in) { in[0] = reinterpret_cast<const double *>(x); in[1] = reinterpret_cast<const double *>(y); }
    __device__ void pair(int64_t i, const double (&v)[2][2], double *) {
        *reinterpret_cast<double2 *>(y + i) = make_double2(fma_rn((double) a, v[1][0], v[0][0]), fma_rn((double) a, v[1][1], v[0][1]));
    }
    __device__ void pack(int64_t i, double *) {
        const Pack32 px = ld256(x + i);
        Pack32 py = ld256(y + i);
#pragma unroll
        for (int e = 0; e < PackOf<T>::N; ++e)
            PackOf<T>::set(py, e, fma_rn(a, PackOf<T>::get(py, e), PackOf<T>::get(px, e)));
        st256(y + i, py);
    }
};

template <typename TT>
struct FillF {  // IndexFill
    using T = TT;
    static constexpr int NRED = 0;
    T *x; T v; const T *vdev;
    __device__ void init() { if (vdev) v = *vdev; }
    __device__ void scalar(int64_t i, double *) { x[i] = v; }
    __device__ void pack(int64_t i, double *) {
        Pack32 px;
#pragma unroll
        for (int e = 0; e < PackOf<T>::N; ++e) PackOf<T>::set(px, e, v);
        st256(x + i, px);
    }
};

struct CopyF {  // PartitionedVector::operator= (a region copy in the reference): dst = src
    using T = double;
    static constexpr int NRED = 0;
    const double *src; double *dst;
    __device__ void init() {}
    __device__ void scalar(int64_t i, double *) { dst[i] = src[i]; }
    __device__ void pack(int64_t i, double *) { st256(dst + i, ld256(src + i)); }
};

template <typename TT>
struct DotF {  // DotTask: sum v*w
    using T = TT;
    static constexpr int NRED = 1;
    const T *v, *w; T *out;
    __device__ void init() {}
    __device__ void outs(T **o) { o[0] = out; }
    __device__ void scalar(int64_t i, double *acc) { acc[0] = fma((double) v[i], (double) w[i], acc[0]); }
    static constexpr int NIN = 2;
    __device__ void inputs(const double **in) { in[0] = reinterpret_cast<const double *>(v); in[1] = reinterpret_cast<const double *>(w); }
    __device__ void pair(int64_t, const double (&u)[2][2], double *acc) {
        acc[0] = fma(u[0][0], u[1][0], acc[0]);
        acc[0] = fma(u[0][1], u[1][1], acc[0]);
    }
    __device__ void pack(int64_t i, double *acc) {
        const Pack32 pv = ld256(v + i);
        const Pack32 pw = ld256(w + i);
#pragma unroll
        for (int e = 0; e < PackOf<T>::N; ++e)
            acc[0] = fma((double) PackOf<T>::get(pv, e), (double) PackOf<T>::get(pw, e), acc[0]);
    }
};

struct Dot2F {  // r.u and u.u in one pass (src/BiCGStabSolver.hpp:75-76)
    using T = double;
    static constexpr int NRED = 2;
    const double *v, *w; double *out_vw, *out_ww;
    __device__ void init() {}
    __device__ void outs(double **o) { o[0] = out_vw; o[1] = out_ww; }
    __device__ void scalar(int64_t i, double *acc) {
        const double b = w[i];
        acc[0] = fma(v[i], b, acc[0]);
        acc[1] = fma(b, b, acc[1]);
    }
    static constexpr int NIN = 2;
    __device__ void inputs(const double **in) { in[0] = v; in[1] = w; }
    __device__ void pair(int64_t, const double (&u)[2][2], double *acc) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            acc[0] = fma(u[0][e], u[1][e], acc[0]);
            acc[1] = fma(u[1][e], u[1][e], acc[1]);
        }
    }
    __device__ void pack(int64_t i, double *acc) {
        const Pack32 pv = ld256(v + i);
        const Pack32 pw = ld256(w + i);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const double b = PackOf<double>::get(pw, e);
            acc[0] = fma(PackOf<double>::get(pv, e), b, acc[0]);
            acc[1] = fma(b, b, acc[1]);
        }
    }
};

struct CgUpdateF {  // src/CGSolver.hpp:50-52: two axpys and the r.r dot in one pass
    using T = double;
    static constexpr int NRED = 1;
    const double *rr_old; double *pq; const double *neg_one; const double *p, *q; double *x, *r; double *rr_new;
    const lsk_peers *resolve_peers;  // non-null: p.q was only SENT by the mat-vec; sum the ranks' packets first
    double a1, a2;
    __device__ void init() {
        __shared__ double s_pq[kMaxRed];
        double pqv;
        if (resolve_peers != nullptr) {
            allreduce_resolve(*resolve_peers, s_pq, 1, pq);
            pqv = s_pq[0];
        } else {
            pqv = *pq;
        }
        a1 = div_rn(*rr_old, pqv);                      // axpy(SOL, rr_old, p_norm, P): f0/f1
        a2 = div_rn(mul_rn(*neg_one, *rr_old), pqv);    // axpy(R, -1, rr_old, p_norm, Q): (f0*f1)/f2
    }
    __device__ void outs(double **o) { o[0] = rr_new; }
    __device__ void scalar(int64_t i, double *acc) {
        x[i] = fma_rn(a1, p[i], x[i]);
        const double rn = fma_rn(a2, q[i], r[i]);
        r[i] = rn;
        acc[0] = fma(rn, rn, acc[0]);
    }
    __device__ void pack(int64_t i, double *acc) {
        const Pack32 pp = ld256(p + i);
        const Pack32 pqv = ld256(q + i);
        Pack32 px = ld256(x + i);
        Pack32 pr = ld256(r + i);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            PackOf<double>::set(px, e, fma_rn(a1, PackOf<double>::get(pp, e), PackOf<double>::get(px, e)));
            const double rn = fma_rn(a2, PackOf<double>::get(pqv, e), PackOf<double>::get(pr, e));
            PackOf<double>::set(pr, e, rn);
            acc[0] = fma(rn, rn, acc[0]);
        }
        st256(x + i, px);
        st256(r + i, pr);
    }
};

struct AxpyDotF {  // y = fma(alpha, x, y); out = sum y*w  (w may alias y)
    using T = double;
    static constexpr int NRED = 1;
    Alpha<double> al; const double *x; double *y; const double *w; double *out; double a;
    __device__ void init() { a = fold_alpha(al); }
    __device__ void outs(double **o) { o[0] = out; }
    __device__ void scalar(int64_t i, double *acc) {
        const double yn = fma_rn(a, x[i], y[i]);
        const double wv = (w == y) ? yn : w[i];
        y[i] = yn;
        acc[0] = fma(yn, wv, acc[0]);
    }
    static constexpr int NIN = 3;
    __device__ void inputs(const double **in) { in[0] = x; in[1] = y; in[2] = w; }
    __device__ void pair(int64_t i, const double (&v)[3][2], double *acc) {
        const bool alias = (w == y);  // then the streamed copy of w is the OLD y: use the new one
        const double y0 = fma_rn(a, v[0][0], v[1][0]), y1 = fma_rn(a, v[0][1], v[1][1]);
        *reinterpret_cast<double2 *>(y + i) = make_double2(y0, y1);
        acc[0] = fma(y0, alias ? y0 : v[2][0], acc[0]);
        acc[0] = fma(y1, alias ? y1 : v[2][1], acc[0]);
    }
    __device__ void pack(int64_t i, double *acc) {
        const Pack32 px = ld256(x + i);
        Pack32 py = ld256(y + i);
        const bool alias = (w == y);
        Pack32 pw;
        if (!alias) pw = ld256(w + i);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const double yn = fma_rn(a, PackOf<double>::get(px, e), PackOf<double>::get(py, e));
            PackOf<double>::set(py, e, yn);
            acc[0] = fma(yn, alias ? yn : PackOf<double>::get(pw, e), acc[0]);
        }
        st256(y + i, py);
    }
};

struct BicgPUpdateF {  // src/BiCGStabSolver.hpp:64-69
    using T = double;
    static constexpr int NRED = 0;
    const double *rho_new, *rho_old, *alpha, *omega; const double *v, *r; double *p;
    double nomega, beta;
    __device__ void init() {
        beta = mul_rn(div_rn(*rho_new, *rho_old), div_rn(*alpha, *omega));
        nomega = -(*omega);
    }
    __device__ void scalar(int64_t i, double *) {
        const double t = fma_rn(nomega, v[i], p[i]);   // axpy(P, -omega, V)
        p[i] = fma_rn(beta, t, r[i]);                  // xpay(P, beta, R)
    }
    static constexpr int NIN = 3;
    __device__ void inputs(const double **in) { in[0] = v; in[1] = r; in[2] = p; }
    __device__ void pair(int64_t i, const double (&u)[3][2], double *) {
        const double t0 = fma_rn(nomega, u[0][0], u[2][0]), t1 = fma_rn(nomega, u[0][1], u[2][1]);
        *reinterpret_cast<double2 *>(p + i) = make_double2(fma_rn(beta, t0, u[1][0]), fma_rn(beta, t1, u[1][1]));
    }
    __device__ void pack(int64_t i, double *) {
        const Pack32 pv = ld256(v + i);
        const Pack32 pr = ld256(r + i);
        Pack32 pp = ld256(p + i);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const double t = fma_rn(nomega, PackOf<double>::get(pv, e), PackOf<double>::get(pp, e));
            PackOf<double>::set(pp, e, fma_rn(beta, t, PackOf<double>::get(pr, e)));
        }
        st256(p + i, pp);
    }
};

struct BicgTailF {  // src/BiCGStabSolver.hpp:77-80 plus the next step's rho = r.rt (:63)
    using T = double;
    static constexpr int NRED = 1;
    const double *alpha, *ru, *uu; const double *p, *u, *rt; double *x, *r; double *rho_next;
    double al, om, nom;
    __device__ void init() {
        al = *alpha;
        om = div_rn(*ru, *uu);
        nom = -om;
    }
    __device__ void outs(double **o) { o[0] = rho_next; }
    __device__ void scalar(int64_t i, double *acc) {
        const double r0 = r[i];
        double xv = fma_rn(al, p[i], x[i]);   // axpy(SOL, alpha, P)
        xv = fma_rn(om, r0, xv);              // axpy(SOL, omega, R)
        x[i] = xv;
        const double rn = fma_rn(nom, u[i], r0);  // axpy(R, -omega, U)
        r[i] = rn;
        acc[0] = fma(rn, rt[i], acc[0]);
    }
    __device__ void pack(int64_t i, double *acc) {
        const Pack32 pp = ld256(p + i);
        const Pack32 pu = ld256(u + i);
        const Pack32 prt = ld256(rt + i);
        Pack32 px = ld256(x + i);
        Pack32 pr = ld256(r + i);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const double r0 = PackOf<double>::get(pr, e);
            double xv = fma_rn(al, PackOf<double>::get(pp, e), PackOf<double>::get(px, e));
            xv = fma_rn(om, r0, xv);
            PackOf<double>::set(px, e, xv);
            const double rn = fma_rn(nom, PackOf<double>::get(pu, e), r0);
            PackOf<double>::set(pr, e, rn);
            acc[0] = fma(rn, PackOf<double>::get(prt, e), acc[0]);
        }
        st256(x + i, px);
        st256(r + i, pr);
    }
};

// ---- scalar futures ------------------------------------------------------------------------------------
template <typename T>
__global__ void scalar_op_kernel(int op, const T *a, const T *b, T *out) {
    const T x = a ? *a : (T) 0;
    const T y = b ? *b : (T) 0;
    T r;
    switch (op) {
    case LSK_OP_NEG: r = -x; break;
    case LSK_OP_ADD: r = add_rn(x, y); break;
    case LSK_OP_SUB: r = add_rn(x, -y); break;
    case LSK_OP_MUL: r = mul_rn(x, y); break;
    case LSK_OP_DIV: r = div_rn(x, y); break;
    case LSK_OP_SQRT: r = sqrt(x); break;                    // IEEE-rounded for fp64 and fp32
    case LSK_OP_RSQRT: r = div_rn((T) 1, (T) sqrt(x)); break;  // 1 / sqrt(x), as RSqrtScalarTask
    case LSK_OP_DUMMY: r = (T) 1; break;
    default: r = x; break;
    }
    *out = r;
}

__global__ void scalar_append_kernel(const double *value, double *hist, long long capacity, long long *count,
                                     double *also) {
    const double v = *value;
    const long long n = *count;
    hist[n % capacity] = v;
    *count = n + 1;
    if (also) *also = v;
}

template <typename T>
static int scalar_op(lsk_ctx *ctx, lsk_stream s, int op, const T *a, const T *b, T *out) {
    if (!ctx || !out || op < 0 || op > LSK_OP_COPY) return LSK_E_INVALID;
    const bool binary = (op >= LSK_OP_ADD && op <= LSK_OP_DIV);
    if (op != LSK_OP_DUMMY && !a) return LSK_E_INVALID;
    if (binary && !b) return LSK_E_INVALID;
    {
        const int rc = settle_pending(ctx, (cudaStream_t) s);
        if (rc != 0) return rc;
    }
    scalar_op_kernel<T><<<1, 1, 0, (cudaStream_t) s>>>(op, a, b, out);
    return after_launch(ctx);
}

template <typename T>
static int do_scal(lsk_ctx *ctx, lsk_stream s, int64_t n, Alpha<T> al, T *x) {
    if (!ctx || n < 0 || (n > 0 && !x) || !alpha_ok(al)) return LSK_E_INVALID;
    ScalF<T> f;
    f.al = al; f.x = x;
    return launch_stream(ctx, s, f, n, plan_span<T>(n, {x}));
}
template <typename T>
static int do_axpy(lsk_ctx *ctx, lsk_stream s, int64_t n, Alpha<T> al, const T *x, T *y) {
    if (!ctx || n < 0 || (n > 0 && (!x || !y)) || !alpha_ok(al)) return LSK_E_INVALID;
    AxpyF<T> f;
    f.al = al; f.x = x; f.y = y;
    return launch_stream(ctx, s, f, n, plan_span<T>(n, {x, y}));
}
template <typename T>
static int do_xpay(lsk_ctx *ctx, lsk_stream s, int64_t n, Alpha<T> al, const T *x, T *y) {
    if (!ctx || n < 0 || (n > 0 && (!x || !y)) || !alpha_ok(al)) return LSK_E_INVALID;
    XpayF<T> f;
    f.al = al; f.x = x; f.y = y;
    return launch_stream(ctx, s, f, n, plan_span<T>(n, {x, y}));
}
template <typename T>
static int do_dot(lsk_ctx *ctx, lsk_stream s, int64_t n, const T *v, const T *w, T *out) {
    if (!ctx || n < 0 || (n > 0 && (!v || !w)) || !out) return LSK_E_INVALID;
    DotF<T> f;
    f.v = v; f.w = w; f.out = out;
    return launch_stream(ctx, s, f, n, plan_span<T>(n, {v, w}));
}
template <typename T>
static int do_fill(lsk_ctx *ctx, lsk_stream s, int64_t n, T value, const T *vdev, T *x) {
    if (!ctx || n < 0 || (n > 0 && !x)) return LSK_E_INVALID;
    FillF<T> f;
    f.x = x; f.v = value; f.vdev = vdev;
    return launch_stream(ctx, s, f, n, plan_span<T>(n, {x}));
}

// opt-in to 64 KB of dynamic shared memory for the two CG vector kernels (once per context, i.e. per device)
static int vec_kernels_configure(lsk_ctx *ctx) {
    static const int family = configure_family_index();
    return configure_once(ctx, family, [] {
        cudaError_t e = cudaFuncSetAttribute(cg_update_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kVecStages * kVecStageBytes);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(cg_direction_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kVecStages * kVecStageBytes);
        return e;
    });
}

}  // namespace lsk

using namespace lsk;

extern "C" {

int lsk_scal_f64(lsk_ctx *ctx, lsk_stream s, int64_t n, int nt, const double *f0, const double *f1,
                 const double *f2, const double *f3, double *x) {
    return do_scal<double>(ctx, s, n, make_alpha(nt, f0, f1, f2, f3), x);
}
int lsk_axpy_f64(lsk_ctx *ctx, lsk_stream s, int64_t n, int nt, const double *f0, const double *f1,
                 const double *f2, const double *f3, const double *x, double *y) {
    return do_axpy<double>(ctx, s, n, make_alpha(nt, f0, f1, f2, f3), x, y);
}
int lsk_xpay_f64(lsk_ctx *ctx, lsk_stream s, int64_t n, int nt, const double *f0, const double *f1,
                 const double *f2, const double *f3, const double *x, double *y) {
    return do_xpay<double>(ctx, s, n, make_alpha(nt, f0, f1, f2, f3), x, y);
}
int lsk_xpay_halo_f64(lsk_ctx *ctx, lsk_stream s, int64_t n, int nt, const double *f0, const double *f1,
                      const double *f2, const double *f3, const double *x, double *y, const lsk_halo_move *moves,
                      int nmoves) {
    Alpha<double> al = make_alpha(nt, f0, f1, f2, f3);
    if (!ctx || n < 0 || !alpha_ok(al) || (n > 0 && (!x || !y)) || nmoves < 0 || nmoves > 4 || (nmoves > 0 && !moves))
        return LSK_E_INVALID;
    if (!ctx->d_peers) return LSK_E_INVALID;  // needs lsk_ctx_set_peers
    {
        const int rc = settle_pending(ctx, (cudaStream_t) s);
        if (rc != 0) return rc;
    }
    HaloSpec h;
    if (!halo_spec_fill(h, moves, nmoves, y, n, ctx->h_peers.nranks)) return LSK_E_INVALID;
    const Span sp = plan_span<double>(n, {x, y});
    const int64_t items = sp.npacks > 0 ? sp.npacks : n;
    const int grid = stream_grid(ctx, items > 0 ? items : 1, kXpayHaloCtasPerSm);  // all CTAs resident (see the kernel)
    xpay_halo_kernel<<<grid, kBlock, 0, (cudaStream_t) s>>>(al, x, y, n, sp.head, sp.npacks, h, ctx->d_peers);
    return after_launch(ctx);
}
int lsk_dot_f64(lsk_ctx *ctx, lsk_stream s, int64_t n, const double *v, const double *w, double *out) {
    return do_dot<double>(ctx, s, n, v, w, out);
}
int lsk_scal_f32(lsk_ctx *ctx, lsk_stream s, int64_t n, int nt, const float *f0, const float *f1,
                 const float *f2, const float *f3, float *x) {
    return do_scal<float>(ctx, s, n, make_alpha(nt, f0, f1, f2, f3), x);
}
int lsk_axpy_f32(lsk_ctx *ctx, lsk_stream s, int64_t n, int nt, const float *f0, const float *f1,
                 const float *f2, const float *f3, const float *x, float *y) {
    return do_axpy<float>(ctx, s, n, make_alpha(nt, f0, f1, f2, f3), x, y);
}
int lsk_xpay_f32(lsk_ctx *ctx, lsk_stream s, int64_t n, int nt, const float *f0, const float *f1,
                 const float *f2, const float *f3, const float *x, float *y) {
    return do_xpay<float>(ctx, s, n, make_alpha(nt, f0, f1, f2, f3), x, y);
}
int lsk_dot_f32(lsk_ctx *ctx, lsk_stream s, int64_t n, const float *v, const float *w, float *out) {
    return do_dot<float>(ctx, s, n, v, w, out);
}

int lsk_fill_f64(lsk_ctx *ctx, lsk_stream s, int64_t n, double value, double *x) {
    return do_fill<double>(ctx, s, n, value, nullptr, x);
}
int lsk_fill_f32(lsk_ctx *ctx, lsk_stream s, int64_t n, float value, float *x) {
    return do_fill<float>(ctx, s, n, value, nullptr, x);
}
int lsk_fill_dev_f64(lsk_ctx *ctx, lsk_stream s, int64_t n, const double *value, double *x) {
    if (!value) return LSK_E_INVALID;
    return do_fill<double>(ctx, s, n, 0.0, value, x);
}
int lsk_copy_f64(lsk_ctx *ctx, lsk_stream s, int64_t n, const double *src, double *dst) {
    if (!ctx || n < 0 || (n > 0 && (!src || !dst))) return LSK_E_INVALID;
    if (n == 0) return 0;
    // A kernel, not cudaMemcpyAsync: a device-to-device copy issued as a memcpy may be scheduled on a copy engine that is
    // busy with the caller's host <-> device traffic (bench.py's end-to-end loop: reset() took 0.19 instead of 0.07 ms)
    const Span sp = plan_span<double>(n, {src, dst});
    if (sp.npacks > 0) {
        CopyF f;
        f.src = src; f.dst = dst;
        return launch_stream(ctx, s, f, n, sp);
    }
    LSK_RETURN_IF_CUDA(cudaMemcpyAsync(dst, src, (size_t) n * sizeof(double), cudaMemcpyDeviceToDevice,
                                       (cudaStream_t) s));
    return 0;
}

int lsk_scalar_op_f64(lsk_ctx *ctx, lsk_stream s, int op, const double *a, const double *b, double *out) {
    return scalar_op<double>(ctx, s, op, a, b, out);
}
int lsk_scalar_op_f32(lsk_ctx *ctx, lsk_stream s, int op, const float *a, const float *b, float *out) {
    return scalar_op<float>(ctx, s, op, a, b, out);
}

int lsk_scalar_append_f64(lsk_ctx *ctx, lsk_stream s, const double *value, double *hist, int64_t capacity,
                          int64_t *count, double *also) {
    if (!ctx || !value || !hist || !count || capacity <= 0) return LSK_E_INVALID;
    {
        const int rc = settle_pending(ctx, (cudaStream_t) s);
        if (rc != 0) return rc;
    }
    scalar_append_kernel<<<1, 1, 0, (cudaStream_t) s>>>(value, hist, capacity, reinterpret_cast<long long *>(count), also);
    return after_launch(ctx);
}

int lsk_cg_update_f64(lsk_ctx *ctx, lsk_stream s, int64_t n, const double *rr_old, const double *pq,
                      const double *p, const double *q, double *x, double *r, double *rr_new) {
    if (!ctx || n < 0 || !rr_old || !pq || !rr_new || (n > 0 && (!p || !q || !x || !r))) return LSK_E_INVALID;
    // deferred all-reduces (lsk_ctx_defer_next_allreduce): this kernel is the designated consumer of a p.q still in flight,
    // and may itself leave the cross-rank sum of r.r to the next lsk_cg_direction_f64
    const bool resolve = ctx->pending_slot != nullptr && ctx->pending_slot == pq && ctx->d_peers != nullptr;
    if (resolve) ctx->pending_slot = nullptr;
    if (!resolve) {
        const int rc = settle_pending(ctx, (cudaStream_t) s);
        if (rc != 0) return rc;
    }
    const bool defer = take_defer(ctx);
    double *pq_rw = const_cast<double *>(pq);  // the global value is stored back into the slot by the resolving kernel
    const Span sp = plan_span<double>(n, {p, q, x, r});
    // Streamed form for passes that do not fit the L2 (measured: 805 MB pass 6.3 -> 7.0 TB/s); for an L2-resident
    // 2 M-row slab the grid-stride kernel with 2048 threads per SM is the faster one (16 vs 14 us).
    if (sp.npacks >= tma_stream_min_packs() && vec_kernels_configure(ctx) == 0) {
        const int64_t nchunks = (sp.npacks * 4 + 511) / 512;
        const int64_t cap = (int64_t) ctx->sm_count * 3;
        const int grid = (int) (nchunks < cap ? nchunks : cap);
        RedScratch rs = next_scratch(ctx);
        if (defer) rs.defer = 1;
        if (defer) ctx->pending_slot = rr_new;
        LSK_RETURN_IF_CUDA(launch_pdl(kPdlUpdate, cg_update_tma_kernel, grid, kBlock, (size_t) kVecStages * kVecStageBytes, (cudaStream_t) s, rr_old, pq_rw,
                                      (const double *) (ctx->consts + 1), p, q, x, r, rr_new, n, sp.head, sp.npacks, rs, resolve));
        return after_launch(ctx);
    }
    CgUpdateF f;
    f.rr_old = rr_old; f.pq = pq_rw; f.neg_one = ctx->consts + 1;
    f.p = p; f.q = q; f.x = x; f.r = r; f.rr_new = rr_new;
    f.resolve_peers = resolve ? ctx->d_peers : nullptr;
    const int rc = launch_stream(ctx, s, f, n, sp, defer, true);
    if (rc == 0 && defer) ctx->pending_slot = rr_new;
    return rc;
}

int lsk_cg_direction_f64(lsk_ctx *ctx, lsk_stream s, int64_t n, double *rr_cur, const double *rr_new, const double *r, double *p,
                         const lsk_halo_move *moves, int nmoves, double *history, int64_t history_capacity,
                         int64_t *history_count) {
    if (!ctx || n < 0 || !rr_cur || !rr_new || (n > 0 && (!r || !p)) || nmoves < 0 || nmoves > 4 || (nmoves > 0 && !moves))
        return LSK_E_INVALID;
    if (history && (!history_count || history_capacity <= 0)) return LSK_E_INVALID;
    if (nmoves > 0 && !ctx->d_peers) return LSK_E_INVALID;  // needs lsk_ctx_set_peers
    const Span sp = plan_span<double>(n, {r, p});
    if (sp.npacks < 1 || vec_kernels_configure(ctx) != 0) return LSK_E_INVALID;  // callers check lsk_cg_direction_supported
    // the designated consumer of an r.r whose cross-rank sum is still in flight (lsk_ctx_defer_next_allreduce)
    const bool resolve = ctx->pending_slot != nullptr && ctx->pending_slot == rr_new && ctx->d_peers != nullptr;
    if (resolve) {
        ctx->pending_slot = nullptr;
    } else {
        const int rc = settle_pending(ctx, (cudaStream_t) s);
        if (rc != 0) return rc;
    }
    HaloSpec h;
    if (!halo_spec_fill(h, moves, nmoves, p, n, ctx->d_peers ? ctx->h_peers.nranks : 1)) return LSK_E_INVALID;
    const int64_t nchunks = (sp.npacks * 4 + 1023) / 1024;
    const int64_t cap = (int64_t) ctx->sm_count * 3;
    const int grid = (int) (nchunks < cap ? nchunks : cap);
    RedScratch rs = next_scratch(ctx);
    if (nmoves == 0) rs.peers = nullptr;
    LSK_RETURN_IF_CUDA(launch_pdl(kPdlDirection, cg_direction_tma_kernel, grid, kBlock, (size_t) kVecStages * kVecStageBytes, (cudaStream_t) s, rr_cur,
                                  const_cast<double *>(rr_new), r, p, n, sp.head, sp.npacks, h, rs, history, (long long) history_capacity,
                                  reinterpret_cast<long long *>(history_count), resolve, (const lsk_peers *) ctx->d_peers));
    return after_launch(ctx);
}

// largest slab the one-launch form is used for (measured on one B200: 2.1 M rows 70.7 -> 67.5 us per iteration, 4.2 M rows
// -- r no longer fits beside the mat-vec's CTA -- 138.5 -> 140.2 us)
constexpr size_t kTailSmemMax = 152 * 1024;
constexpr int64_t kTailPacksPerCta = (int64_t) (kTailSmemMax / sizeof(Pack32));
int lsk_cg_tail_supported(lsk_ctx *ctx, int64_t n, const double *p, const double *q, const double *x, const double *r) {
    if (!ctx || n <= 0 || !p || !q || !x || !r) return 0;
    static const bool off = [] { const char *e = getenv("LSK_CG_TAIL"); return e && e[0] == '0'; }();  // developer A/B switch
    if (off) return 0;
    const Span sp = plan_span<double>(n, {p, q, x, r});
    // every CTA keeps its share of r in shared memory, next to one CTA of the following mat-vec (72 KB)
    return (sp.npacks >= 1 && sp.npacks <= (int64_t) ctx->sm_count * kTailPacksPerCta && ctx->tail_sync != nullptr) ? 1 : 0;
}

int lsk_cg_tail_stats(lsk_ctx *ctx, lsk_stream s, uint64_t *host_out6) {
    if (!ctx || !host_out6 || !ctx->tail_sync) return LSK_E_INVALID;
    const TailSync *ts = static_cast<const TailSync *>(ctx->tail_sync);
    LSK_RETURN_IF_CUDA(cudaMemcpyAsync(host_out6, ts->ns, 6 * sizeof(uint64_t), cudaMemcpyDeviceToHost, (cudaStream_t) s));
    LSK_RETURN_IF_CUDA(cudaStreamSynchronize((cudaStream_t) s));
    return 0;
}

int lsk_cg_tail_f64(lsk_ctx *ctx, lsk_stream s, int64_t n, double *rr_cur, const double *pq, double *rr_new, double *p, const double *q,
                    double *x, double *r, const lsk_halo_move *moves, int nmoves, double *history, int64_t history_capacity,
                    int64_t *history_count) {
    if (!ctx || n <= 0 || !rr_cur || !pq || !rr_new || !p || !q || !x || !r || nmoves < 0) return LSK_E_INVALID;
    if (history && (!history_count || history_capacity <= 0)) return LSK_E_INVALID;
    if (nmoves > 0 && !ctx->d_peers) return LSK_E_INVALID;  // needs lsk_ctx_set_peers
    if (!lsk_cg_tail_supported(ctx, n, p, q, x, r)) return LSK_E_INVALID;
    const bool resolve = ctx->pending_slot != nullptr && ctx->pending_slot == pq && ctx->d_peers != nullptr;
    if (resolve) {
        ctx->pending_slot = nullptr;
    } else {
        const int rc = settle_pending(ctx, (cudaStream_t) s);
        if (rc != 0) return rc;
    }
    CgTailArgs a;
    if (!halo_spec_fill(a.h, moves, nmoves, p, n, ctx->d_peers ? ctx->h_peers.nranks : 1)) return LSK_E_INVALID;
    const Span sp = plan_span<double>(n, {p, q, x, r});
    a.rr_cur = rr_cur; a.pq = const_cast<double *>(pq); a.rr_new = rr_new; a.neg_one = ctx->consts + 1;
    a.p = p; a.q = q; a.x = x; a.r = r; a.n = n; a.head = sp.head; a.npacks = sp.npacks;
    // one CTA per SM, at least ~2 packs per thread so that tiny vectors do not pay for a wide barrier
    int64_t grid = (sp.npacks + 2 * kTailThreads - 1) / (2 * kTailThreads);
    if (grid > ctx->sm_count) grid = ctx->sm_count;
    if (grid < 1) grid = 1;
    a.per = (sp.npacks + grid - 1) / grid;
    grid = (sp.npacks + a.per - 1) / a.per;
    const size_t smem = (size_t) a.per * sizeof(Pack32);  // <= kTailSmemMax (lsk_cg_tail_supported)
    static const int family = configure_family_index();
    {
        const int rc = configure_once(ctx, family, [] { return cudaFuncSetAttribute(cg_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kTailSmemMax); });
        if (rc != 0) return rc;
    }
    const RedScratch rs = next_scratch(ctx);
    a.partials = rs.partials;
    a.sync = static_cast<TailSync *>(ctx->tail_sync);
    a.peers = ctx->d_peers;
    a.resolve_pq = resolve ? 1 : 0;
    static const bool tail_pdl = [] { const char *e = getenv("LSK_TAIL_PDL"); return !(e && e[0] == '0'); }();  // developer A/B switch
    a.pdl = tail_pdl ? 1 : 0;
    a.hist = history; a.hist_cap = (long long) history_capacity; a.hist_count = reinterpret_cast<long long *>(history_count);
    cg_tail_kernel<<<(unsigned) grid, kTailThreads, smem, (cudaStream_t) s>>>(a);
    return after_launch(ctx);
}

int lsk_cg_direction_supported(int64_t n, const double *r, const double *p) {
    if (n <= 0 || !r || !p) return 0;
    return plan_span<double>(n, {r, p}).npacks >= 1 ? 1 : 0;
}

int lsk_axpy_dot_f64(lsk_ctx *ctx, lsk_stream s, int64_t n, int nt, const double *f0, const double *f1,
                     const double *f2, const double *f3, const double *x, double *y, const double *w,
                     double *out) {
    Alpha<double> al = make_alpha(nt, f0, f1, f2, f3);
    if (!ctx || n < 0 || !out || !alpha_ok(al) || (n > 0 && (!x || !y || !w))) return LSK_E_INVALID;
    AxpyDotF f;
    f.al = al; f.x = x; f.y = y; f.w = w; f.out = out;
    return launch_stream(ctx, s, f, n, plan_span<double>(n, {x, y, w}));
}

int lsk_dot2_f64(lsk_ctx *ctx, lsk_stream s, int64_t n, const double *v, const double *w, double *out_vw,
                 double *out_ww) {
    if (!ctx || n < 0 || !out_vw || !out_ww || (n > 0 && (!v || !w))) return LSK_E_INVALID;
    Dot2F f;
    f.v = v; f.w = w; f.out_vw = out_vw; f.out_ww = out_ww;
    return launch_stream(ctx, s, f, n, plan_span<double>(n, {v, w}));
}

int lsk_bicg_p_update_f64(lsk_ctx *ctx, lsk_stream s, int64_t n, const double *rho_new,
                          const double *rho_old, const double *alpha, const double *omega,
                          const double *v, const double *r, double *p) {
    if (!ctx || n < 0 || !rho_new || !rho_old || !alpha || !omega || (n > 0 && (!v || !r || !p)))
        return LSK_E_INVALID;
    BicgPUpdateF f;
    f.rho_new = rho_new; f.rho_old = rho_old; f.alpha = alpha; f.omega = omega;
    f.v = v; f.r = r; f.p = p;
    return launch_stream(ctx, s, f, n, plan_span<double>(n, {v, r, p}));
}

int lsk_bicg_tail_f64(lsk_ctx *ctx, lsk_stream s, int64_t n, const double *alpha, const double *ru,
                      const double *uu, const double *p, const double *u, const double *rt, double *x,
                      double *r, double *rho_next) {
    if (!ctx || n < 0 || !alpha || !ru || !uu || !rho_next || (n > 0 && (!p || !u || !rt || !x || !r)))
        return LSK_E_INVALID;
    BicgTailF f;
    f.alpha = alpha; f.ru = ru; f.uu = uu; f.p = p; f.u = u; f.rt = rt; f.x = x; f.r = r;
    f.rho_next = rho_next;
    return launch_stream(ctx, s, f, n, plan_span<double>(n, {p, u, rt, x, r}));
}

}  // extern "C"
