// lsk_common.cuh -- shared device/host helpers for the sm_100a Krylov kernels.
//
// Everything here is written for Blackwell B200 only (compile with
// -gencode arch=compute_100a,code=sm_100a): 256-bit global loads/stores (LDG.E.256 / STG.E.256),
// L2 eviction-priority hints on the streamed matrix arrays, and grids sized from the SM count.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/lsk.h"

namespace lsk {

constexpr int kBlock = 256;          // threads per CTA for every kernel in this library
constexpr int kWarps = kBlock / 32;
constexpr int kMaxPartials = 4096;   // >= largest grid of any reducing kernel
constexpr int kScratchSets = 32;     // rotating reduction scratch (several reductions in flight)
constexpr int kMaxRed = 2;           // outputs per fused reduction

}  // namespace lsk

// The per-GPU context: replaces the reference's CUDALibraryContext (stream + cuBLAS + cuSPARSE
// handles, src/CUDAUtilities.hpp:44-66) with the only state these kernels need.
struct lsk_ctx {
    int device;
    int sm_count;
    double *partials;          // [kScratchSets][kMaxRed][kMaxPartials]
    unsigned int *tickets;     // [kScratchSets], zero between launches
    double *consts;            // {1.0, -1.0, 0.0}
    int cursor;                // next scratch set
    unsigned long long launches;
    lsk_peers *d_peers;        // device copy of the peer windows; non-null = reducing kernels all-reduce in their tail
    lsk_peers h_peers;         // host copy (valid while d_peers is non-null)
    int defer_next;            // lsk_ctx_defer_next_allreduce: the next fused reduction sends without waiting
    const void *pending_slot;  // device scalar whose cross-rank sum is still in flight (resolved by its consumer)
    unsigned long long *work;  // [kScratchSets] dynamic work counters of the TMA-streamed vector kernels, zero between launches
    void *tail_sync;           // lsk::TailSync of the one-launch CG tail (lsk_cg_tail_f64), zeroed
    unsigned long long configured;  // one bit per kernel family whose dynamic shared-memory opt-in was done on THIS device
                                    // (function attributes are per device: a process may hold contexts on several GPUs)
};

namespace lsk {

struct RedScratch {
    double *partials;      // [kMaxRed][kMaxPartials]
    unsigned int *ticket;
    const lsk_peers *peers;  // non-null: finish the reduction with a cross-rank sum
    unsigned long long *work;  // dynamic work counter (zero at launch; the last CTA resets it)
    int defer;                 // with peers: send the rank sums to the peers and return; the consumer kernel sums them up
};

inline RedScratch next_scratch(lsk_ctx *ctx) {
    const int set = ctx->cursor;
    ctx->cursor = (ctx->cursor + 1) % kScratchSets;
    RedScratch r;
    r.partials = ctx->partials + (size_t) set * kMaxRed * kMaxPartials;
    r.ticket = ctx->tickets + set;
    r.peers = ctx->d_peers;
    r.work = ctx->work + set;
    r.defer = 0;
    ctx->defer_next = 0;  // a request nobody claimed (take_defer) does not carry over to a later launch
    return r;
}

// Streaming kernels are persistent: one wave of CTAs sized from the SM count, grid-stride inside.
inline int stream_grid(const lsk_ctx *ctx, int64_t work_items, int ctas_per_sm) {
    int64_t want = (work_items + kBlock - 1) / kBlock;
    int64_t cap = (int64_t) ctx->sm_count * ctas_per_sm;
    if (cap > kMaxPartials) cap = kMaxPartials;
    if (want < 1) want = 1;
    return (int) (want < cap ? want : cap);
}

template <typename T>
struct Alpha {  // device-resident get_alpha (src/LegionUtilities.cpp:72-97)
    const T *f[4];
    int n;
};

template <typename T>
inline Alpha<T> make_alpha(int n, const T *f0, const T *f1, const T *f2, const T *f3) {
    Alpha<T> a;
    a.f[0] = f0; a.f[1] = f1; a.f[2] = f2; a.f[3] = f3;
    a.n = n;
    return a;
}

template <typename T>
inline bool alpha_ok(const Alpha<T> &a) {
    if (a.n < 0 || a.n > 4) return false;
    for (int i = 0; i < a.n; ++i)
        if (a.f[i] == nullptr) return false;
    return true;
}

#ifdef __CUDACC__

// ---- exact-rounding arithmetic (never let the compiler contract these) ----------------------------
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double div_rn(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ float div_rn(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ double fma_rn(double a, double b, double c) { return __fma_rn(a, b, c); }
__device__ __forceinline__ float fma_rn(float a, float b, float c) { return __fmaf_rn(a, b, c); }

template <typename T>
__device__ __forceinline__ T fold_alpha(const Alpha<T> &a) {
    switch (a.n) {
    case 0: return (T) 1;
    case 1: return *a.f[0];
    case 2: return div_rn(*a.f[0], *a.f[1]);
    case 3: return div_rn(mul_rn(*a.f[0], *a.f[1]), *a.f[2]);
    default: return div_rn(mul_rn(*a.f[0], *a.f[1]), mul_rn(*a.f[2], *a.f[3]));
    }
}

// ---- 256-bit global memory access (sm_100: LDG.E.256 / STG.E.256) ---------------------------------
struct alignas(32) Pack32 {
    unsigned long long q[4];
};

// default cache policy: vectors, which should stay L2-resident between solver passes
__device__ __forceinline__ Pack32 ld256(const void *p) {
    Pack32 r;
    asm volatile("ld.global.v4.b64 {%0, %1, %2, %3}, [%4];"
                 : "=l"(r.q[0]), "=l"(r.q[1]), "=l"(r.q[2]), "=l"(r.q[3])
                 : "l"(p));
    return r;
}
// streamed-once data (matrix values / column indices): bypass L1, mark evict-first in L2 so the
// 126 MB L2 keeps the solver's vectors instead
__device__ __forceinline__ Pack32 ld256_stream(const void *p) {
    Pack32 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v4.b64 {%0, %1, %2, %3}, [%4];"
                 : "=l"(r.q[0]), "=l"(r.q[1]), "=l"(r.q[2]), "=l"(r.q[3])
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void st256(void *p, const Pack32 &v) {
    asm volatile("st.global.v4.b64 [%0], {%1, %2, %3, %4};" ::"l"(p), "l"(v.q[0]), "l"(v.q[1]),
                 "l"(v.q[2]), "l"(v.q[3])
                 : "memory");
}
// 8-byte streamed loads for the ragged edges of a tile
__device__ __forceinline__ unsigned long long ld64_stream(const void *p) {
    unsigned long long r;
    asm volatile("ld.global.nc.L1::no_allocate.b64 %0, [%1];" : "=l"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ unsigned int ld32_stream(const void *p) {
    unsigned int r;
    asm volatile("ld.global.nc.L1::no_allocate.b32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}

// element view of a 32-byte pack
template <typename T>
struct PackOf;
template <>
struct PackOf<double> {
    static constexpr int N = 4;
    __device__ static __forceinline__ double get(const Pack32 &p, int i) {
        return __longlong_as_double((long long) p.q[i]);
    }
    __device__ static __forceinline__ void set(Pack32 &p, int i, double v) {
        p.q[i] = (unsigned long long) __double_as_longlong(v);
    }
};
template <>
struct PackOf<float> {
    static constexpr int N = 8;
    __device__ static __forceinline__ float get(const Pack32 &p, int i) {
        const unsigned long long w = p.q[i >> 1];
        return __uint_as_float((unsigned int) ((i & 1) ? (w >> 32) : (w & 0xffffffffull)));
    }
    __device__ static __forceinline__ void set(Pack32 &p, int i, float v) {
        const unsigned long long b = __float_as_uint(v);
        unsigned long long &w = p.q[i >> 1];
        w = (i & 1) ? ((w & 0x00000000ffffffffull) | (b << 32)) : ((w & 0xffffffff00000000ull) | b);
    }
};
template <>
struct PackOf<long long> {
    static constexpr int N = 4;
    __device__ static __forceinline__ long long get(const Pack32 &p, int i) { return (long long) p.q[i]; }
};

// ---- peer-memory collectives (see lsk_comm.cu) --------------------------------------------------------
struct CommWindow {
    // written by PEERS (remote stores over NVLink)
    unsigned long long ar_pkt[2][LSK_MAX_RANKS][kMaxRed][2];  // all-reduce packets {32 data bits | epoch << 32},
                                                               // by epoch parity, source rank, value, half
    // local state
    unsigned long long ar_epoch;
    // Halo exchanges are counted PER PAIR of ranks: halo_sent[r] = exchanges this rank has completed with peer r.  Both
    // ranks of a pair list each other in the same exchanges (lsk_halo_move), so their pair counters advance together even
    // when other pairs of the job skip that exchange (different halos per block, ranks without neighbours).
    unsigned long long halo_sent[LSK_MAX_RANKS];
    unsigned int halo_ticket;
    int error;
    // accounting (cheap, always on): all-reduce = time inside it on the thread that closes it; halo = per exchange, the
    // longest time any thread spent polling for a packet that had not landed yet
    unsigned long long ar_calls, ar_wait_ns, halo_calls, halo_wait_ns;
    unsigned long long halo_poll_ns;  // of the exchange in progress (atomicMax; folded into halo_wait_ns by the last CTA)
};

__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

constexpr unsigned long long kSpinLimitNs = 4ull * 1000 * 1000 * 1000;  // 4 s, then give up instead of hanging the GPU

// true once a spin loop that started at `t_start` (0 = not started yet) has used up its time budget;
// the clock is read only every 1024 polls
__device__ __forceinline__ bool spin_expired(unsigned int &polls, unsigned long long &t_start) {
    if ((++polls & 1023u) != 0) return false;
    const unsigned long long now = global_ns();
    if (t_start == 0) {
        t_start = now;
        return false;
    }
    return now - t_start > kSpinLimitNs;
}

__device__ __forceinline__ bool spin_until(const volatile unsigned long long *flag, unsigned long long want, int *err) {
    unsigned int polls = 0;
    unsigned long long t_start = 0;
    while (*flag < want) {
        if (spin_expired(polls, t_start)) {
            *err = 1;
            return false;
        }
    }
    return true;
}

// Cross-rank sum of `count` values, executed by ONE warp (all 32 lanes call it).  LL-style protocol: a
// double travels as two 8-byte packets {32 data bits, 32-bit epoch}; an aligned 8-byte store is atomic, so
// the packet both delivers the data and signals its arrival -- no fence, no separate flag, one NVLink
// one-way latency.  Lane r stores this rank's packets into rank r's window and polls rank r's packets in
// its own window; lane 0 then adds the contributions in rank order (identical bits on every rank).
// Slots alternate with the epoch's parity: a peer can run at most one reduction ahead.
__device__ __forceinline__ void allreduce_warp(const lsk_peers &peers, double *v, int count) {
    CommWindow *me = static_cast<CommWindow *>(peers.window[peers.rank]);
    const int r = threadIdx.x & 31;
    const unsigned long long e = me->ar_epoch + 1;
    const unsigned long long tag = (e & 0xffffffffull) << 32;
    const int par = (int) (e & 1);
    const unsigned long long t0 = (r == 0) ? global_ns() : 0ull;
    double got[kMaxRed];
    bool timed_out = false;
    if (r < peers.nranks) {
        CommWindow *dst = static_cast<CommWindow *>(peers.window[r]);
        for (int j = 0; j < count; ++j) {
            const unsigned long long bits = (unsigned long long) __double_as_longlong(v[j]);
            volatile unsigned long long *pk = &dst->ar_pkt[par][peers.rank][j][0];
            pk[0] = tag | (bits & 0xffffffffull);
            pk[1] = tag | (bits >> 32);
        }
        for (int j = 0; j < count; ++j) {
            const volatile unsigned long long *pk = &me->ar_pkt[par][r][j][0];
            unsigned long long a, b;
            unsigned int polls = 0;
            unsigned long long t_start = 0;
            do {
                a = pk[0];
                b = pk[1];
                if (spin_expired(polls, t_start)) {
                    me->error = 1;
                    timed_out = true;
                    break;
                }
            } while ((a >> 32) != (tag >> 32) || (b >> 32) != (tag >> 32));
            // a sum assembled from a packet that never arrived must not look like a number
            got[j] = timed_out ? __longlong_as_double(0x7ff8000000000000ll)
                               : __longlong_as_double((long long) ((a & 0xffffffffull) | (b << 32)));
        }
    }
    // rank-order sum: lane 0 collects lane q's value
    for (int j = 0; j < count; ++j) {
        double sum = 0.0;
        for (int q = 0; q < peers.nranks; ++q) sum += __shfl_sync(0xffffffffu, got[j], q);
        v[j] = sum;
    }
    if (r == 0) {
        me->ar_epoch = e;
        me->ar_calls += 1;
        me->ar_wait_ns += global_ns() - t0;
    }
    __syncwarp();
}

// The same all-reduce SPLIT in two, for a reduction whose only consumer is the next kernel on the stream (the p.q and
// r.r of a fused CG step): the producing kernel's last CTA only SENDS (allreduce_send: packets to every peer, epoch
// advanced, no polling) and exits; every CTA of the consuming kernel polls the packets in its OWN window at its start
// (allreduce_resolve) and adds them in rank order -- identical bits in every CTA of every rank.  The NVLink flight and
// the wait for the slowest rank then overlap the kernel boundary instead of extending the producer's tail.
__device__ __forceinline__ void allreduce_send(const lsk_peers &peers, const double *v, int count) {
    CommWindow *me = static_cast<CommWindow *>(peers.window[peers.rank]);
    const int r = threadIdx.x & 31;
    const unsigned long long e = me->ar_epoch + 1;
    const unsigned long long tag = (e & 0xffffffffull) << 32;
    const int par = (int) (e & 1);
    if (r < peers.nranks) {
        CommWindow *dst = static_cast<CommWindow *>(peers.window[r]);
        for (int j = 0; j < count; ++j) {
            const unsigned long long bits = (unsigned long long) __double_as_longlong(v[j]);
            volatile unsigned long long *pk = &dst->ar_pkt[par][peers.rank][j][0];
            pk[0] = tag | (bits & 0xffffffffull);
            pk[1] = tag | (bits >> 32);
        }
    }
    __syncwarp();
    if (r == 0) {
        me->ar_epoch = e;
        me->ar_calls += 1;
    }
}
// all threads of the CTA call it; s_out[count] is shared memory; `slot` (may be null): CTA 0 also stores the sums there
__device__ __forceinline__ void allreduce_resolve(const lsk_peers &peers, double *s_out, int count, double *slot) {
    if (threadIdx.x < 32) {
        CommWindow *me = static_cast<CommWindow *>(peers.window[peers.rank]);
        const int r = threadIdx.x;
        const unsigned long long e = *reinterpret_cast<volatile unsigned long long *>(&me->ar_epoch);  // advanced by the sender kernel
        const unsigned long long tag = (e & 0xffffffffull) << 32;
        const int par = (int) (e & 1);
        const unsigned long long t0 = (r == 0 && blockIdx.x == 0) ? global_ns() : 0ull;
        double got[kMaxRed];
        for (int j = 0; j < kMaxRed; ++j) got[j] = 0.0;
        if (r < peers.nranks) {
            for (int j = 0; j < count; ++j) {
                const volatile unsigned long long *pk = &me->ar_pkt[par][r][j][0];
                unsigned long long a, b;
                unsigned int polls = 0;
                unsigned long long t_start = 0;
                bool timed_out = false;
                do {
                    a = pk[0];
                    b = pk[1];
                    if (spin_expired(polls, t_start)) {
                        me->error = 1;
                        timed_out = true;
                        break;
                    }
                } while ((a >> 32) != (tag >> 32) || (b >> 32) != (tag >> 32));
                got[j] = timed_out ? __longlong_as_double(0x7ff8000000000000ll)
                                   : __longlong_as_double((long long) ((a & 0xffffffffull) | (b << 32)));
            }
        }
        for (int j = 0; j < count; ++j) {
            double sum = 0.0;
            for (int q = 0; q < peers.nranks; ++q) sum += __shfl_sync(0xffffffffu, got[j], q);
            if (r == 0) {
                s_out[j] = sum;
                if (slot != nullptr && blockIdx.x == 0) slot[j] = sum;
            }
        }
        if (r == 0 && blockIdx.x == 0) me->ar_wait_ns += global_ns() - t0;
    }
    __syncthreads();
}

// ---- halo exchange as LL packets (lsk.h: lsk_halo_move) ----------------------------------------------------------------
// A double travels as one 16-byte packet {lo32 | tag, hi32 | tag}, tag = exchange number of the pair << 32: the packet
// delivers the data AND announces it, so the exchange needs no fence and no flag (a system-scope fence is 1.9 us on an
// idle B200 and 13-20 us while a PCIe copy is in flight; see tools/probe_fence.cu).  Packets land in a buffer owned by the
// receiver (two exchanges, alternating); the receiver's own CTAs copy them into its ghost region.
constexpr int kMaxFusedMoves = 4;  // moves a fused (xpay / direction) kernel carries by value

// bytes of one half of a landing buffer for `count` values: count + 1 packets (the token), rounded up so that the second half
// is 32-byte aligned like the first
__host__ __device__ __forceinline__ size_t ll_half_bytes(int64_t count) { return ((size_t) (count + 1) * 16 + 31) & ~size_t(31); }
__device__ __forceinline__ void ll_store(char *slot, int64_t idx, double v, unsigned long long tag) {
    const unsigned long long bits = (unsigned long long) __double_as_longlong(v);
    const unsigned long long a = tag | (bits & 0xffffffffull), b = tag | (bits >> 32);
    asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(slot + idx * 16), "l"(a), "l"(b) : "memory");
}
// two consecutive packets in ONE 32-byte store (slot + idx * 16 must be 32-byte aligned): a warp then writes 1 KB of whole
// sectors.  Measured on 8 GPUs: with one 16-byte store per packet -- every warp store instruction leaving half-filled
// sectors -- the 2 x 1 MB of a 256^3 halo took 15-40 us to cross NVLink.
__device__ __forceinline__ void ll_store2(char *slot, int64_t idx, double v0, double v1, unsigned long long tag) {
    const unsigned long long b0 = (unsigned long long) __double_as_longlong(v0), b1 = (unsigned long long) __double_as_longlong(v1);
    asm volatile("st.global.v4.b64 [%0], {%1, %2, %3, %4};" ::"l"(slot + idx * 16), "l"(tag | (b0 & 0xffffffffull)), "l"(tag | (b0 >> 32)),
                 "l"(tag | (b1 & 0xffffffffull)), "l"(tag | (b1 >> 32))
                 : "memory");
}
__device__ __forceinline__ bool ll_try_load(const char *slot, int64_t idx, unsigned long long tag, double &v) {
    unsigned long long a, b;
    asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(slot + idx * 16) : "memory");
    if ((a >> 32) != (tag >> 32) || (b >> 32) != (tag >> 32)) return false;
    v = __longlong_as_double((long long) ((a & 0xffffffffull) | (b << 32)));
    return true;
}

struct HaloSpec {
    int nmoves;
    lsk_halo_move m[kMaxFusedMoves];
    int64_t lo[kMaxFusedMoves];  // send range start as an element index into the vector the kernel updates
};
// the exchange a kernel is performing, per move (shared memory; halo_begin fills it)
struct HaloLive {
    unsigned long long tag[kMaxFusedMoves];
    char *send_slot[kMaxFusedMoves];        // the half of the peer's landing buffer this exchange writes
    const char *recv_slot[kMaxFusedMoves];  // the half of this rank's landing buffer the peer writes
};
__device__ __forceinline__ void halo_live_move(HaloLive &hl, int q, const lsk_halo_move &mv, const CommWindow *me) {
    // the pair counter is advanced by the LAST CTA of an exchanging kernel, after every CTA has read it here
    const unsigned long long e = *reinterpret_cast<const volatile unsigned long long *>(&me->halo_sent[mv.peer]) + 1;
    hl.tag[q] = (e & 0xffffffffull) << 32;
    hl.send_slot[q] = static_cast<char *>(mv.ll_send) + (size_t) (e & 1) * ll_half_bytes(mv.n);
    hl.recv_slot[q] = static_cast<const char *>(mv.ll_recv) + (size_t) (e & 1) * ll_half_bytes(mv.recv_n);
}
// All threads of the CTA call it (after pdl_wait: the pair counters belong to the preceding kernel).  CTA 0 also sends
// the token packets: the (n + 1)st packet of every move, in both directions, data or not.
__device__ __forceinline__ void halo_begin(HaloLive &hl, const lsk_halo_move *m, int nmoves, const lsk_peers *peers) {
    const CommWindow *me = static_cast<const CommWindow *>(peers->window[peers->rank]);
    if ((int) threadIdx.x < nmoves) {
        halo_live_move(hl, threadIdx.x, m[threadIdx.x], me);
        if (blockIdx.x == 0) ll_store(hl.send_slot[threadIdx.x], m[threadIdx.x].n, 0.0, hl.tag[threadIdx.x]);
    }
    __syncthreads();
}
// All threads of all CTAs call it once their own packets are out: copy the neighbours' values into the ghost regions.
// Packet idx of move q is handled by global thread (idx mod grid threads); a packet that has not landed yet is polled.
// ADD: the values are added to what recv_dst holds (the reverse exchange of a transposed mat-vec) instead of replacing it
template <bool ADD = false>
__device__ __forceinline__ void halo_unpack(const HaloLive &hl, const lsk_halo_move *m, int nmoves, const lsk_peers *peers) {
    CommWindow *me = static_cast<CommWindow *>(peers->window[peers->rank]);
    const int64_t g = (int64_t) blockIdx.x * blockDim.x + threadIdx.x, stride = (int64_t) gridDim.x * blockDim.x;
    unsigned long long waited = 0;
    for (int q = 0; q < nmoves; ++q) {
        const int64_t cnt = m[q].recv_n;
        double *dst = m[q].recv_dst;
        {   // CTA-level gate: ONE thread looks (at the last packet of the CTA's first pass) until it is there; then everybody
            // reads.  Reading a packet's sector BEFORE the NVLink write lands is what makes the landing slow.  Measured with
            // the one-launch CG tail on 2 GPUs (2.1 M rows each, gpurun_out/r2ak-r2am): every thread polling its own packet
            // right away 94.5 us per iteration (15 us of it waiting for packets that are in place after 10 us when nobody
            // looks); one lane per warp first 88.6; this gate 77.7 -- the two-kernel form, whose unpacking starts 10+ us
            // after the sends anyway, 78-79.
            const int64_t first = (int64_t) blockIdx.x * blockDim.x;
            if (threadIdx.x == 0 && first <= cnt) {
                const int64_t sentinel = first + blockDim.x - 1 < cnt ? first + blockDim.x - 1 : cnt;
                double vs;
                unsigned int nap0 = 128, polls0 = 0;
                unsigned long long t_start0 = 0;
                const unsigned long long t00 = global_ns();
                bool hit = true;
                while (!ll_try_load(hl.recv_slot[q], sentinel, hl.tag[q], vs)) {
                    hit = false;
                    __nanosleep(nap0);
                    if (nap0 < 1024) nap0 <<= 1;
                    if (spin_expired(polls0, t_start0)) break;  // (the loop below reports the timeout)
                }
                if (!hit) waited += global_ns() - t00;
            }
            __syncthreads();
        }
        for (int64_t idx = g; idx <= cnt; idx += stride) {  // idx == cnt: the token
            double v;
            // (later passes, and packets the gate did not vouch for:) lane 0 looks first, the other 31 lanes only once its
            // packet is there -- a warp's packets were sent by one or two store instructions of the peer and land together
            if ((threadIdx.x & 31) == 0) {
                unsigned int nap0 = 128;
                unsigned int polls0 = 0;
                unsigned long long t_start0 = 0;
                const unsigned long long t00 = global_ns();
                bool first = true;
                while (!ll_try_load(hl.recv_slot[q], idx, hl.tag[q], v)) {
                    first = false;
                    __nanosleep(nap0);
                    if (nap0 < 2048) nap0 <<= 1;
                    if (spin_expired(polls0, t_start0)) break;  // (the loop below reports the timeout)
                }
                if (!first) waited += global_ns() - t00;
            }
            __syncwarp(__activemask());
            if (!ll_try_load(hl.recv_slot[q], idx, hl.tag[q], v)) {
                // Not there yet: back off between polls.
                const unsigned long long t0 = global_ns();
                unsigned int polls = 0, nap = 64;
                unsigned long long t_start = 0;
                while (!ll_try_load(hl.recv_slot[q], idx, hl.tag[q], v)) {
                    __nanosleep(nap);
                    if (nap < 1024) nap <<= 1;
                    if (spin_expired(polls, t_start)) {
                        me->error = 1;
                        v = __longlong_as_double(0x7ff8000000000000ll);  // a value that never arrived must not look like a number
                        break;
                    }
                }
                waited += global_ns() - t0;
            }
            if (idx < cnt) dst[idx] = ADD ? add_rn(dst[idx], v) : v;
        }
    }
    if (waited) atomicMax(&me->halo_poll_ns, waited);
}
// Executed by all threads of the LAST CTA of the kernel (after the ticket that made it the last): the exchange is complete
// on this rank -- advance the pair counters.
__device__ __forceinline__ void halo_finish(const lsk_halo_move *m, int nmoves, const lsk_peers *peers) {
    CommWindow *me = static_cast<CommWindow *>(peers->window[peers->rank]);
    if ((int) threadIdx.x < nmoves) me->halo_sent[m[threadIdx.x].peer] += 1;
    if (threadIdx.x == 0) {
        me->halo_calls += 1;
        me->halo_wait_ns += me->halo_poll_ns;
        me->halo_poll_ns = 0;
    }
}

// ---- deterministic reductions ----------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// All threads of the CTA call this; the result is valid in thread 0.  NT = threads per CTA (a multiple of 32).
// The warp totals are combined by the same xor tree for every NT (steps that only meet zeros change nothing), so a
// kernel with extra non-contributing warps produces the bits of the 256-thread form.
template <int NT = kBlock>
__device__ __forceinline__ double block_sum(double v, double *smem /*[NT / 32]*/) {
    constexpr int NW = NT / 32;
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();  // smem may still be in use by a previous call
    if (lane == 0) smem[warp] = v;
    __syncthreads();
    double r = 0.0;
    if (warp == 0) {
        r = lane < NW ? smem[lane] : 0.0;
#pragma unroll
        for (int o = (NW > 16 ? 16 : NW > 8 ? 8 : NW > 4 ? 4 : NW > 2 ? 2 : 1); o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
    }
    return r;
}

// Two-stage grid reduction with a fixed summation order: every CTA publishes its partial(s); the
// CTA that draws the last ticket folds them (thread-strided, then block_sum) and writes the result.
// No second launch, no atomics on the values, bitwise reproducible for a given grid size.
// Accumulation is fp64 for both entry types; the result is narrowed on the final store.
template <int NRED, typename T, int NT = kBlock>
__device__ __forceinline__ void grid_reduce_finish(const double (&acc)[NRED], double *partials,
                                                   unsigned int *ticket, T *const (&out)[NRED],
                                                   const lsk_peers *peers = nullptr, unsigned long long *reset = nullptr, bool defer = false) {
    __shared__ double s_red[NT / 32];
    __shared__ double s_tot[NRED];
    __shared__ bool s_last;
#pragma unroll
    for (int j = 0; j < NRED; ++j) {
        const double b = block_sum<NT>(acc[j], s_red);
        if (threadIdx.x == 0) partials[(size_t) j * kMaxPartials + blockIdx.x] = b;
    }
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int t = atomicAdd(ticket, 1u);
        s_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
#pragma unroll
    for (int j = 0; j < NRED; ++j) {
        double v = 0.0;
        const volatile double *pj = partials + (size_t) j * kMaxPartials;
        // the fold is done by the first kBlock threads in the 256-thread pattern whatever NT is (same bits)
        if (threadIdx.x < kBlock)
            for (unsigned int i = threadIdx.x; i < gridDim.x; i += kBlock) v += pj[i];
        v = block_sum<NT>(v, s_red);
        if (threadIdx.x == 0) s_tot[j] = v;
    }
    __syncthreads();
    if (peers != nullptr && peers->nranks > 1) {
        // fused all-reduce: the cross-rank sum rides in the tail of the producing kernel (no extra launch)
        if (threadIdx.x < 32) {
            double v[NRED];
#pragma unroll
            for (int j = 0; j < NRED; ++j) v[j] = s_tot[j];
            if (defer) {
                allreduce_send(*peers, v, NRED);  // the consumer kernel forms the cross-rank sum (allreduce_resolve)
            } else {
                allreduce_warp(*peers, v, NRED);
                if (threadIdx.x == 0) {
#pragma unroll
                    for (int j = 0; j < NRED; ++j) s_tot[j] = v[j];
                }
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
#pragma unroll
        for (int j = 0; j < NRED; ++j)
            if (out[j] != nullptr) *out[j] = (T) s_tot[j];
        *ticket = 0u;  // ready for the next launch on this scratch set
        if (reset != nullptr) *reset = 0ull;  // every CTA has drawn its ticket, i.e. is done with the work counter
    }
}

#endif  // __CUDACC__

// ---- programmatic dependent launch (PDL) -- OPT-IN (LSK_PDL=1), measured harmful for the full CG chain ----------
// The three kernels of a CG iteration depend on each other, so every boundary costs the launch latency plus the ramp of
// the next kernel's pipeline.  With PDL the next kernel is scheduled as soon as this one's CTAs have all started
// (pdl_launch_dependents at the top), runs its prologue on SMs as they drain -- barrier set-up, and for the mat-vec the
// rowptr loads and the TMA copy of its first matrix tile, which do not depend on the predecessor -- and blocks in
// pdl_wait until the predecessor has completed and its writes are visible.  Both are no-ops for a kernel launched
// without the attribute / followed by one without it.
// Measured on B200 (256^3, same box, alternating runs): back-to-back mat-vecs gain 1 % (442 -> 439 us) and the CG
// iteration whose x/r update is the grid-stride kernel gains 1 %, but the chain mat-vec -> cg_update_tma ->
// cg_direction_tma with all three edges programmatic drops from 1618 to 1180 it/s (power 630 -> 520 W: the GPU idles
// ~65 us per boundary).  Not understood yet, so the attribute is only set when LSK_PDL=1.
#ifndef LSK_PDL_DEFAULT
#define LSK_PDL_DEFAULT 1  // the mat-vec only (kPdlSpmv): measured +3 % on an 8-GPU slab; the vector kernels lose with it
#endif
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// kernel classes for the LSK_PDL bit mask (developer switch): which launches carry the programmatic attribute
enum PdlClass { kPdlSpmv = 1, kPdlUpdate = 2, kPdlDirection = 4 };
inline int pdl_mask() {
    static const int m = [] { const char *e = getenv("LSK_PDL"); return e ? atoi(e) : LSK_PDL_DEFAULT; }();
    return m;
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(int cls, void (*kernel)(KArgs...), int grid, int block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned) grid);
    cfg.blockDim = dim3((unsigned) block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (pdl_mask() & cls) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#endif

// ---- host-side launch bookkeeping --------------------------------------------------------------------
#define LSK_RETURN_IF_CUDA(expr)                  \
    do {                                          \
        cudaError_t e_ = (expr);                  \
        if (e_ != cudaSuccess) return (int) e_;   \
    } while (0)

// Per-context, per-kernel-family opt-in to more than 48 KB of dynamic shared memory.  `family` indices are handed out
// process-wide on first use (configure_family_index); the "done" bit lives in the context, i.e. per device.
inline int configure_family_index() {
    static int next = 0;
    return next++;
}
template <class Setup>
inline int configure_once(lsk_ctx *ctx, int family, Setup setup) {
    if (family < 0 || family >= 64) return (int) cudaErrorInvalidValue;
    const unsigned long long bit = 1ull << family;
    if (ctx->configured & bit) return 0;
    int current = -1;
    (void) cudaGetDevice(&current);
    if (current != ctx->device) (void) cudaSetDevice(ctx->device);  // function attributes apply to the current device
    const cudaError_t e = setup();
    if (current >= 0 && current != ctx->device) (void) cudaSetDevice(current);
    if (e != cudaSuccess) {
        (void) cudaGetLastError();
        return (int) e;
    }
    ctx->configured |= bit;
    return 0;
}

// one halo move as the entry points accept it; y / n: the vector a fused kernel updates (send ranges are sub-ranges of it), or null
inline bool halo_move_ok(const lsk_halo_move &m, const double *y, int64_t n, int nranks) {
    if (m.peer < 0 || m.peer >= nranks || m.n < 0 || m.recv_n < 0) return false;
    if (!m.ll_send || !m.ll_recv || ((reinterpret_cast<uintptr_t>(m.ll_send) | reinterpret_cast<uintptr_t>(m.ll_recv)) & 15) != 0) return false;
    if (m.n > 0 && (!m.src || (y != nullptr && (m.src < y || m.src + m.n > y + n)))) return false;
    if (m.recv_n > 0 && !m.recv_dst) return false;
    return true;
}
#ifdef __CUDACC__
inline bool halo_spec_fill(HaloSpec &h, const lsk_halo_move *moves, int nmoves, const double *y, int64_t n, int nranks) {
    if (nmoves < 0 || nmoves > kMaxFusedMoves || (nmoves > 0 && !moves)) return false;
    h.nmoves = nmoves;
    for (int i = 0; i < kMaxFusedMoves; ++i) {
        h.lo[i] = 0;
        h.m[i] = lsk_halo_move{};
    }
    for (int i = 0; i < nmoves; ++i) {
        if (!halo_move_ok(moves[i], y, n, nranks)) return false;
        for (int j = 0; j < i; ++j)
            if (moves[j].peer == moves[i].peer) return false;  // exchanges are numbered per pair: one move per peer
        h.m[i] = moves[i];
        h.lo[i] = (moves[i].n > 0 && y != nullptr) ? (int64_t) (moves[i].src - y) : 0;
    }
    return true;
}
#endif

// claim a pending lsk_ctx_defer_next_allreduce request for this launch (only meaningful with peers)
inline bool take_defer(lsk_ctx *ctx) {
    const bool d = ctx->defer_next != 0 && ctx->d_peers != nullptr;
    ctx->defer_next = 0;
    return d;
}
// if a deferred reduction is still in flight and THIS launch is not its designated consumer, finish it first
int settle_pending(lsk_ctx *ctx, cudaStream_t st);  // lsk_comm.cu

inline int after_launch(lsk_ctx *ctx) {
    ctx->launches += 1;
    const cudaError_t e = cudaPeekAtLastError();
    return e == cudaSuccess ? 0 : (int) cudaGetLastError();
}

inline uintptr_t mod32(const void *p) { return (uintptr_t) p & 31u; }

}  // namespace lsk
