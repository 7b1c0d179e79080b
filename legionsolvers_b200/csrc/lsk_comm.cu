// lsk_comm.cu -- the two collectives of the Krylov path over NVLink / NVSwitch peer memory.
//
// One process per GPU; each rank's comm window and its halo landing buffers are mapped into every peer
// with CUDA IPC (host side: host/Runtime.cpp).  Stores to a mapped peer address travel over NVLink
// through the NVSwitch and land in the peer's L2 / HBM.  Everything that crosses the link is an LL
// packet: 32 data bits + the 32-bit number of the exchange in one atomic 8-byte word, so there is no
// fence and no flag anywhere (lsk_common.cuh); the reader polls the packets in its OWN memory with
// volatile (L1-bypassing) loads.  Exchange numbers are monotonic device-side counters, so the kernels
// can be recorded in a CUDA graph and replayed without patching arguments.
#include "lsk_common.cuh"

namespace lsk {

// ---- all-reduce: one CTA, thread r talks to rank r ----------------------------------------------------
__global__ void __launch_bounds__(32) allreduce_kernel(lsk_peers peers, double *slots, int count) {
    double v[kMaxRed];
    for (int j = 0; j < kMaxRed; ++j) v[j] = j < count ? slots[j] : 0.0;
    allreduce_warp(peers, v, count);
    if (threadIdx.x == 0)
        for (int j = 0; j < count; ++j) slots[j] = v[j];
}

// finishes a deferred all-reduce whose designated consumer did not come: one warp polls, sums, stores the slot
__global__ void __launch_bounds__(32) allreduce_resolve_kernel(const lsk_peers *peers, double *slot) {
    __shared__ double s_v[kMaxRed];
    allreduce_resolve(*peers, s_v, 1, slot);
}

int settle_pending(lsk_ctx *ctx, cudaStream_t st) {
    if (ctx->pending_slot == nullptr) return 0;
    double *slot = const_cast<double *>(static_cast<const double *>(ctx->pending_slot));
    ctx->pending_slot = nullptr;
    if (ctx->d_peers == nullptr) return 0;
    allreduce_resolve_kernel<<<1, 32, 0, st>>>(ctx->d_peers, slot);
    return after_launch(ctx);
}

// ---- halo exchange -------------------------------------------------------------------------------------
struct HaloArgs {
    int nmoves;
    lsk_halo_move m[LSK_MAX_HALO_MOVES];
};

// Stand-alone exchange (the fused forms live in lsk_blas1.cu): send my boundary values as packets, unpack the
// neighbours'.  At most 32 CTAs: all resident, so a CTA polling for a packet never keeps a sender off the SMs.
template <bool ADD>
__global__ void __launch_bounds__(kBlock) halo_exchange_kernel(lsk_peers peers, HaloArgs a) {
    CommWindow *me = static_cast<CommWindow *>(peers.window[peers.rank]);
    __shared__ bool s_last;
    const int64_t tid = (int64_t) blockIdx.x * kBlock + threadIdx.x, stride = (int64_t) gridDim.x * kBlock;
    for (int i = 0; i < a.nmoves; ++i) {
        HaloLive one;
        halo_live_move(one, 0, a.m[i], me);
        const double *src = a.m[i].src;
        const int64_t n = a.m[i].n;
        if ((reinterpret_cast<uintptr_t>(one.send_slot[0]) & 31) == 0) {  // two packets per 32-byte store, a warp writes whole sectors
            for (int64_t k = 2 * tid; k + 1 < n; k += 2 * stride) ll_store2(one.send_slot[0], k, src[k], src[k + 1], one.tag[0]);
            if (tid == 0 && (n & 1)) ll_store(one.send_slot[0], n - 1, src[n - 1], one.tag[0]);
        } else {
            for (int64_t k = tid; k < n; k += stride) ll_store(one.send_slot[0], k, src[k], one.tag[0]);
        }
        if (tid == 0) ll_store(one.send_slot[0], n, 0.0, one.tag[0]);  // the token
    }
    for (int i = 0; i < a.nmoves; ++i) {
        HaloLive one;
        halo_live_move(one, 0, a.m[i], me);
        halo_unpack<ADD>(one, &a.m[i], 1, &peers);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int t = atomicAdd(&me->halo_ticket, 1u);
        s_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;
    for (int i0 = 0; i0 < a.nmoves; i0 += kBlock) {
        const int i = i0 + (int) threadIdx.x;
        if (i < a.nmoves) me->halo_sent[a.m[i].peer] += 1;
    }
    if (threadIdx.x == 0) {
        me->halo_calls += 1;
        me->halo_wait_ns += me->halo_poll_ns;
        me->halo_poll_ns = 0;
        me->halo_ticket = 0u;
    }
}

}  // namespace lsk

using namespace lsk;

extern "C" {

size_t lsk_comm_window_bytes(void) { return (sizeof(CommWindow) + 255) & ~size_t(255); }

static bool peers_ok(const lsk_peers *p) {
    if (!p || p->nranks < 1 || p->nranks > LSK_MAX_RANKS || p->rank < 0 || p->rank >= p->nranks) return false;
    for (int r = 0; r < p->nranks; ++r)
        if (!p->window[r]) return false;
    return true;
}

int lsk_allreduce_sum_f64(lsk_ctx *ctx, lsk_stream s, const lsk_peers *peers, double *slots, int count) {
    if (!ctx || !slots || count < 1 || count > kMaxRed || !peers_ok(peers)) return LSK_E_INVALID;
    if (peers->nranks > 32) return LSK_E_INVALID;
    allreduce_kernel<<<1, 32, 0, (cudaStream_t) s>>>(*peers, slots, count);
    return after_launch(ctx);
}

size_t lsk_halo_landing_bytes(int64_t count) { return count < 0 ? 0 : 2 * ll_half_bytes(count); }

static int halo_exchange(lsk_ctx *ctx, lsk_stream s, const lsk_peers *peers, const lsk_halo_move *moves, int nmoves, bool add) {
    if (!ctx || !peers_ok(peers) || nmoves < 0 || nmoves > LSK_MAX_HALO_MOVES || (nmoves > 0 && !moves)) return LSK_E_INVALID;
    if (nmoves == 0) return 0;
    HaloArgs a;
    a.nmoves = nmoves;
    int64_t total = 0;
    for (int i = 0; i < nmoves; ++i) {
        a.m[i] = moves[i];
        if (!halo_move_ok(moves[i], nullptr, 0, peers->nranks) || moves[i].peer == peers->rank) return LSK_E_INVALID;
        for (int j = 0; j < i; ++j)
            if (moves[j].peer == moves[i].peer) return LSK_E_INVALID;  // exchanges are numbered per pair: one move per peer
        total += moves[i].n + moves[i].recv_n;
    }
    // enough CTAs to keep NVLink busy for a few hundred KB, few enough that all of them are resident
    int grid = (int) ((total + 4 * kBlock - 1) / (4 * kBlock));
    if (grid < 1) grid = 1;
    if (grid > 32) grid = 32;
    if (add) halo_exchange_kernel<true><<<grid, kBlock, 0, (cudaStream_t) s>>>(*peers, a);
    else halo_exchange_kernel<false><<<grid, kBlock, 0, (cudaStream_t) s>>>(*peers, a);
    return after_launch(ctx);
}

int lsk_halo_exchange_f64(lsk_ctx *ctx, lsk_stream s, const lsk_peers *peers, const lsk_halo_move *moves, int nmoves) {
    return halo_exchange(ctx, s, peers, moves, nmoves, false);
}
int lsk_halo_reduce_f64(lsk_ctx *ctx, lsk_stream s, const lsk_peers *peers, const lsk_halo_move *moves, int nmoves) {
    return halo_exchange(ctx, s, peers, moves, nmoves, true);
}

int lsk_comm_stats(lsk_ctx *ctx, lsk_stream s, const lsk_peers *peers, uint64_t *host_out4) {
    if (!ctx || !peers_ok(peers) || !host_out4) return LSK_E_INVALID;
    const CommWindow *me = static_cast<const CommWindow *>(peers->window[peers->rank]);
    LSK_RETURN_IF_CUDA(cudaMemcpyAsync(host_out4, &me->ar_calls, 4 * sizeof(uint64_t), cudaMemcpyDeviceToHost, (cudaStream_t) s));
    LSK_RETURN_IF_CUDA(cudaStreamSynchronize((cudaStream_t) s));
    return 0;
}

int lsk_comm_error(lsk_ctx *ctx, lsk_stream s, const lsk_peers *peers, int *host_out) {
    if (!ctx || !peers_ok(peers) || !host_out) return LSK_E_INVALID;
    const CommWindow *me = static_cast<const CommWindow *>(peers->window[peers->rank]);
    LSK_RETURN_IF_CUDA(cudaMemcpyAsync(host_out, &me->error, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t) s));
    LSK_RETURN_IF_CUDA(cudaStreamSynchronize((cudaStream_t) s));
    return 0;
}

}  // extern "C"
