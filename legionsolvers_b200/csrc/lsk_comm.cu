// lsk_comm.cu -- the two collectives of the Krylov path over NVLink / NVSwitch peer memory.
//
// One process per GPU; each rank's comm window and its vectors' buffers are mapped into every peer
// with CUDA IPC (host side: host/Runtime.cpp).  Stores to a mapped peer address travel over NVLink
// through the NVSwitch and land in the peer's L2 / HBM; visibility is ordered with
// __threadfence_system() before the epoch-flag store, and the reader polls the flag in its OWN
// memory with volatile (L1-bypassing) loads.  Epochs are monotonic device-side counters, so the
// kernels can be recorded in a CUDA graph and replayed without patching arguments.
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

__global__ void __launch_bounds__(kBlock) halo_exchange_kernel(lsk_peers peers, HaloArgs a) {
    CommWindow *me = static_cast<CommWindow *>(peers.window[peers.rank]);
    __shared__ bool s_last;
    // exchange number of each pair (me, peer): the pair counters are written only by the last CTA, after everyone read them
    // 1. tell every peer I receive from that my ghost region may be overwritten (all my earlier kernels
    //    on this stream -- the readers of the previous ghost values -- have completed)
    if (blockIdx.x == 0 && threadIdx.x < a.nmoves && a.m[threadIdx.x].expect) {
        const int p = a.m[threadIdx.x].peer;
        CommWindow *dst = static_cast<CommWindow *>(peers.window[p]);
        *reinterpret_cast<volatile unsigned long long *>(&dst->halo_ready[peers.rank]) = me->halo_sent[p] + 1;
    }
    // 2. wait until every peer I send to is ready
    if (threadIdx.x < a.nmoves && a.m[threadIdx.x].n > 0) {
        const int p = a.m[threadIdx.x].peer;
        spin_until(&me->halo_ready[p], me->halo_sent[p] + 1, &me->error);
    }
    __syncthreads();
    // 3. store my boundary values straight into the peers' ghost regions
    for (int i = 0; i < a.nmoves; ++i) {
        const double *src = a.m[i].src;
        double *dst = a.m[i].dst;
        const int64_t n = a.m[i].n;
        const bool vec = ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0;
        const int64_t tid = (int64_t) blockIdx.x * kBlock + threadIdx.x, stride = (int64_t) gridDim.x * kBlock;
        if (vec) {
            const int64_t n2 = n >> 1;
            for (int64_t k = tid; k < n2; k += stride)
                reinterpret_cast<double2 *>(dst)[k] = reinterpret_cast<const double2 *>(src)[k];
            if (tid == 0 && (n & 1)) dst[n - 1] = src[n - 1];
        } else {
            for (int64_t k = tid; k < n; k += stride) dst[k] = src[k];
        }
    }
    // 4. the last CTA to finish publishes "done" to the peers, waits for theirs, and advances the pair counters
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int t = atomicAdd(&me->halo_ticket, 1u);
        s_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;
    halo_publish(&peers, a.m, a.nmoves, true);
    if (threadIdx.x == 0) me->halo_ticket = 0u;
}

// Closes an OPEN exchange (halo_publish without wait) for consumers that cannot wait per row block: one thread per
// peer this rank receives from blocks until that peer's data of the pair's latest exchange has landed.
__global__ void __launch_bounds__(32) halo_wait_kernel(lsk_peers peers, HaloArgs a) {
    CommWindow *me = static_cast<CommWindow *>(peers.window[peers.rank]);
    if (threadIdx.x < a.nmoves && a.m[threadIdx.x].expect) {
        const int p = a.m[threadIdx.x].peer;
        spin_until(&me->halo_done[p], me->halo_sent[p], &me->error);
    }
    __syncwarp();
    __threadfence_system();
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

int lsk_halo_exchange_f64(lsk_ctx *ctx, lsk_stream s, const lsk_peers *peers, const lsk_halo_move *moves, int nmoves) {
    if (!ctx || !peers_ok(peers) || nmoves < 0 || nmoves > LSK_MAX_HALO_MOVES || (nmoves > 0 && !moves)) return LSK_E_INVALID;
    if (nmoves == 0) return 0;
    HaloArgs a;
    a.nmoves = nmoves;
    int64_t total = 0;
    for (int i = 0; i < nmoves; ++i) {
        a.m[i] = moves[i];
        if (moves[i].peer < 0 || moves[i].peer >= peers->nranks || moves[i].n < 0) return LSK_E_INVALID;
        if (moves[i].n > 0 && (!moves[i].src || !moves[i].dst)) return LSK_E_INVALID;
        for (int j = 0; j < i; ++j)
            if (moves[j].peer == moves[i].peer) return LSK_E_INVALID;  // exchanges are numbered per pair: one move per peer
        total += moves[i].n;
    }
    // enough CTAs to keep NVLink busy for a few hundred KB, few enough that the epilogue stays cheap
    int grid = (int) ((total + 4 * kBlock - 1) / (4 * kBlock));
    if (grid < 1) grid = 1;
    if (grid > 32) grid = 32;
    halo_exchange_kernel<<<grid, kBlock, 0, (cudaStream_t) s>>>(*peers, a);
    return after_launch(ctx);
}

int lsk_halo_wait_f64(lsk_ctx *ctx, lsk_stream s, const lsk_peers *peers, const lsk_halo_move *moves, int nmoves) {
    if (!ctx || !peers_ok(peers) || nmoves < 0 || nmoves > 32 || (nmoves > 0 && !moves)) return LSK_E_INVALID;
    if (nmoves == 0) return 0;
    HaloArgs a;
    a.nmoves = nmoves;
    for (int i = 0; i < nmoves; ++i) {
        if (moves[i].peer < 0 || moves[i].peer >= peers->nranks) return LSK_E_INVALID;
        a.m[i] = moves[i];
    }
    halo_wait_kernel<<<1, 32, 0, (cudaStream_t) s>>>(*peers, a);
    return after_launch(ctx);
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
