// lsk_ctx.cu -- per-GPU context: the replacement for the reference's CUDALibraryContext /
// LoadCUDALibsTask (src/CUDAUtilities.cpp:66-145, src/CudaLibs.cu:11-66).  It owns the only
// persistent device state the kernels need: rotating reduction scratch, ticket counters and three
// constant scalars.  Created once per GPU; never touched by the kernels' callers afterwards.
#include <new>

#include "lsk_common.cuh"

using namespace lsk;

extern "C" {

int lsk_version(void) { return LSK_VERSION; }

const char *lsk_error_string(int status) {
    if (status == 0) return "success";
    if (status > 0) return cudaGetErrorString((cudaError_t) status);
    switch (status) {
    case LSK_E_INVALID: return "lsk: invalid argument";
    case LSK_E_NO_DEVICE: return "lsk: no CUDA device (there is no CPU fallback)";
    case LSK_E_CAPACITY: return "lsk: context scratch capacity exceeded";
    case LSK_E_NCCL: return "lsk: NCCL failure";
    default: return "lsk: unknown error";
    }
}

int lsk_ctx_create(int device, lsk_ctx **out) {
    if (!out) return LSK_E_INVALID;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) {
        (void) cudaGetLastError();
        return LSK_E_NO_DEVICE;
    }
    if (device < 0 || device >= count) return LSK_E_INVALID;
    LSK_RETURN_IF_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    LSK_RETURN_IF_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return LSK_E_NO_DEVICE;  // built for sm_100a only

    lsk_ctx *ctx = new (std::nothrow) lsk_ctx();
    if (!ctx) return LSK_E_CAPACITY;
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->cursor = 0;
    ctx->launches = 0;
    ctx->d_peers = nullptr;
    ctx->work = nullptr;
    ctx->tail_sync = nullptr;
    ctx->configured = 0;
    ctx->defer_next = 0;
    ctx->pending_slot = nullptr;
    const size_t pbytes = sizeof(double) * (size_t) kScratchSets * kMaxRed * kMaxPartials;
    cudaError_t e = cudaMalloc(&ctx->partials, pbytes);
    if (e == cudaSuccess) e = cudaMalloc(&ctx->tickets, sizeof(unsigned int) * kScratchSets);
    if (e == cudaSuccess) e = cudaMalloc(&ctx->consts, sizeof(double) * 4);
    if (e == cudaSuccess) e = cudaMalloc(&ctx->work, sizeof(unsigned long long) * kScratchSets);
    if (e == cudaSuccess) e = cudaMemset(ctx->work, 0, sizeof(unsigned long long) * kScratchSets);
    if (e == cudaSuccess) e = cudaMalloc(&ctx->tail_sync, 256);
    if (e == cudaSuccess) e = cudaMemset(ctx->tail_sync, 0, 256);
    if (e == cudaSuccess) e = cudaMemset(ctx->partials, 0, pbytes);
    if (e == cudaSuccess) e = cudaMemset(ctx->tickets, 0, sizeof(unsigned int) * kScratchSets);
    const double consts[4] = {1.0, -1.0, 0.0, 0.0};
    if (e == cudaSuccess) e = cudaMemcpy(ctx->consts, consts, sizeof(consts), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        lsk_ctx_destroy(ctx);
        return (int) e;
    }
    *out = ctx;
    return 0;
}

int lsk_ctx_destroy(lsk_ctx *ctx) {
    if (!ctx) return 0;
    cudaSetDevice(ctx->device);
    if (ctx->partials) cudaFree(ctx->partials);
    if (ctx->tickets) cudaFree(ctx->tickets);
    if (ctx->consts) cudaFree(ctx->consts);
    if (ctx->d_peers) cudaFree(ctx->d_peers);
    if (ctx->work) cudaFree(ctx->work);
    if (ctx->tail_sync) cudaFree(ctx->tail_sync);
    delete ctx;
    return 0;
}

int lsk_ctx_set_peers(lsk_ctx *ctx, const lsk_peers *peers) {
    if (!ctx) return LSK_E_INVALID;
    LSK_RETURN_IF_CUDA(cudaSetDevice(ctx->device));
    if (!peers) {
        if (ctx->d_peers) {
            LSK_RETURN_IF_CUDA(cudaDeviceSynchronize());
            cudaFree(ctx->d_peers);
            ctx->d_peers = nullptr;
        }
        return 0;
    }
    if (peers->nranks < 1 || peers->nranks > LSK_MAX_RANKS || peers->rank < 0 || peers->rank >= peers->nranks) return LSK_E_INVALID;
    if (!ctx->d_peers) LSK_RETURN_IF_CUDA(cudaMalloc(&ctx->d_peers, sizeof(lsk_peers)));
    LSK_RETURN_IF_CUDA(cudaMemcpy(ctx->d_peers, peers, sizeof(lsk_peers), cudaMemcpyHostToDevice));
    ctx->h_peers = *peers;
    return 0;
}

int lsk_ctx_defer_next_allreduce(lsk_ctx *ctx) {
    if (!ctx) return LSK_E_INVALID;
    ctx->defer_next = ctx->d_peers != nullptr ? 1 : 0;
    return 0;
}
int lsk_ctx_settle(lsk_ctx *ctx, lsk_stream s) {
    if (!ctx) return LSK_E_INVALID;
    return lsk::settle_pending(ctx, (cudaStream_t) s);
}

int lsk_ctx_device(const lsk_ctx *ctx) { return ctx ? ctx->device : -1; }
int lsk_ctx_sm_count(const lsk_ctx *ctx) { return ctx ? ctx->sm_count : 0; }
uint64_t lsk_ctx_launch_count(const lsk_ctx *ctx) { return ctx ? ctx->launches : 0; }
const double *lsk_ctx_const_f64(const lsk_ctx *ctx, int which) {
    if (!ctx || which < 0 || which > 2) return nullptr;
    return ctx->consts + which;
}

}  // extern "C"
