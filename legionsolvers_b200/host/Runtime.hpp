// Runtime.hpp -- the minimal region / partition / future shim under the host layer.
//
// The reference runs on Legion + Realm (absent here).  What the Krylov path actually needs from that
// runtime is small, and on one NVSwitch box maps onto CUDA + NCCL directly:
//
//   Legion / Realm concept (reference)                      here
//   -----------------------------------------------------   -----------------------------------------
//   one TOC_PROC + its Realm task stream                    one process per GPU, one CUDA stream
//   CUDALibraryContext (src/CUDAUtilities.hpp:44-66)        lsk_ctx (reduction scratch)
//   Future / Scalar<T> (src/Scalar.hpp)                     8-byte slot in a device-resident arena
//   FutureMap sum reduction across pieces / shards          colour-order fold on device + ncclAllReduce
//   ghost-region instance copies (implicit Realm DMA)       grouped ncclSend/ncclRecv of halo intervals
//   begin_trace / end_trace + memoize (BenchmarkStencil)    CUDA-graph capture on first use, replay after
//   BlockingShardingFunctor (LegionSolversMapper.cpp:140)   colour c -> rank c / ceil(P / nranks)
//
// Errors: the reference prints and aborts (CHECK_* -> assert(false)); this layer throws
// std::runtime_error, which the C ABI in lsk_solvers.cpp turns into a status + message.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <functional>
#include <map>
#include <memory>
#include <set>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/lsk.h"

struct ncclComm;

namespace LegionSolvers {

class Runtime {
public:
    // `external_stream` (may be null): enqueue on the caller's stream instead of a private one
    Runtime(int device, int rank, int nranks, void *external_stream = nullptr);
    ~Runtime();
    Runtime(const Runtime &) = delete;
    Runtime &operator=(const Runtime &) = delete;

    lsk_ctx *ctx() const { return ctx_; }
    cudaStream_t stream() const { return stream_; }
    int device() const { return device_; }
    int rank() const { return rank_; }
    int nranks() const { return nranks_; }

    // ---- communication ---------------------------------------------------------------------------
    static void comm_unique_id(void *out128);
    void comm_init(const void *uid128);
    bool has_comm() const { return comm_ != nullptr; }
    void allreduce_sum(double *slots, int count);  // in place, stream-ordered; no-op on one rank
    // ---- peer memory (CUDA IPC over NVLink): the low-latency path of the two collectives -----------
    struct Exported {            // one cudaMalloc'ed allocation mapped into every peer
        std::vector<char *> base;     // base[r] = rank r's allocation as mapped here (base[rank] = local)
        std::vector<int64_t> tag0, tag1;  // two user values per rank (e.g. byte offset of element 0, buf_lo)
    };
    bool p2p() const { return p2p_; }
    const lsk_peers &peers() const { return peers_; }
    // COLLECTIVE (every rank, same order): export `raw` and map everybody else's
    Exported export_allocation(void *raw, int64_t tag0, int64_t tag1);
    void halo_exchange_p2p(const lsk_halo_move *moves, int nmoves);
    void halo_reduce_p2p(const lsk_halo_move *moves, int nmoves);  // values received are added (lsk_halo_reduce_f64)
    // Fused collectives: every reducing kernel finishes with the cross-rank sum in its own tail, and
    // lsk_xpay_halo_f64 may be used.  Valid only while every rank launches exactly the same reducing
    // kernels (one local piece per rank); the planner switches it on when that holds.
    void set_fused_collectives(bool on);
    bool fused_collectives() const { return fused_; }
    int comm_error();   // non-zero if a spin-wait of a collective gave up
    void allgather_i64(const int64_t *send_dev, int64_t *recv_dev, int count_per_rank);
    void group_start();
    void group_end();
    void send(const void *ptr, size_t bytes, int peer);
    void recv(void *ptr, size_t bytes, int peer);

    // ---- device memory / scalar arena ---------------------------------------------------------------
    void *alloc(size_t bytes);
    void free(void *p);
    double *new_slot();  // 8-byte device slot, zero-initialised, lives as long as the runtime

    // ---- tracing: the begin_trace/end_trace of test/BenchmarkStencil.cpp:219-241 -------------------
    void begin_trace(int id);
    void end_trace(int id);
    bool replaying() const { return mode_ == Mode::Replay; }
    bool capturing() const { return mode_ == Mode::Capture; }

    // Every device operation of the host layer goes through here: skipped while a recorded trace
    // is being replayed (the graph launch at end_trace does the work), checked otherwise.
    template <class F>
    void enqueue(const char *what, F &&f) {
        if (mode_ == Mode::Replay) return;
        const int rc = f();
        if (rc != 0) fail(rc, what);
    }

    void fence();  // issue_execution_fence: wait for the stream
    uint64_t kernel_launches() const;  // lsk kernels launched, graph replays included

    [[noreturn]] void fail(int status, const char *what) const;
    void check_cuda(cudaError_t e, const char *what) const {
        if (e != cudaSuccess) fail((int) e, what);
    }

private:
    enum class Mode { Eager, Capture, Replay };
    struct Trace {
        cudaGraphExec_t exec = nullptr;
        uint64_t kernels = 0;
    };
    int device_, rank_, nranks_;
    lsk_ctx *ctx_ = nullptr;
    cudaStream_t stream_ = nullptr;
    bool own_stream_ = false;
    ncclComm *comm_ = nullptr;
    bool p2p_ = false;
    bool fused_ = false;
    lsk_peers peers_{};
    void *window_ = nullptr;
    std::vector<void *> ipc_opened_;
    std::map<std::string, void *> ipc_cache_;
    Mode mode_ = Mode::Eager;
    int active_trace_ = -1;
    bool eager_trace_ = false;
    uint64_t capture_mark_ = 0;
    uint64_t replayed_kernels_ = 0;
    std::map<int, Trace> traces_;
    std::vector<void *> allocations_;
    std::set<void *> exported_;     // allocations some peer may have mapped (CUDA IPC)
    std::vector<void *> retired_;   // exported allocations released by their owner: freed at the collective teardown
    std::vector<double *> arena_chunks_;
    size_t arena_used_ = 0;
    static constexpr size_t kArenaChunk = 1 << 16;
};

// RAII device buffer
template <typename T>
struct DeviceBuffer {
    Runtime *rt = nullptr;
    T *ptr = nullptr;
    size_t count = 0;
    DeviceBuffer() = default;
    DeviceBuffer(Runtime *rt_, size_t n) : rt(rt_), ptr(n ? static_cast<T *>(rt_->alloc(n * sizeof(T))) : nullptr), count(n) {}
    DeviceBuffer(const DeviceBuffer &) = delete;
    DeviceBuffer &operator=(const DeviceBuffer &) = delete;
    DeviceBuffer(DeviceBuffer &&o) noexcept : rt(o.rt), ptr(o.ptr), count(o.count) { o.ptr = nullptr; o.count = 0; }
    DeviceBuffer &operator=(DeviceBuffer &&o) noexcept {
        if (this != &o) {
            release();
            rt = o.rt; ptr = o.ptr; count = o.count;
            o.ptr = nullptr; o.count = 0;
        }
        return *this;
    }
    ~DeviceBuffer() { release(); }
    void release() {
        if (ptr && rt) rt->free(ptr);
        ptr = nullptr;
        count = 0;
    }
};

}  // namespace LegionSolvers
