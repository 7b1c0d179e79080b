// Runtime.cpp -- see Runtime.hpp.
#include "Runtime.hpp"

#include <nccl.h>

#include <algorithm>
#include <cstring>
#include <cstdlib>
#include <set>
#include <sstream>

namespace LegionSolvers {

Runtime::Runtime(int device, int rank, int nranks, void *external_stream)
    : device_(device), rank_(rank), nranks_(nranks) {
    if (nranks < 1 || rank < 0 || rank >= nranks) throw std::runtime_error("Runtime: bad rank / nranks");
    const int rc = lsk_ctx_create(device, &ctx_);
    if (rc != 0) fail(rc, "lsk_ctx_create");
    check_cuda(cudaSetDevice(device), "cudaSetDevice");
    if (external_stream) {
        stream_ = static_cast<cudaStream_t>(external_stream);
    } else {
        check_cuda(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking), "cudaStreamCreate");
        own_stream_ = true;
    }
    (void) new_slot();  // first arena chunk up front, so that traces never have to grow it
}

Runtime::~Runtime() {
    cudaSetDevice(device_);
    cudaStreamSynchronize(stream_);
    for (auto &kv : traces_)
        if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
    // Exported allocations must outlive every importer's mapping (freeing IPC-exported memory that a peer still has
    // open is undefined): close MY mappings of the peers' buffers, then meet the peers at a barrier -- after it, nobody
    // maps my buffers any more -- and only then free.
    for (void *p : ipc_opened_) cudaIpcCloseMemHandle(p);
    if (comm_ && p2p_ && nranks_ > 1) {
        int *flag = nullptr;
        if (cudaMalloc(&flag, sizeof(int)) == cudaSuccess) {
            cudaMemsetAsync(flag, 0, sizeof(int), stream_);
            if (ncclAllReduce(flag, flag, 1, ncclInt, ncclSum, reinterpret_cast<ncclComm_t>(comm_), stream_) == ncclSuccess)
                cudaStreamSynchronize(stream_);
            cudaFree(flag);
        }
    }
    for (void *p : retired_) cudaFree(p);
    if (comm_) ncclCommDestroy(reinterpret_cast<ncclComm_t>(comm_));
    if (window_) cudaFree(window_);
    for (void *p : allocations_) cudaFree(p);
    for (double *p : arena_chunks_) cudaFree(p);
    if (own_stream_) cudaStreamDestroy(stream_);
    lsk_ctx_destroy(ctx_);
}

void Runtime::fail(int status, const char *what) const {
    std::ostringstream os;
    os << "[LegionSolvers] " << what << " failed on rank " << rank_ << ": " << lsk_error_string(status) << " ("
       << status << ")";
    throw std::runtime_error(os.str());
}

// ---- communication ---------------------------------------------------------------------------------
static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is passed through the C ABI as 128 bytes");

void Runtime::comm_unique_id(void *out128) {
    ncclUniqueId id;
    if (ncclGetUniqueId(&id) != ncclSuccess) throw std::runtime_error("ncclGetUniqueId failed");
    std::memcpy(out128, &id, sizeof(id));
}

void Runtime::comm_init(const void *uid128) {
    if (nranks_ == 1) return;
    ncclUniqueId id;
    std::memcpy(&id, uid128, sizeof(id));
    check_cuda(cudaSetDevice(device_), "cudaSetDevice");
    ncclComm_t c;
    if (ncclCommInitRank(&c, nranks_, id, rank_) != ncclSuccess) fail(LSK_E_NCCL, "ncclCommInitRank");
    comm_ = reinterpret_cast<ncclComm *>(c);
    // Peer-memory windows for the latency-bound collectives.  LSK_COMM=nccl keeps everything on NCCL.
    const char *mode = std::getenv("LSK_COMM");
    if (mode && std::string(mode) == "nccl") return;
    if (nranks_ > LSK_MAX_RANKS) return;
    const size_t wbytes = lsk_comm_window_bytes();
    check_cuda(cudaMalloc(&window_, wbytes), "cudaMalloc(comm window)");
    check_cuda(cudaMemset(window_, 0, wbytes), "cudaMemset(comm window)");
    try {
        const Exported ex = export_allocation(window_, 0, 0);
        peers_.rank = rank_;
        peers_.nranks = nranks_;
        for (int r = 0; r < nranks_; ++r) peers_.window[r] = ex.base[(size_t) r];
        p2p_ = true;
    } catch (const std::exception &) {
        p2p_ = false;  // no IPC / no peer access: stay on NCCL
    }
    // all ranks must agree (a rank falling back alone would deadlock the others)
    DeviceBuffer<int64_t> flag(this, 1), all(this, (size_t) nranks_);
    const int64_t ok = p2p_ ? 1 : 0;
    check_cuda(cudaMemcpyAsync(flag.ptr, &ok, sizeof(ok), cudaMemcpyHostToDevice, stream_), "p2p vote");
    allgather_i64(flag.ptr, all.ptr, 1);
    std::vector<int64_t> votes((size_t) nranks_);
    check_cuda(cudaMemcpyAsync(votes.data(), all.ptr, sizeof(int64_t) * votes.size(), cudaMemcpyDeviceToHost, stream_), "p2p vote");
    fence();
    for (int64_t v : votes) p2p_ = p2p_ && (v == 1);
}

Runtime::Exported Runtime::export_allocation(void *raw, int64_t tag0, int64_t tag1) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle travels as 8 int64");
    constexpr int W = 10;  // 8 words of handle + 2 tags
    int64_t mine[W];
    cudaIpcMemHandle_t h;
    const cudaError_t ge = cudaIpcGetMemHandle(&h, raw);
    std::memset(mine, 0, sizeof(mine));
    if (ge == cudaSuccess) std::memcpy(mine, &h, sizeof(h));
    else (void) cudaGetLastError();
    mine[8] = tag0;
    mine[9] = tag1;
    DeviceBuffer<int64_t> send(this, W), recv(this, (size_t) W * nranks_);
    check_cuda(cudaMemcpyAsync(send.ptr, mine, sizeof(mine), cudaMemcpyHostToDevice, stream_), "export H2D");
    allgather_i64(send.ptr, recv.ptr, W);
    std::vector<int64_t> all((size_t) W * nranks_);
    check_cuda(cudaMemcpyAsync(all.data(), recv.ptr, sizeof(int64_t) * all.size(), cudaMemcpyDeviceToHost, stream_), "export D2H");
    fence();
    if (ge == cudaSuccess) exported_.insert(raw);
    Exported ex;
    ex.base.assign((size_t) nranks_, nullptr);
    ex.tag0.assign((size_t) nranks_, 0);
    ex.tag1.assign((size_t) nranks_, 0);
    bool ok = (ge == cudaSuccess);
    for (int r = 0; r < nranks_; ++r) {
        ex.tag0[(size_t) r] = all[(size_t) r * W + 8];
        ex.tag1[(size_t) r] = all[(size_t) r * W + 9];
        if (r == rank_) {
            ex.base[(size_t) r] = static_cast<char *>(raw);
            continue;
        }
        cudaIpcMemHandle_t ph;
        std::memcpy(&ph, &all[(size_t) r * W], sizeof(ph));
        const std::string key(reinterpret_cast<const char *>(&ph), sizeof(ph));
        auto hit = ipc_cache_.find(key);
        if (hit != ipc_cache_.end()) {  // the peer re-exported an allocation that is already mapped here
            ex.base[(size_t) r] = static_cast<char *>(hit->second);
            continue;
        }
        void *mapped = nullptr;
        const cudaError_t oe = cudaIpcOpenMemHandle(&mapped, ph, cudaIpcMemLazyEnablePeerAccess);
        if (oe != cudaSuccess) {
            (void) cudaGetLastError();
            ok = false;
            continue;
        }
        ipc_opened_.push_back(mapped);
        ipc_cache_[key] = mapped;
        ex.base[(size_t) r] = static_cast<char *>(mapped);
    }
    if (!ok) throw std::runtime_error("[LegionSolvers] CUDA IPC export failed (no peer access between ranks?)");
    return ex;
}

void Runtime::halo_exchange_p2p(const lsk_halo_move *moves, int nmoves) {
    enqueue("halo exchange", [&] { return lsk_halo_exchange_f64(ctx_, stream_, &peers_, moves, nmoves); });
}

void Runtime::halo_reduce_p2p(const lsk_halo_move *moves, int nmoves) {
    enqueue("halo reduce", [&] { return lsk_halo_reduce_f64(ctx_, stream_, &peers_, moves, nmoves); });
}

void Runtime::set_fused_collectives(bool on) {
    if (on && !p2p_) return;
    if (on == fused_) return;
    if (mode_ != Mode::Eager) fail(LSK_E_INVALID, "set_fused_collectives inside a trace");
    const char *mode = std::getenv("LSK_COMM");
    if (on && mode && std::string(mode) == "p2p-unfused") return;  // developer A/B switch: standalone collective kernels
    const int rc = lsk_ctx_set_peers(ctx_, on ? &peers_ : nullptr);
    if (rc != 0) fail(rc, "lsk_ctx_set_peers");
    fused_ = on;
}

int Runtime::comm_error() {
    int e = 0;
    if (!p2p_) return 0;
    const int rc = lsk_comm_error(ctx_, stream_, &peers_, &e);
    if (rc != 0) fail(rc, "lsk_comm_error");
    return e;
}

#define LSK_NCCL(expr, what)                              \
    do {                                                  \
        if ((expr) != ncclSuccess) fail(LSK_E_NCCL, what); \
    } while (0)

void Runtime::allreduce_sum(double *slots, int count) {
    if (nranks_ == 1 || mode_ == Mode::Replay) return;
    if (fused_) fail(LSK_E_INVALID, "stand-alone all-reduce while reductions are fused (would double count)");
    if (p2p_) {
        enqueue("allreduce", [&] { return lsk_allreduce_sum_f64(ctx_, stream_, &peers_, slots, count); });
        return;
    }
    if (!comm_) fail(LSK_E_NCCL, "allreduce_sum without comm_init");
    LSK_NCCL(ncclAllReduce(slots, slots, (size_t) count, ncclDouble, ncclSum, reinterpret_cast<ncclComm_t>(comm_), stream_),
             "ncclAllReduce");
}

void Runtime::allgather_i64(const int64_t *send_dev, int64_t *recv_dev, int count_per_rank) {
    if (nranks_ == 1) {
        check_cuda(cudaMemcpyAsync(recv_dev, send_dev, sizeof(int64_t) * (size_t) count_per_rank,
                                   cudaMemcpyDeviceToDevice, stream_), "allgather copy");
        return;
    }
    if (!comm_) fail(LSK_E_NCCL, "allgather without comm_init");
    LSK_NCCL(ncclAllGather(send_dev, recv_dev, (size_t) count_per_rank, ncclInt64, reinterpret_cast<ncclComm_t>(comm_), stream_),
             "ncclAllGather");
}

void Runtime::group_start() {
    if (nranks_ > 1 && mode_ != Mode::Replay) LSK_NCCL(ncclGroupStart(), "ncclGroupStart");
}
void Runtime::group_end() {
    if (nranks_ > 1 && mode_ != Mode::Replay) LSK_NCCL(ncclGroupEnd(), "ncclGroupEnd");
}
void Runtime::send(const void *ptr, size_t bytes, int peer) {
    if (mode_ == Mode::Replay) return;
    LSK_NCCL(ncclSend(ptr, bytes, ncclChar, peer, reinterpret_cast<ncclComm_t>(comm_), stream_), "ncclSend");
}
void Runtime::recv(void *ptr, size_t bytes, int peer) {
    if (mode_ == Mode::Replay) return;
    LSK_NCCL(ncclRecv(ptr, bytes, ncclChar, peer, reinterpret_cast<ncclComm_t>(comm_), stream_), "ncclRecv");
}

// ---- memory ----------------------------------------------------------------------------------------------
void *Runtime::alloc(size_t bytes) {
    void *p = nullptr;
    check_cuda(cudaSetDevice(device_), "cudaSetDevice");
    check_cuda(cudaMalloc(&p, bytes ? bytes : 1), "cudaMalloc");
    allocations_.push_back(p);
    return p;
}

void Runtime::free(void *p) {
    if (!p) return;
    auto it = std::find(allocations_.begin(), allocations_.end(), p);
    if (it != allocations_.end()) {
        allocations_.erase(it);
        cudaStreamSynchronize(stream_);
        if (exported_.count(p)) {  // peers may still map it (CUDA IPC): keep it until the collective teardown in ~Runtime
            retired_.push_back(p);
            return;
        }
        cudaFree(p);
    }
}

double *Runtime::new_slot() {
    if (mode_ == Mode::Replay) return arena_chunks_.empty() ? nullptr : arena_chunks_.front();  // value unused on replay
    if (arena_chunks_.empty() || arena_used_ == kArenaChunk) {
        if (mode_ == Mode::Capture) fail(LSK_E_CAPACITY, "scalar arena growth inside a trace");
        double *chunk = nullptr;
        check_cuda(cudaMalloc(&chunk, sizeof(double) * kArenaChunk), "cudaMalloc(arena)");
        check_cuda(cudaMemset(chunk, 0, sizeof(double) * kArenaChunk), "cudaMemset(arena)");
        arena_chunks_.push_back(chunk);
        arena_used_ = 0;
    }
    return arena_chunks_.back() + arena_used_++;
}

// ---- tracing ---------------------------------------------------------------------------------------------
void Runtime::begin_trace(int id) {
    if (mode_ != Mode::Eager || eager_trace_) fail(LSK_E_INVALID, "begin_trace: traces do not nest");
    active_trace_ = id;
    static const bool eager = [] { const char *e = std::getenv("LSK_TRACE"); return e && std::string(e) == "eager"; }();
    if (eager) {  // developer switch: run traced regions launch by launch (no CUDA graph)
        mode_ = Mode::Eager;
        eager_trace_ = true;
        return;
    }
    if (traces_.count(id)) {
        mode_ = Mode::Replay;
        return;
    }
    check_cuda(cudaStreamBeginCapture(stream_, cudaStreamCaptureModeThreadLocal), "cudaStreamBeginCapture");
    capture_mark_ = lsk_ctx_launch_count(ctx_);
    mode_ = Mode::Capture;
}

void Runtime::end_trace(int id) {
    if (eager_trace_ && id == active_trace_) {
        eager_trace_ = false;
        active_trace_ = -1;
        return;
    }
    if (mode_ == Mode::Eager || id != active_trace_) fail(LSK_E_INVALID, "end_trace without matching begin_trace");
    if (mode_ == Mode::Capture) {
        cudaGraph_t graph = nullptr;
        mode_ = Mode::Eager;
        check_cuda(cudaStreamEndCapture(stream_, &graph), "cudaStreamEndCapture");
        Trace t;
        t.kernels = lsk_ctx_launch_count(ctx_) - capture_mark_;
        check_cuda(cudaGraphInstantiate(&t.exec, graph, 0), "cudaGraphInstantiate");
        cudaGraphDestroy(graph);
        traces_[id] = t;
    }
    mode_ = Mode::Eager;
    active_trace_ = -1;
    const Trace &t = traces_[id];
    check_cuda(cudaGraphLaunch(t.exec, stream_), "cudaGraphLaunch");
    // launches made while capturing were counted by the context but only run now; replays add theirs
    static_cast<void>(0);
    replayed_kernels_ += t.kernels;
}

void Runtime::fence() {
    check_cuda(cudaStreamSynchronize(stream_), "cudaStreamSynchronize");
    // A spin-wait of a peer-memory collective that gave up only raises a device
    // flag (and poisons the reduced value with NaN): surface it here, at the first host synchronisation after it, instead
    // of returning numbers computed from stale ghosts with status 0.
    if (nranks_ > 1 && p2p_ && mode_ == Mode::Eager) {
        int e = 0;
        if (lsk_comm_error(ctx_, stream_, &peers_, &e) == 0 && e != 0)
            fail(LSK_E_NCCL, "a peer-memory collective timed out (dead or diverged peer): results are invalid");
    }
}

uint64_t Runtime::kernel_launches() const {
    // the capture pass itself is counted once by lsk_ctx and once by the graph launch at end_trace:
    // subtract one copy per recorded trace so each executed kernel is counted exactly once
    uint64_t captured_once = 0;
    for (const auto &kv : traces_) captured_once += kv.second.kernels;
    return lsk_ctx_launch_count(ctx_) + replayed_kernels_ - captured_once;
}

}  // namespace LegionSolvers
