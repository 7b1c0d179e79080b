// PartitionedVector.hpp -- the reference's PartitionedVector<ENTRY_T> (src/PartitionedVector.hpp,
// src/PartitionedVector.cpp) over device memory: a 1-D index space [0, N), a disjoint + complete
// partition into `pieces` colours, and one field.  Same operations (constant_fill, operator= copy,
// scal, axpy with 1-3 scalars, xpay with 1-2, dot); each is one kernel launch per LOCAL piece --
// the analogue of one point task per piece -- on this rank's stream, with scalars read on the device.
//
// Placement follows the reference's BlockingShardingFunctor (src/LegionSolversMapper.cpp:140-151):
// colour c lives on rank c / ceil(pieces / nranks).  A rank stores the contiguous run of rows of its
// colours, plus the halo its matrices' ghost partitions ask for; element i is addressed through a
// pointer shifted to global index 0, the convention of the reference's mat-vec tasks
// (src/CuSPARSEHelpers.hpp:188-201).
#pragma once

#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "Runtime.hpp"
#include "Scalar.hpp"

namespace LegionSolvers {

// create_equal_partition of [0, volume) + the colours this rank owns
struct IndexPartition {
    int64_t volume = 0;
    int pieces = 0;
    std::vector<int64_t> lo, hi;  // inclusive bounds per colour
    int first_color = 0, end_color = 0;  // local colours [first, end)

    static std::shared_ptr<IndexPartition> equal(const Runtime &rt, int64_t volume, int pieces) {
        auto p = std::make_shared<IndexPartition>();
        p->volume = volume;
        p->pieces = pieces;
        p->lo.resize((size_t) pieces);
        p->hi.resize((size_t) pieces);
        if (lsk_equal_partition(volume, pieces, p->lo.data(), p->hi.data()) != 0)
            throw std::runtime_error("create_equal_partition: bad arguments");
        const int per = (pieces + rt.nranks() - 1) / rt.nranks();  // BlockingShardingFunctor
        p->first_color = std::min(pieces, rt.rank() * per);
        p->end_color = std::min(pieces, (rt.rank() + 1) * per);
        return p;
    }
    bool owns_any() const { return end_color > first_color; }
    int64_t own_lo() const { return owns_any() ? lo[(size_t) first_color] : 0; }
    int64_t own_hi() const { return owns_any() ? hi[(size_t) end_color - 1] : -1; }
    int64_t piece_size(int c) const { return hi[(size_t) c] - lo[(size_t) c] + 1; }
    bool same_as(const IndexPartition &o) const { return volume == o.volume && pieces == o.pieces && lo == o.lo && hi == o.hi; }
};

// [lo, hi] per colour (only the local colours are meaningful): kernel and ghost partitions
struct IntervalPartition {
    std::vector<int64_t> lo, hi;
    explicit IntervalPartition(int pieces = 0) : lo((size_t) pieces, 0), hi((size_t) pieces, -1) {}
};

template <typename T>
struct VectorKernels;
template <>
struct VectorKernels<double> {
    static int scal(lsk_ctx *c, cudaStream_t s, int64_t n, int nt, double *const *f, double *x) {
        return lsk_scal_f64(c, s, n, nt, f[0], f[1], f[2], f[3], x);
    }
    static int axpy(lsk_ctx *c, cudaStream_t s, int64_t n, int nt, double *const *f, const double *x, double *y) {
        return lsk_axpy_f64(c, s, n, nt, f[0], f[1], f[2], f[3], x, y);
    }
    static int xpay(lsk_ctx *c, cudaStream_t s, int64_t n, int nt, double *const *f, const double *x, double *y) {
        return lsk_xpay_f64(c, s, n, nt, f[0], f[1], f[2], f[3], x, y);
    }
    static int dot(lsk_ctx *c, cudaStream_t s, int64_t n, const double *v, const double *w, double *o) {
        return lsk_dot_f64(c, s, n, v, w, o);
    }
    static int fill(lsk_ctx *c, cudaStream_t s, int64_t n, double v, double *x) { return lsk_fill_f64(c, s, n, v, x); }
};
template <>
struct VectorKernels<float> {
    static int scal(lsk_ctx *c, cudaStream_t s, int64_t n, int nt, float *const *f, float *x) {
        return lsk_scal_f32(c, s, n, nt, f[0], f[1], f[2], f[3], x);
    }
    static int axpy(lsk_ctx *c, cudaStream_t s, int64_t n, int nt, float *const *f, const float *x, float *y) {
        return lsk_axpy_f32(c, s, n, nt, f[0], f[1], f[2], f[3], x, y);
    }
    static int xpay(lsk_ctx *c, cudaStream_t s, int64_t n, int nt, float *const *f, const float *x, float *y) {
        return lsk_xpay_f32(c, s, n, nt, f[0], f[1], f[2], f[3], x, y);
    }
    static int dot(lsk_ctx *c, cudaStream_t s, int64_t n, const float *v, const float *w, float *o) {
        return lsk_dot_f32(c, s, n, v, w, o);
    }
    static int fill(lsk_ctx *c, cudaStream_t s, int64_t n, float v, float *x) { return lsk_fill_f32(c, s, n, v, x); }
};

template <typename T>
class PartitionedVector {
    struct Storage {
        Runtime *rt;
        std::string name;
        std::shared_ptr<IndexPartition> part;
        int64_t buf_lo = 0, buf_hi = -1;  // global range held on this rank (owned rows + halo)
        void *raw = nullptr;              // allocation
        T *base = nullptr;                // element buf_lo
        ~Storage() {
            if (raw) rt->free(raw);
        }
    };
    std::shared_ptr<Storage> st;

    // allocate so that the first OWNED element sits on a 256-byte boundary (full-width vector path)
    static void allocate(Storage &s, int64_t lo, int64_t hi) {
        const int64_t own_lo = s.part->own_lo();
        const size_t n = hi >= lo ? (size_t) (hi - lo + 1) : 0;
        const size_t lead = (size_t) (own_lo > lo ? own_lo - lo : 0) * sizeof(T);
        s.raw = s.rt->alloc(n * sizeof(T) + 512);
        const uintptr_t want = (reinterpret_cast<uintptr_t>(s.raw) + lead + 255) & ~uintptr_t(255);
        s.base = reinterpret_cast<T *>(want - lead);
        s.buf_lo = lo;
        s.buf_hi = hi;
        s.rt->check_cuda(cudaMemsetAsync(s.base, 0, n * sizeof(T), s.rt->stream()), "vector init");
    }

public:
    PartitionedVector() = delete;
    explicit PartitionedVector(Runtime *rt, const std::string &name, std::shared_ptr<IndexPartition> part)
        : st(std::make_shared<Storage>()) {
        st->rt = rt;
        st->name = name;
        st->part = std::move(part);
        allocate(*st, st->part->own_lo(), st->part->own_hi());
    }
    PartitionedVector(const PartitionedVector &) = default;  // aliases the same region, like the reference

    Runtime *runtime() const { return st->rt; }
    const std::string &get_name() const { return st->name; }
    const IndexPartition &partition() const { return *st->part; }
    std::shared_ptr<IndexPartition> get_index_partition() const { return st->part; }
    int64_t volume() const { return st->part->volume; }

    T *ptr(int64_t global_index) const { return st->base + (global_index - st->buf_lo); }
    T *shifted() const {  // indexable by GLOBAL index; only dereferenced inside [buf_lo, buf_hi]
        return reinterpret_cast<T *>(reinterpret_cast<uintptr_t>(st->base) - (uintptr_t) st->buf_lo * sizeof(T));
    }
    int64_t buf_lo() const { return st->buf_lo; }
    int64_t buf_hi() const { return st->buf_hi; }

    // make room for a ghost interval [lo, hi]; owned data is preserved (set-up time only).  Ghost values are only ever
    // written by this rank's own kernels (the halo exchange unpacks the neighbours' packets locally), so a vector's
    // buffer is never mapped into another process and can be re-allocated freely.
    void ensure_range(int64_t lo, int64_t hi) {
        if (hi < lo) return;
        const int64_t nlo = std::min(lo, st->buf_lo), nhi = std::max(hi, st->buf_hi);
        if (st->buf_hi >= st->buf_lo && nlo == st->buf_lo && nhi == st->buf_hi) return;
        Runtime *rt = st->rt;
        if (rt->capturing() || rt->replaying()) rt->fail(LSK_E_INVALID, "ensure_range inside a trace");
        void *old_raw = st->raw;
        T *old_base = st->base;
        const int64_t old_lo = st->buf_lo, old_hi = st->buf_hi;
        st->raw = nullptr;
        allocate(*st, nlo, nhi);
        if (old_hi >= old_lo)
            rt->check_cuda(cudaMemcpyAsync(ptr(old_lo), old_base, sizeof(T) * (size_t) (old_hi - old_lo + 1),
                                           cudaMemcpyDeviceToDevice, rt->stream()), "vector regrow");
        if (old_raw) rt->free(old_raw);
    }

    template <class F>
    void for_local_pieces(F &&f) const {
        const IndexPartition &p = *st->part;
        for (int c = p.first_color; c < p.end_color; ++c) f(c, p.lo[(size_t) c], p.piece_size(c));
    }

    // ---- IndexFill (src/PartitionedVector.cpp:150-173) ----------------------------------------------
    void constant_fill(T value) {
        Runtime *rt = st->rt;
        for_local_pieces([&](int, int64_t lo, int64_t n) {
            T *x = ptr(lo);
            rt->enqueue("fill", [&] { return VectorKernels<T>::fill(rt->ctx(), rt->stream(), n, value, x); });
        });
    }
    void constant_fill(const Scalar<T> &value) {
        static_assert(std::is_same<T, double>::value, "device-scalar fill is instantiated for fp64");
        Runtime *rt = st->rt;
        for_local_pieces([&](int, int64_t lo, int64_t n) {
            T *x = ptr(lo);
            const T *v = value.ptr();
            rt->enqueue("fill", [&] { return lsk_fill_dev_f64(rt->ctx(), rt->stream(), n, v, x); });
        });
    }
    void zero_fill() { constant_fill(static_cast<T>(0)); }
    // the ghost elements this rank holds beside its owned rows := 0 (before a transposed mat-vec accumulates into them)
    void zero_ghosts() {
        Runtime *rt = st->rt;
        const int64_t own_lo = st->part->own_lo(), own_hi = st->part->own_hi();
        const int64_t lo_n = own_lo > st->buf_lo ? own_lo - st->buf_lo : 0;
        const int64_t hi_n = st->buf_hi > own_hi ? st->buf_hi - std::max(own_hi, st->buf_lo - 1) : 0;
        if (lo_n > 0) {
            T *x = ptr(st->buf_lo);
            rt->enqueue("fill", [&] { return VectorKernels<T>::fill(rt->ctx(), rt->stream(), lo_n, (T) 0, x); });
        }
        if (hi_n > 0) {
            T *x = ptr(st->buf_hi - hi_n + 1);
            rt->enqueue("fill", [&] { return VectorKernels<T>::fill(rt->ctx(), rt->stream(), hi_n, (T) 0, x); });
        }
    }
    T operator=(T value) {
        constant_fill(value);
        return value;
    }

    // ---- IndexCopy (src/PartitionedVector.cpp:176-192) ------------------------------------------------
    const PartitionedVector &operator=(const PartitionedVector &x) {
        if (st == x.st) return x;
        require_same(x);
        Runtime *rt = st->rt;
        for_local_pieces([&](int, int64_t lo, int64_t n) {
            T *d = ptr(lo);
            const T *s = x.ptr(lo);
            rt->enqueue("copy", [&] {
                if (n == 0) return 0;
                if constexpr (std::is_same<T, double>::value) return lsk_copy_f64(rt->ctx(), rt->stream(), n, s, d);  // a kernel: never queues behind host copies
                else return (int) cudaMemcpyAsync(d, s, sizeof(T) * (size_t) n, cudaMemcpyDeviceToDevice, rt->stream());
            });
        });
        return x;
    }

    // ---- BLAS-1 launchers (src/PartitionedVector.cpp:195-358) -------------------------------------------
    void scal(const Scalar<T> &alpha) { launch1(0, {alpha.ptr()}, nullptr); }
    void axpy(const Scalar<T> &alpha, const PartitionedVector &x) { launch1(1, {alpha.ptr()}, &x); }
    void axpy(T alpha, const PartitionedVector &x) { axpy(Scalar<T>(st->rt, alpha), x); }
    void axpy(const Scalar<T> &numer, const Scalar<T> &denom, const PartitionedVector &x) {
        launch1(1, {numer.ptr(), denom.ptr()}, &x);
    }
    void axpy(const Scalar<T> &n1, const Scalar<T> &n2, const Scalar<T> &denom, const PartitionedVector &x) {
        launch1(1, {n1.ptr(), n2.ptr(), denom.ptr()}, &x);
    }
    void xpay(const Scalar<T> &alpha, const PartitionedVector &x) { launch1(2, {alpha.ptr()}, &x); }
    void xpay(T alpha, const PartitionedVector &x) { xpay(Scalar<T>(st->rt, alpha), x); }
    void xpay(const Scalar<T> &numer, const Scalar<T> &denom, const PartitionedVector &x) {
        launch1(2, {numer.ptr(), denom.ptr()}, &x);
    }

    // DotTask per piece + sum of the per-piece futures (src/PartitionedVector.cpp:337-358): partials
    // are folded on the device in colour order, then summed across ranks.
    Scalar<T> dot(const PartitionedVector &x) const {
        Scalar<T> result(st->rt);
        dot_into(x, result);
        return result;
    }
    void dot_into(const PartitionedVector &x, const Scalar<T> &out) const {
        require_same(x);
        Runtime *rt = st->rt;
        const IndexPartition &p = *st->part;
        bool first = true;
        std::vector<Scalar<T>> parts;
        for (int c = p.first_color; c < p.end_color; ++c) {
            Scalar<T> part = first ? out : Scalar<T>(rt);
            const T *v = ptr(p.lo[(size_t) c]), *w = x.ptr(p.lo[(size_t) c]);
            T *o = part.ptr();
            const int64_t n = p.piece_size(c);
            rt->enqueue("dot", [&] { return VectorKernels<T>::dot(rt->ctx(), rt->stream(), n, v, w, o); });
            if (!first) parts.push_back(part);
            first = false;
        }
        if (first) {  // this rank owns no piece: contribute 0
            T *o = out.ptr();
            rt->enqueue("dot", [&] { return VectorKernels<T>::fill(rt->ctx(), rt->stream(), 1, (T) 0, o); });
        }
        for (const Scalar<T> &part : parts) {
            T *o = out.ptr();
            const T *b = part.ptr();
            rt->enqueue("dot fold", [&] { return ScalarKernels<T>::op(rt->ctx(), rt->stream(), LSK_OP_ADD, o, b, o); });
        }
        if (rt->fused_collectives()) {  // the dot kernel's tail already summed across ranks
            if (p.end_color - p.first_color != 1) rt->fail(LSK_E_INVALID, "fused reductions need exactly one local piece per rank");
        } else if (std::is_same<T, double>::value) rt->allreduce_sum(reinterpret_cast<double *>(out.ptr()), 1);
        else if (rt->nranks() > 1) rt->fail(LSK_E_INVALID, "multi-rank dot is instantiated for fp64");
    }

    // ---- host access to the OWNED rows (tests, benchmark I/O) -----------------------------------------------
    void copy_from_host(const T *global_array) {  // global_array indexed by global row
        Runtime *rt = st->rt;
        const IndexPartition &p = *st->part;
        if (!p.owns_any()) return;
        rt->check_cuda(cudaMemcpyAsync(ptr(p.own_lo()), global_array + p.own_lo(),
                                       sizeof(T) * (size_t) (p.own_hi() - p.own_lo() + 1), cudaMemcpyHostToDevice,
                                       rt->stream()), "vector H2D");
    }
    void copy_to_host(T *global_array) const {
        Runtime *rt = st->rt;
        const IndexPartition &p = *st->part;
        if (!p.owns_any()) return;
        rt->check_cuda(cudaMemcpyAsync(global_array + p.own_lo(), ptr(p.own_lo()),
                                       sizeof(T) * (size_t) (p.own_hi() - p.own_lo() + 1), cudaMemcpyDeviceToHost,
                                       rt->stream()), "vector D2H");
    }

    // Asynchronous forms on a stream of the caller's choice (null = the runtime's stream), for overlapping the I/O of
    // one solve with the iterations of another: `global_array` may be pinned host memory or device memory
    // (cudaMemcpyDefault).  On a foreign stream the caller orders the copy against the solver's work with events.
    void copy_from_async(const T *global_array, cudaStream_t s) {
        Runtime *rt = st->rt;
        const IndexPartition &p = *st->part;
        if (!p.owns_any()) return;
        rt->check_cuda(cudaMemcpyAsync(ptr(p.own_lo()), global_array + p.own_lo(), sizeof(T) * (size_t) (p.own_hi() - p.own_lo() + 1),
                                       cudaMemcpyDefault, s ? s : rt->stream()), "vector copy-in (async)");
    }
    void copy_to_async(T *global_array, cudaStream_t s) const {
        Runtime *rt = st->rt;
        const IndexPartition &p = *st->part;
        if (!p.owns_any()) return;
        rt->check_cuda(cudaMemcpyAsync(global_array + p.own_lo(), ptr(p.own_lo()), sizeof(T) * (size_t) (p.own_hi() - p.own_lo() + 1),
                                       cudaMemcpyDefault, s ? s : rt->stream()), "vector copy-out (async)");
    }

    void require_same(const PartitionedVector &x) const {
        if (st->part != x.st->part && !st->part->same_as(*x.st->part))
            st->rt->fail(LSK_E_INVALID, "vectors live on different partitions");
    }

private:
    void launch1(int which, std::initializer_list<T *> terms, const PartitionedVector *x) {
        if (x) require_same(*x);
        Runtime *rt = st->rt;
        T *f[4] = {nullptr, nullptr, nullptr, nullptr};
        int nt = 0;
        for (T *t : terms) f[nt++] = t;
        for_local_pieces([&](int, int64_t lo, int64_t n) {
            T *y = ptr(lo);
            const T *xs = x ? x->ptr(lo) : nullptr;
            rt->enqueue(which == 0 ? "scal" : which == 1 ? "axpy" : "xpay", [&] {
                if (which == 0) return VectorKernels<T>::scal(rt->ctx(), rt->stream(), n, nt, f, y);
                if (which == 1) return VectorKernels<T>::axpy(rt->ctx(), rt->stream(), n, nt, f, xs, y);
                return VectorKernels<T>::xpay(rt->ctx(), rt->stream(), n, nt, f, xs, y);
            });
        });
    }
};

}  // namespace LegionSolvers
