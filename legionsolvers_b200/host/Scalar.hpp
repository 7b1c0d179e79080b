// Scalar.hpp -- device-resident Scalar<T>, the stand-in for the reference's future-backed Scalar
// (src/Scalar.hpp, src/Scalar.cpp).  Same operator set; each operation is one single-thread kernel
// writing a fresh arena slot instead of one CPU task launch, and nothing ever waits on the host
// except get_value() -- the analogue of Future::get_result.
#pragma once

#include <type_traits>
#include <vector>

#include "Runtime.hpp"

namespace LegionSolvers {

template <typename T>
struct ScalarKernels;
template <>
struct ScalarKernels<double> {
    static int op(lsk_ctx *c, cudaStream_t s, int o, const double *a, const double *b, double *out) {
        return lsk_scalar_op_f64(c, s, o, a, b, out);
    }
    static int fill(lsk_ctx *c, cudaStream_t s, int64_t n, double v, double *x) { return lsk_fill_f64(c, s, n, v, x); }
};
template <>
struct ScalarKernels<float> {
    static int op(lsk_ctx *c, cudaStream_t s, int o, const float *a, const float *b, float *out) {
        return lsk_scalar_op_f32(c, s, o, a, b, out);
    }
    static int fill(lsk_ctx *c, cudaStream_t s, int64_t n, float v, float *x) { return lsk_fill_f32(c, s, n, v, x); }
};

template <typename T>
class Scalar {
    static_assert(std::is_same<T, double>::value || std::is_same<T, float>::value, "entry type");
    Runtime *rt;
    T *slot;  // device

    Scalar unary(int op) const {
        Scalar r(rt);
        const T *a = slot;
        T *o = r.slot;
        rt->enqueue("scalar op", [&] { return ScalarKernels<T>::op(rt->ctx(), rt->stream(), op, a, nullptr, o); });
        return r;
    }
    Scalar binary(int op, const Scalar &rhs) const {
        Scalar r(rt);
        const T *a = slot, *b = rhs.slot;
        T *o = r.slot;
        rt->enqueue("scalar op", [&] { return ScalarKernels<T>::op(rt->ctx(), rt->stream(), op, a, b, o); });
        return r;
    }

public:
    Scalar() = delete;
    // fresh slot, value undefined until written
    explicit Scalar(Runtime *rt_) : rt(rt_), slot(reinterpret_cast<T *>(rt_->new_slot())) {}
    // Future::from_value (src/Scalar.hpp:30-33)
    explicit Scalar(Runtime *rt_, const T &value) : Scalar(rt_) {
        T *o = slot;
        rt->enqueue("scalar from value", [&] { return ScalarKernels<T>::fill(rt->ctx(), rt->stream(), 1, value, o); });
    }
    // view of an existing device slot
    explicit Scalar(Runtime *rt_, T *existing) : rt(rt_), slot(existing) {}
    Scalar(const Scalar &) = default;
    Scalar &operator=(const Scalar &rhs) {
        slot = rhs.slot;  // like the reference: rebinds the handle (no need to overwrite rt)
        return *this;
    }

    Runtime *runtime() const { return rt; }
    T *ptr() const { return slot; }

    // the only host synchronisation point (Future::get_result)
    T get_value() const {
        T v;
        rt->check_cuda(cudaMemcpyAsync(&v, slot, sizeof(T), cudaMemcpyDeviceToHost, rt->stream()), "scalar D2H");
        rt->fence();
        return v;
    }
    // copy the VALUE into this handle's slot (device-side), keeping the slot
    void assign_value(const Scalar &src) {
        const T *a = src.slot;
        T *o = slot;
        rt->enqueue("scalar copy", [&] { return ScalarKernels<T>::op(rt->ctx(), rt->stream(), LSK_OP_COPY, a, nullptr, o); });
    }

    // write op(a[, b]) into THIS handle's existing slot (no allocation: usable inside a trace)
    void set(int op, const Scalar &a) {
        const T *pa = a.slot;
        T *o = slot;
        rt->enqueue("scalar op", [&] { return ScalarKernels<T>::op(rt->ctx(), rt->stream(), op, pa, nullptr, o); });
    }
    void set(int op, const Scalar &a, const Scalar &b) {
        const T *pa = a.slot, *pb = b.slot;
        T *o = slot;
        rt->enqueue("scalar op", [&] { return ScalarKernels<T>::op(rt->ctx(), rt->stream(), op, pa, pb, o); });
    }

    Scalar operator+() const { return *this; }
    Scalar operator-() const { return unary(LSK_OP_NEG); }
    Scalar operator+(const Scalar &rhs) const { return binary(LSK_OP_ADD, rhs); }
    Scalar operator-(const Scalar &rhs) const { return binary(LSK_OP_SUB, rhs); }
    Scalar operator*(const Scalar &rhs) const { return binary(LSK_OP_MUL, rhs); }
    Scalar operator/(const Scalar &rhs) const { return binary(LSK_OP_DIV, rhs); }
    Scalar sqrt() const { return unary(LSK_OP_SQRT); }
    Scalar rsqrt() const { return unary(LSK_OP_RSQRT); }
    static Scalar dummy(Runtime *rt_) {  // DummyTask (src/UtilityTasks.cpp:96-99): returns 1
        Scalar r(rt_);
        T *o = r.slot;
        rt_->enqueue("dummy", [&] { return ScalarKernels<T>::op(rt_->ctx(), rt_->stream(), LSK_OP_DUMMY, nullptr, nullptr, o); });
        return r;
    }
};

// Growing list of scalars whose LENGTH lives on the device, so that a recorded trace can append to
// it on every replay (CGSolver::residual_norm_squared, BiCGStabSolver::rho/alpha/omega).  Circular
// with a fixed capacity -- the change the reference's TODO asks for (src/CGSolver.hpp:25-26).
class ScalarHistory {
    Runtime *rt;
    DeviceBuffer<double> hist;
    DeviceBuffer<int64_t> count;
    int64_t capacity;

public:
    ScalarHistory(Runtime *rt_, int64_t capacity_ = 1 << 16)
        : rt(rt_), hist(rt_, (size_t) capacity_), count(rt_, 1), capacity(capacity_) {
        rt->check_cuda(cudaMemsetAsync(count.ptr, 0, sizeof(int64_t), rt->stream()), "history init");
    }
    // push_back(value); optionally also copy the value into `also` (the "current" slot of a solver)
    void push_back(const Scalar<double> &value, const Scalar<double> *also = nullptr) {
        const double *v = value.ptr();
        double *a = also ? also->ptr() : nullptr;
        rt->enqueue("history append", [&] {
            return lsk_scalar_append_f64(rt->ctx(), rt->stream(), v, hist.ptr, capacity, count.ptr, a);
        });
    }
    // raw views for kernels that append by themselves (lsk_cg_direction_f64)
    double *data() const { return hist.ptr; }
    int64_t *count_ptr() const { return count.ptr; }
    int64_t get_capacity() const { return capacity; }
    void clear() {
        int64_t *c = count.ptr;
        rt->enqueue("history clear", [&] { return (int) cudaMemsetAsync(c, 0, sizeof(int64_t), rt->stream()); });
    }
    int64_t size() const {
        int64_t n = 0;
        rt->check_cuda(cudaMemcpyAsync(&n, count.ptr, sizeof(n), cudaMemcpyDeviceToHost, rt->stream()), "history size");
        rt->fence();
        return n;
    }
    // the most recent min(size, capacity) values, oldest first (synchronises)
    std::vector<double> to_host() const {
        const int64_t n = size();
        const int64_t keep = n < capacity ? n : capacity;
        std::vector<double> ring((size_t) capacity), out((size_t) keep);
        rt->check_cuda(cudaMemcpyAsync(ring.data(), hist.ptr, sizeof(double) * (size_t) capacity, cudaMemcpyDeviceToHost,
                                       rt->stream()), "history D2H");
        rt->fence();
        for (int64_t i = 0; i < keep; ++i) out[(size_t) i] = ring[(size_t) ((n - keep + i) % capacity)];
        return out;
    }
};

}  // namespace LegionSolvers
