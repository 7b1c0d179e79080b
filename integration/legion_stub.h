// legion_stub.h -- the handful of Legion / LegionSolvers declarations that the patched `cuda_task_body` bodies in
// cuda_task_bodies.cpp touch, so that the binding is type-checked in an image without Legion (tests/test_abi.py
// compiles it with -fsyntax-only).  Shapes follow legion.h and the reference's src/LegionUtilities.hpp; nothing here
// executes.  With a real Legion install, compile cuda_task_bodies.cpp with -DLSK_WITH_LEGION instead.
#pragma once
#include <cstddef>
#include <set>
#include <vector>

typedef struct CUstream_st *cudaStream_t;

namespace Legion {
typedef long long coord_t;
typedef unsigned int FieldID;
typedef unsigned int ReductionOpID;
struct Context {};
class Runtime;

template <int DIM, typename T = coord_t>
struct Point {
    T x[DIM];
    T &operator[](int i) { return x[i]; }
    const T &operator[](int i) const { return x[i]; }
};
template <int DIM, typename T = coord_t>
struct Rect {
    Point<DIM, T> lo, hi;  // INCLUSIVE bounds
    std::size_t volume() const {
        std::size_t v = 1;
        for (int d = 0; d < DIM; ++d) v *= hi[d] >= lo[d] ? (std::size_t) (hi[d] - lo[d] + 1) : 0;
        return v;
    }
};
template <int DIM, typename T = coord_t>
struct DomainT {
    Rect<DIM, T> bounds;
    bool dense() const { return true; }
    bool empty() const { return bounds.volume() == 0; }
};
struct Future {
    template <typename T>
    T get_result() const { return T(); }
};
struct RegionRequirement {
    std::set<FieldID> privilege_fields;
};
struct Task {
    std::vector<Future> futures;
    std::vector<RegionRequirement> regions;
    void *args;
    std::size_t arglen;
};
struct PhysicalRegion {
    template <int DIM, typename T>
    DomainT<DIM, T> get_bounds() const { return DomainT<DIM, T>(); }
};
}  // namespace Legion

namespace LegionSolvers {
// src/LegionUtilities.hpp: FieldAccessor aliases with an affine layout; only ptr(point) is used by the GPU bodies
template <typename FT, int DIM, typename COORD_T>
struct AffineReader {
    AffineReader(const Legion::PhysicalRegion &, Legion::FieldID) {}
    const FT *ptr(const Legion::Point<DIM, COORD_T> &) const { return nullptr; }
};
template <typename FT, int DIM, typename COORD_T>
struct AffineReaderWriter {
    AffineReaderWriter(const Legion::PhysicalRegion &, Legion::FieldID) {}
    FT *ptr(const Legion::Point<DIM, COORD_T> &) const { return nullptr; }
};
template <typename FT, int DIM, typename COORD_T>
struct AffineSumAccessor {
    AffineSumAccessor(const Legion::PhysicalRegion &, Legion::FieldID, Legion::ReductionOpID) {}
    FT *ptr(const Legion::Point<DIM, COORD_T> &) const { return nullptr; }
};
template <typename T>
constexpr Legion::ReductionOpID LEGION_REDOP_SUM = 0;
// src/LegionUtilities.cpp:72-97
template <typename T>
T get_alpha(const std::vector<Legion::Future> &) { return T(1); }
// src/CUDAUtilities.cpp:66-75: Realm's task stream
inline cudaStream_t get_cuda_stream() { return nullptr; }
}  // namespace LegionSolvers

extern "C" int cudaMemcpyAsync(void *, const void *, std::size_t, int, cudaStream_t);
extern "C" int cudaStreamSynchronize(cudaStream_t);
#define LSK_STUB_MEMCPY_H2D 1
#define LSK_STUB_MEMCPY_D2H 2
