// StencilGenerator.hpp -- create_linearized_csr_stencil_matrix (src/StencilGenerator.hpp:533-643) on
// the GPU.  The reference equal-partitions the kernel and rowptr regions into `num_pieces` and has
// every piece walk the whole grid sequentially on a CPU (FillLinearizedCSRStencilTask,
// src/StencilGenerator.cpp:380-543).  Here each rank fills exactly the slab its colours need -- rows
// [own_lo, own_hi] of the equal row partition and their non-zeros -- with one thread per row:
// count, exclusive scan, fill.  Same sorted-offset order, same inclusive global-k rects: bit-identical.
#pragma once

#include <memory>

#include "Matrices.hpp"

namespace LegionSolvers {

enum class IndexOrder { ROW_MAJOR = 0, COLUMN_MAJOR = 1 };

inline std::unique_ptr<CSRMatrix<double>> create_linearized_csr_stencil_matrix(Runtime *rt, lsk_stencil stencil,
                                                                              std::size_t num_pieces) {
    if (lsk_stencil_sort(&stencil) != 0) rt->fail(LSK_E_INVALID, "stencil description");
    int64_t n = 1;
    for (int d = 0; d < stencil.dim; ++d) n *= stencil.shape[d];
    const int64_t nnz_global = lsk_stencil_size(&stencil);
    auto rows = IndexPartition::equal(*rt, n, (int) num_pieces);
    const int64_t r_lo = rows->own_lo(), r_hi = rows->own_hi();
    // global k of the slab's first non-zero = non-zeros of all earlier rows; slab size likewise
    DeviceBuffer<int64_t> counts(rt, 2);
    rt->enqueue("stencil count", [&] { return lsk_stencil_count_f64(rt->ctx(), rt->stream(), &stencil, 0, r_lo - 1, counts.ptr); });
    rt->enqueue("stencil count", [&] { return lsk_stencil_count_f64(rt->ctx(), rt->stream(), &stencil, r_lo, r_hi, counts.ptr + 1); });
    int64_t h[2];
    rt->check_cuda(cudaMemcpyAsync(h, counts.ptr, sizeof(h), cudaMemcpyDeviceToHost, rt->stream()), "stencil counts D2H");
    rt->fence();
    const int64_t k_lo = h[0], k_hi = h[0] + h[1] - 1;
    auto m = std::make_unique<CSRMatrix<double>>(rt, n, n, nnz_global, r_lo, r_hi, k_lo, k_hi);
    DeviceBuffer<int64_t> scratch(rt, (size_t) std::max<int64_t>(1, r_hi - r_lo + 2));
    rt->enqueue("stencil fill", [&] {
        return lsk_stencil_fill_csr_f64(rt->ctx(), rt->stream(), &stencil, r_lo, r_hi, k_lo, m->entry_ptr(), m->col_ptr(),
                                        m->rowptr_ptr(), scratch.ptr);
    });
    rt->fence();
    return m;
}

// The stencils of test/BenchmarkStencil.cpp:33-131 (-dim 1, 2, 3; 4 = 3-D 27-point).
inline lsk_stencil benchmark_stencil(int dim_flag, int64_t nx, int64_t ny, int64_t nz) {
    lsk_stencil st{};
    st.order = 0;
    auto add = [&](int64_t a, int64_t b, int64_t c, double v) {
        st.offsets[st.noff][0] = a;
        st.offsets[st.noff][1] = b;
        st.offsets[st.noff][2] = c;
        st.values[st.noff] = v;
        ++st.noff;
    };
    switch (dim_flag) {
    case 1:
        st.dim = 1; st.shape[0] = nx;
        add(0, 0, 0, 2.0); add(-1, 0, 0, -1.0); add(1, 0, 0, -1.0);
        break;
    case 2:
        st.dim = 2; st.shape[0] = nx; st.shape[1] = ny;
        add(0, 0, 0, 4.0); add(-1, 0, 0, -1.0); add(1, 0, 0, -1.0); add(0, -1, 0, -1.0); add(0, 1, 0, -1.0);
        break;
    case 3:
        st.dim = 3; st.shape[0] = nx; st.shape[1] = ny; st.shape[2] = nz;
        add(0, 0, 0, 6.0);
        add(-1, 0, 0, -1.0); add(1, 0, 0, -1.0); add(0, -1, 0, -1.0); add(0, 1, 0, -1.0); add(0, 0, -1, -1.0); add(0, 0, 1, -1.0);
        break;
    case 4:
        st.dim = 3; st.shape[0] = nx; st.shape[1] = ny; st.shape[2] = nz;
        for (int a = -1; a <= 1; ++a)
            for (int b = -1; b <= 1; ++b)
                for (int c = -1; c <= 1; ++c) {
                    const int nzc = (a != 0) + (b != 0) + (c != 0);
                    const double v = nzc == 0 ? 88.0 / 26.0 : nzc == 1 ? -6.0 / 26.0 : nzc == 2 ? -3.0 / 26.0 : -2.0 / 26.0;
                    add(a, b, c, v);
                }
        break;
    default: throw std::runtime_error("INVALID DIM");
    }
    return st;
}

}  // namespace LegionSolvers
