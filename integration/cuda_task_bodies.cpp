// cuda_task_bodies.cpp -- the GPU variants of the reference's six hot-path leaf tasks, re-written on liblsk.so.
//
// Each function below is the body a maintainer puts in place of the cuBLAS / cuSPARSE call sequence of the
// corresponding `cuda_task_body` (same signature, same region / future / argument contract, same TaskID -- see
// lsk_task_ids.h -- and the same registration through TaskTDI / TaskTDDDIII::preregister, which picks the function up
// by name).  Compiled against integration/legion_stub.h here (syntax and types only; tests/test_abi.py), against
// <legion.h> with -DLSK_WITH_LEGION in a real build.
//
//   task (reference file:line of the body replaced)                      TaskID fp64 / 1-D / long long
//   ScalTask::cuda_task_body      src/LinearAlgebraTasks.cu:14-56          557044
//   AxpyTask::cuda_task_body      src/LinearAlgebraTasks.cu:59-113         562228
//   XpayTask::cuda_task_body      src/LinearAlgebraTasks.cu:118-176        567412
//   DotTask::cuda_task_body       src/LinearAlgebraTasks.cu:179-238        572596
//   COOMatvecTask::cuda_task_body src/COOMatrixTasks.cu:12-146             577888
//   CSRMatvecTask::cuda_task_body src/CSRMatrixTasks.cu:14-156             593440
//   (COORmatvecTask 583072 / CSRRmatvecTask 598624: `assert(false)` in the reference; bodies given at the end)
#ifdef LSK_WITH_LEGION
#include <legion.h>

#include "LegionUtilities.hpp"
#include "CudaLibs.hpp"
#else
#include "legion_stub.h"
#endif

#include <cassert>
#include <cstdint>
#include <cstdio>
#include <vector>

#include "lsk.h"
#include "lsk_task_ids.h"

namespace LegionSolvers {

// Rect<1, long long> and lsk_rect are the same 16 bytes: the rowptr field is passed through unconverted
static_assert(sizeof(Legion::Rect<1, long long>) == sizeof(lsk_rect), "rowptr field layout");
static_assert(sizeof(Legion::Point<1, long long>) == sizeof(int64_t), "col / row field layout");

#define CHECK_LSK(expr)                                                                   \
    do {                                                                                  \
        const int s_ = (expr);                                                            \
        if (s_ != 0) {                                                                    \
            std::fprintf(stderr, "%s:%d: %s\n", __FILE__, __LINE__, lsk_error_string(s_)); \
            assert(false);                                                                \
        }                                                                                 \
    } while (0)

// per-GPU-processor context and a small ring of device scalar slots, next to the cuBLAS / cuSPARSE handles of
// src/CUDAUtilities.cpp:125-145 (indexing by proc.id as there; one processor shown)
inline lsk_ctx *get_lsk_ctx() {
    static lsk_ctx *ctx = nullptr;
    if (ctx == nullptr) CHECK_LSK(lsk_ctx_create(0, &ctx));
    return ctx;
}
double *next_scalar_slot();  // 8-byte device slot from a per-processor ring (cudaMalloc'ed once)

#define LSK_TASK_ARGS                                                                                  \
    const Legion::Task *task, const std::vector<Legion::PhysicalRegion> &regions, Legion::Context ctx, \
        Legion::Runtime *rt

// the folded alpha of `task->futures` parked in a device slot (binding (a) of INTEGRATION.md section 2)
inline const double *alpha_slot(const Legion::Task *task, cudaStream_t stream) {
    const double alpha = get_alpha<double>(task->futures);
    double *slot = next_scalar_slot();
#ifdef LSK_WITH_LEGION
    CHECK_CUDA(cudaMemcpyAsync(slot, &alpha, sizeof(double), cudaMemcpyHostToDevice, stream));
#else
    (void) cudaMemcpyAsync(slot, &alpha, sizeof(double), LSK_STUB_MEMCPY_H2D, stream);
#endif
    return slot;
}

template <int DIM, typename COORD_T>
struct ScalTaskF64 {  // ScalTask<double, DIM, COORD_T>
    static void cuda_task_body(LSK_TASK_ARGS) {
        (void) ctx; (void) rt;
        assert(regions.size() == 1 && task->regions.size() == 1 && task->regions[0].privilege_fields.size() == 1);
        const Legion::FieldID x_fid = *task->regions[0].privilege_fields.begin();
        AffineReaderWriter<double, DIM, COORD_T> x(regions[0], x_fid);
        const auto x_domain = regions[0].template get_bounds<DIM, COORD_T>();
        assert(x_domain.dense());
        if (x_domain.empty()) return;
        auto stream = get_cuda_stream();
        CHECK_LSK(lsk_scal_f64(get_lsk_ctx(), stream, (int64_t) x_domain.bounds.volume(), 1, alpha_slot(task, stream), nullptr, nullptr, nullptr,
                               x.ptr(x_domain.bounds.lo)));
    }
};

template <int DIM, typename COORD_T>
struct AxpyTaskF64 {  // AxpyTask<double, DIM, COORD_T>: y = fma(alpha, x, y)
    static void cuda_task_body(LSK_TASK_ARGS) {
        (void) ctx; (void) rt;
        assert(regions.size() == 2 && task->regions.size() == 2);
        const Legion::FieldID y_fid = *task->regions[0].privilege_fields.begin();
        const Legion::FieldID x_fid = *task->regions[1].privilege_fields.begin();
        AffineReaderWriter<double, DIM, COORD_T> y(regions[0], y_fid);
        AffineReader<double, DIM, COORD_T> x(regions[1], x_fid);
        const auto y_domain = regions[0].template get_bounds<DIM, COORD_T>();
        const auto x_domain = regions[1].template get_bounds<DIM, COORD_T>();
        assert(y_domain.dense());
        if (y_domain.empty()) return;
        auto stream = get_cuda_stream();
        CHECK_LSK(lsk_axpy_f64(get_lsk_ctx(), stream, (int64_t) y_domain.bounds.volume(), 1, alpha_slot(task, stream), nullptr, nullptr, nullptr,
                               x.ptr(x_domain.bounds.lo), y.ptr(y_domain.bounds.lo)));
    }
};

template <int DIM, typename COORD_T>
struct XpayTaskF64 {  // XpayTask<double, DIM, COORD_T>: y = fma(alpha, y, x); replaces xpay_kernel + Pitches (a dense rect is one run)
    static void cuda_task_body(LSK_TASK_ARGS) {
        (void) ctx; (void) rt;
        assert(regions.size() == 2 && task->regions.size() == 2);
        const Legion::FieldID y_fid = *task->regions[0].privilege_fields.begin();
        const Legion::FieldID x_fid = *task->regions[1].privilege_fields.begin();
        AffineReaderWriter<double, DIM, COORD_T> y(regions[0], y_fid);
        AffineReader<double, DIM, COORD_T> x(regions[1], x_fid);
        const auto y_domain = regions[0].template get_bounds<DIM, COORD_T>();
        const auto x_domain = regions[1].template get_bounds<DIM, COORD_T>();
        assert(y_domain.dense());
        if (y_domain.empty()) return;
        auto stream = get_cuda_stream();
        CHECK_LSK(lsk_xpay_f64(get_lsk_ctx(), stream, (int64_t) y_domain.bounds.volume(), 1, alpha_slot(task, stream), nullptr, nullptr, nullptr,
                               x.ptr(x_domain.bounds.lo), y.ptr(y_domain.bounds.lo)));
    }
};

template <int DIM, typename COORD_T>
struct DotTaskF64 {  // DotTask<double, DIM, COORD_T>: returns the value, so the task synchronises as the reference does (:233-237)
    static double cuda_task_body(LSK_TASK_ARGS) {
        (void) ctx; (void) rt;
        assert(regions.size() == 2 && task->regions.size() == 2);
        const Legion::FieldID v_fid = *task->regions[0].privilege_fields.begin();
        const Legion::FieldID w_fid = *task->regions[1].privilege_fields.begin();
        AffineReader<double, DIM, COORD_T> v(regions[0], v_fid), w(regions[1], w_fid);
        const auto v_domain = regions[0].template get_bounds<DIM, COORD_T>();
        const auto w_domain = regions[1].template get_bounds<DIM, COORD_T>();
        assert(v_domain.dense());
        double result = 0.0;
        if (v_domain.empty()) return result;
        auto stream = get_cuda_stream();
        double *slot = next_scalar_slot();
        CHECK_LSK(lsk_dot_f64(get_lsk_ctx(), stream, (int64_t) v_domain.bounds.volume(), v.ptr(v_domain.bounds.lo), w.ptr(w_domain.bounds.lo), slot));
#ifdef LSK_WITH_LEGION
        CHECK_CUDA(cudaMemcpyAsync(&result, slot, sizeof(double), cudaMemcpyDeviceToHost, stream));
        CHECK_CUDA(cudaStreamSynchronize(stream));
#else
        (void) cudaMemcpyAsync(&result, slot, sizeof(double), LSK_STUB_MEMCPY_D2H, stream);
        (void) cudaStreamSynchronize(stream);
#endif
        return result;
    }
};

struct MatvecArgs {  // CSRMatvecTask::Args / COOMatvecTask::Args (src/CSRMatrixTasks.hpp:19-22, src/COOMatrixTasks.hpp:19-23)
    Legion::FieldID fid_entry, fid_row_or_unused, fid_col;
};

// CSRMatvecTask<double, 1, 1, 1, long long, long long, long long>: regions = {dst RW, kernel RO (entry, col), rowptr RO, ghost src RO}
struct CSRMatvecTaskF64 {
    static void cuda_task_body(LSK_TASK_ARGS) {
        (void) ctx; (void) rt;
        assert(regions.size() == 4 && task->regions.size() == 4);
        const Legion::FieldID out_fid = *task->regions[0].privilege_fields.begin();
        const Legion::FieldID rowptr_fid = *task->regions[2].privilege_fields.begin();
        const Legion::FieldID in_fid = *task->regions[3].privilege_fields.begin();
        const MatvecArgs args = *reinterpret_cast<const MatvecArgs *>(task->args);
        const auto output_domain = regions[0].get_bounds<1, long long>();
        const auto kernel_domain = regions[1].get_bounds<1, long long>();
        const auto rowptr_domain = regions[2].get_bounds<1, long long>();
        const auto input_domain = regions[3].get_bounds<1, long long>();
        const AffineSumAccessor<double, 1, long long> output_writer(regions[0], out_fid, LEGION_REDOP_SUM<double>);
        const AffineReader<double, 1, long long> entry_reader(regions[1], args.fid_entry);
        const AffineReader<Legion::Point<1, long long>, 1, long long> col_reader(regions[1], args.fid_col);
        const AffineReader<Legion::Rect<1, long long>, 1, long long> rowptr_reader(regions[2], rowptr_fid);
        const AffineReader<double, 1, long long> input_reader(regions[3], in_fid);
        assert(output_domain.dense());
        const auto rows = output_domain.bounds.volume();
        if (rows == 0) return;
        // everything from `get_cusparse_handle()` to the three cusparseDestroy* calls (src/CSRMatrixTasks.cu:85-155) becomes:
        CHECK_LSK(lsk_csr_spmv_f64(
            get_lsk_ctx(), get_cuda_stream(), (int64_t) rows, (int64_t) kernel_domain.bounds.volume(),
            entry_reader.ptr(kernel_domain.bounds.lo),
            reinterpret_cast<const int64_t *>(col_reader.ptr(kernel_domain.bounds.lo)),
            reinterpret_cast<const lsk_rect *>(rowptr_reader.ptr(rowptr_domain.bounds.lo)),  // inclusive rects of GLOBAL k, no indptr conversion
            (int64_t) kernel_domain.bounds.lo[0],                                             // k_base
            input_reader.ptr(input_domain.bounds.lo) - input_domain.bounds.lo[0],             // shifted x, as makeShiftedCuSparseDnVec
            output_writer.ptr(output_domain.bounds.lo),                                       // beta = 0, as the reference calls cuSPARSE
            nullptr, nullptr, nullptr, LSK_SPMV_AUTO));
    }
};

// COOMatvecTask<double, 1, 1, 1, long long x 3>: regions = {dst RW, kernel RO (entry, row, col), ghost src RO}
struct COOMatvecTaskF64 {
    static void cuda_task_body(LSK_TASK_ARGS) {
        (void) ctx; (void) rt;
        assert(regions.size() == 3 && task->regions.size() == 3);
        const Legion::FieldID out_fid = *task->regions[0].privilege_fields.begin();
        const Legion::FieldID in_fid = *task->regions[2].privilege_fields.begin();
        const MatvecArgs args = *reinterpret_cast<const MatvecArgs *>(task->args);
        const auto output_domain = regions[0].get_bounds<1, long long>();
        const auto coo_domain = regions[1].get_bounds<1, long long>();
        const auto input_domain = regions[2].get_bounds<1, long long>();
        const AffineSumAccessor<double, 1, long long> output_writer(regions[0], out_fid, LEGION_REDOP_SUM<double>);
        const AffineReader<double, 1, long long> entry_reader(regions[1], args.fid_entry);
        const AffineReader<Legion::Point<1, long long>, 1, long long> row_reader(regions[1], args.fid_row_or_unused);
        const AffineReader<Legion::Point<1, long long>, 1, long long> col_reader(regions[1], args.fid_col);
        const AffineReader<double, 1, long long> input_reader(regions[2], in_fid);
        if (coo_domain.empty()) return;
        CHECK_LSK(lsk_coo_spmv_f64(
            get_lsk_ctx(), get_cuda_stream(), (int64_t) coo_domain.bounds.volume(), entry_reader.ptr(coo_domain.bounds.lo),
            reinterpret_cast<const int64_t *>(row_reader.ptr(coo_domain.bounds.lo)),
            reinterpret_cast<const int64_t *>(col_reader.ptr(coo_domain.bounds.lo)),
            input_reader.ptr(input_domain.bounds.lo) - input_domain.bounds.lo[0],    // both vectors shifted to index 0 (src/COOMatrixTasks.cu:78-99)
            output_writer.ptr(output_domain.bounds.lo) - output_domain.bounds.lo[0],
            output_domain.bounds.lo[0], output_domain.bounds.hi[0], input_domain.bounds.lo[0], input_domain.bounds.hi[0]));  // beta = 1, as today
    }
};

// The two reserved transposed tasks: same regions as their forward twins with the roles of the vectors exchanged
// (dst lives on the DOMAIN space and is accumulated into; src is the piece of the range space).
struct CSRRmatvecTaskF64 {
    static void cuda_task_body(LSK_TASK_ARGS) {
        (void) ctx; (void) rt;
        assert(regions.size() == 4 && task->regions.size() == 4);
        const MatvecArgs args = *reinterpret_cast<const MatvecArgs *>(task->args);
        const auto output_domain = regions[0].get_bounds<1, long long>();  // ghost piece of the domain space
        const auto kernel_domain = regions[1].get_bounds<1, long long>();
        const auto rowptr_domain = regions[2].get_bounds<1, long long>();
        const auto input_domain = regions[3].get_bounds<1, long long>();   // range piece
        const AffineSumAccessor<double, 1, long long> output_writer(regions[0], *task->regions[0].privilege_fields.begin(), LEGION_REDOP_SUM<double>);
        const AffineReader<double, 1, long long> entry_reader(regions[1], args.fid_entry);
        const AffineReader<Legion::Point<1, long long>, 1, long long> col_reader(regions[1], args.fid_col);
        const AffineReader<Legion::Rect<1, long long>, 1, long long> rowptr_reader(regions[2], *task->regions[2].privilege_fields.begin());
        const AffineReader<double, 1, long long> input_reader(regions[3], *task->regions[3].privilege_fields.begin());
        if (input_domain.empty()) return;
        CHECK_LSK(lsk_csr_rspmv_f64(get_lsk_ctx(), get_cuda_stream(), (int64_t) input_domain.bounds.volume(), (int64_t) kernel_domain.bounds.volume(),
                                    entry_reader.ptr(kernel_domain.bounds.lo), reinterpret_cast<const int64_t *>(col_reader.ptr(kernel_domain.bounds.lo)),
                                    reinterpret_cast<const lsk_rect *>(rowptr_reader.ptr(rowptr_domain.bounds.lo)), (int64_t) kernel_domain.bounds.lo[0],
                                    input_reader.ptr(input_domain.bounds.lo), output_writer.ptr(output_domain.bounds.lo) - output_domain.bounds.lo[0],
                                    output_domain.bounds.lo[0], output_domain.bounds.hi[0]));
    }
};

// the ids these bodies are registered under (TaskTDI / TaskTDDDIII::preregister, Processor::TOC_PROC)
static_assert(LSK_TASK_BLOCK_SIZE == 5184, "block size");
inline int registered_ids(int (&out)[7]) {
    out[0] = LSK_TID_F64_1D_S64(LSK_BLOCK_SCAL);
    out[1] = LSK_TID_F64_1D_S64(LSK_BLOCK_AXPY);
    out[2] = LSK_TID_F64_1D_S64(LSK_BLOCK_XPAY);
    out[3] = LSK_TID_F64_1D_S64(LSK_BLOCK_DOT);
    out[4] = LSK_TID_MATVEC_F64_1D_S64(LSK_BLOCK_COO_MATVEC);
    out[5] = LSK_TID_MATVEC_F64_1D_S64(LSK_BLOCK_CSR_MATVEC);
    out[6] = LSK_TID_MATVEC_F64_1D_S64(LSK_BLOCK_CSR_RMATVEC);
    return 7;
}

// force the templates through the type checker for the configuration every BASELINE config uses
template struct ScalTaskF64<1, long long>;
template struct AxpyTaskF64<1, long long>;
template struct XpayTaskF64<1, long long>;
template struct DotTaskF64<1, long long>;
template struct ScalTaskF64<3, int>;

}  // namespace LegionSolvers
