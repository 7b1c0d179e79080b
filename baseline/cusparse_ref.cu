// cusparse_ref.cu -- COMPARATOR ONLY (never linked into or loaded by legionsolvers_b200).
//
// The GPU call sequence of the reference's leaf tasks, issued through cuSPARSE / cuBLAS 12.9 exactly as the
// reference issues it, on the reference's own data layout (fp64 entries, int64 column ids, rowptr = inclusive
// Rect<1> of GLOBAL k, x shifted to global column 0).  bench.py times it next to the hand-written kernels so that
// "faster than the library the reference runs" is a measured statement:
//
//   CSRMatvecTask::cuda_task_body   src/CSRMatrixTasks.cu:88-155 + src/CuSPARSEHelpers.hpp:9-101,162-201
//       per mat-vec: scratch indptr[rows + 1] (Legion::DeferredBuffer -> cudaMallocAsync here),
//       convertGlobalRowptrToLocalIndPtr (128 threads per block), cusparseCreateCsr (64-bit indices),
//       two cusparseCreateDnVec, cusparseSpMV_bufferSize, scratch workspace, cusparseSpMV(ALG_DEFAULT, beta = 0),
//       three destroys
//   ScalTask / AxpyTask / DotTask   src/LinearAlgebraTasks.cu:14-56, 59-113, 179-238   (cublasD*, host-pointer mode;
//       the dot synchronises the stream to return its value)
//   XpayTask                        src/LinearAlgebraTasks.cu:118-176                  (one element per thread)
//   CGSolver::step                  src/CGSolver.hpp:46-55 through SquarePlanner (zero fill before the mat-vec,
//       src/SquarePlanner.hpp:340-357); alpha folded on the host from the dot results (get_alpha,
//       src/LegionUtilities.cpp:72-97)
//
// Not modelled: Legion's own task-launch and mapping overhead (the comparator is what one GPU processor executes).
#include <cublas_v2.h>
#include <cuda_runtime.h>
#include <cusparse.h>
#include <stdint.h>
#include <stdio.h>

#define REF_THREADS 128  // THREADS_PER_BLOCK of the reference (src/CudaLibs.hpp:15)

struct ref_rect {
    long long lo, hi;
};

struct ref_ctx {
    cusparseHandle_t sp;
    cublasHandle_t bl;
};

#define REF_CUDA(x)                                                      \
    do {                                                                 \
        cudaError_t e_ = (x);                                            \
        if (e_ != cudaSuccess) {                                         \
            fprintf(stderr, "cusparse_ref: %s -> %s\n", #x, cudaGetErrorString(e_)); \
            return 1000 + (int) e_;                                      \
        }                                                                \
    } while (0)
#define REF_SP(x)                                                        \
    do {                                                                 \
        cusparseStatus_t s_ = (x);                                       \
        if (s_ != CUSPARSE_STATUS_SUCCESS) {                             \
            fprintf(stderr, "cusparse_ref: %s -> %d\n", #x, (int) s_);   \
            return 2000 + (int) s_;                                      \
        }                                                                \
    } while (0)
#define REF_BL(x)                                                        \
    do {                                                                 \
        cublasStatus_t s_ = (x);                                         \
        if (s_ != CUBLAS_STATUS_SUCCESS) {                               \
            fprintf(stderr, "cusparse_ref: %s -> %d\n", #x, (int) s_);   \
            return 3000 + (int) s_;                                      \
        }                                                                \
    } while (0)

// the reference's conversion kernel: Rect rowptr of global k -> 0-based indptr of this piece
__global__ void ref_rowptr_to_indptr(size_t rows, const ref_rect *rowptr, long long *indptr) {
    const size_t idx = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows) return;
    indptr[idx] = rowptr[idx].lo - rowptr[0].lo;
    if (idx == 0) indptr[rows] = rowptr[rows - 1].hi + 1 - rowptr[0].lo;
}

// the reference's xpay kernel: one element per thread, y = fma(alpha, y, x)
__global__ void ref_xpay(size_t n, double alpha, double *y, const double *x) {
    const size_t idx = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    y[idx] = fma(alpha, y[idx], x[idx]);
}

extern "C" {

int ref_create(ref_ctx **out) {
    ref_ctx *c = new ref_ctx();
    REF_SP(cusparseCreate(&c->sp));
    REF_BL(cublasCreate(&c->bl));
    REF_BL(cublasSetPointerMode(c->bl, CUBLAS_POINTER_MODE_HOST));
    *out = c;
    return 0;
}

int ref_destroy(ref_ctx *c) {
    if (!c) return 0;
    cusparseDestroy(c->sp);
    cublasDestroy(c->bl);
    delete c;
    return 0;
}

// CSRMatvecTask::cuda_task_body: y = A x (beta = 0); `cols` = input_domain.hi + 1, x_shifted indexable by global column
int ref_csr_matvec(ref_ctx *c, cudaStream_t st, int64_t rows, int64_t cols, int64_t nnz, const ref_rect *rowptr, const int64_t *col,
                   const double *entry, const double *x_shifted, double *y) {
    if (rows == 0) return 0;
    REF_SP(cusparseSetStream(c->sp, st));
    long long *indptr = nullptr;
    REF_CUDA(cudaMallocAsync(&indptr, sizeof(long long) * (size_t) (rows + 1), st));
    const unsigned blocks = (unsigned) ((rows + REF_THREADS - 1) / REF_THREADS);
    ref_rowptr_to_indptr<<<blocks, REF_THREADS, 0, st>>>((size_t) rows, rowptr, indptr);
    cusparseSpMatDescr_t A;
    REF_SP(cusparseCreateCsr(&A, rows, cols, nnz, indptr, const_cast<int64_t *>(col), const_cast<double *>(entry), CUSPARSE_INDEX_64I,
                             CUSPARSE_INDEX_64I, CUSPARSE_INDEX_BASE_ZERO, CUDA_R_64F));
    cusparseDnVecDescr_t vx, vy;
    REF_SP(cusparseCreateDnVec(&vx, cols, const_cast<double *>(x_shifted), CUDA_R_64F));
    REF_SP(cusparseCreateDnVec(&vy, rows, y, CUDA_R_64F));
    const double alpha = 1.0, beta = 0.0;
    size_t buf_size = 0;
    REF_SP(cusparseSpMV_bufferSize(c->sp, CUSPARSE_OPERATION_NON_TRANSPOSE, &alpha, A, vx, &beta, vy, CUDA_R_64F, CUSPARSE_SPMV_ALG_DEFAULT, &buf_size));
    void *workspace = nullptr;
    if (buf_size > 0) REF_CUDA(cudaMallocAsync(&workspace, buf_size, st));
    REF_SP(cusparseSpMV(c->sp, CUSPARSE_OPERATION_NON_TRANSPOSE, &alpha, A, vx, &beta, vy, CUDA_R_64F, CUSPARSE_SPMV_ALG_DEFAULT, workspace));
    REF_SP(cusparseDestroyDnVec(vx));
    REF_SP(cusparseDestroyDnVec(vy));
    REF_SP(cusparseDestroySpMat(A));
    if (workspace) REF_CUDA(cudaFreeAsync(workspace, st));
    REF_CUDA(cudaFreeAsync(indptr, st));
    return 0;
}

int ref_dot(ref_ctx *c, cudaStream_t st, int64_t n, const double *v, const double *w, double *host_out) {
    REF_BL(cublasSetStream(c->bl, st));
    REF_BL(cublasDdot(c->bl, (int) n, v, 1, w, 1, host_out));
    REF_CUDA(cudaStreamSynchronize(st));  // DotTask returns the value: the task synchronises (src/LinearAlgebraTasks.cu:233-237)
    return 0;
}
int ref_axpy(ref_ctx *c, cudaStream_t st, int64_t n, double alpha, const double *x, double *y) {
    REF_BL(cublasSetStream(c->bl, st));
    REF_BL(cublasDaxpy(c->bl, (int) n, &alpha, x, 1, y, 1));
    return 0;
}
int ref_scal(ref_ctx *c, cudaStream_t st, int64_t n, double alpha, double *x) {
    REF_BL(cublasSetStream(c->bl, st));
    REF_BL(cublasDscal(c->bl, (int) n, &alpha, x, 1));
    return 0;
}
int ref_xpay(ref_ctx *, cudaStream_t st, int64_t n, double alpha, const double *x, double *y) {
    if (n == 0) return 0;
    ref_xpay<<<(unsigned) ((n + REF_THREADS - 1) / REF_THREADS), REF_THREADS, 0, st>>>((size_t) n, alpha, y, x);
    return 0;
}

// `iters` CGSolver::step()s on ONE piece (src/CGSolver.hpp:46-55), scalars on the host like Legion futures read by
// get_alpha.  rr_io: in = r.r of the current residual, out = r.r after the last step.
int ref_cg_steps(ref_ctx *c, cudaStream_t st, int iters, int64_t n, int64_t nnz, const ref_rect *rowptr, const int64_t *col,
                 const double *entry, double *sol, double *p, double *q, double *r, double *rr_io) {
    double rr_old = *rr_io;
    for (int it = 0; it < iters; ++it) {
        REF_CUDA(cudaMemsetAsync(q, 0, sizeof(double) * (size_t) n, st));           // SquarePlanner::matvec zero-fills dst
        int rc = ref_csr_matvec(c, st, n, n, nnz, rowptr, col, entry, p, q);         // matvec(Q, P)
        if (rc) return rc;
        double p_norm = 0.0, rr_new = 0.0;
        if ((rc = ref_dot(c, st, n, p, q, &p_norm))) return rc;                      // dot(P, Q)
        if ((rc = ref_axpy(c, st, n, rr_old / p_norm, p, sol))) return rc;           // axpy(SOL, rr_old, p_norm, P)
        if ((rc = ref_axpy(c, st, n, (-1.0 * rr_old) / p_norm, q, r))) return rc;    // axpy(R, -1, rr_old, p_norm, Q)
        if ((rc = ref_dot(c, st, n, r, r, &rr_new))) return rc;                      // dot(R, R)
        if ((rc = ref_xpay(c, st, n, rr_new / rr_old, r, p))) return rc;             // xpay(P, rr_new, rr_old, R)
        rr_old = rr_new;
    }
    *rr_io = rr_old;
    return 0;
}

}  // extern "C"
