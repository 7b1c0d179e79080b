/*
 * lsk.h -- C ABI of the B200-native Krylov inner-loop kernels (liblsk.so).
 *
 * This is the drop-in boundary for the reference's GPU leaf tasks.  Each entry point replaces the
 * cuBLAS / cuSPARSE / hand-kernel call made by one `cuda_task_body` of dzhang314/LegionSolvers
 * (reference file:line cited per function) and takes exactly what that task body already holds:
 * raw device pointers borrowed from Legion physical instances (`accessor.ptr(domain.lo())`),
 * extents, and the task's CUDA stream.  INTEGRATION.md shows the binding for each task.
 *
 * Conventions (mirroring SURVEY.md section 8b):
 *   - plain C types only; every function returns 0 on success, a cudaError_t value (>0) for a CUDA
 *     failure, or a negative LSK_E_* code for a contract violation.  Nothing is thrown, nothing
 *     aborts; the reference's CHECK_* macros (src/CUDAUtilities.cpp:13-63) can wrap the result.
 *   - no entry point synchronises the stream or the device, allocates, or frees.  All scratch
 *     lives in the per-GPU `lsk_ctx` (the replacement for CUDALibraryContext,
 *     src/CUDAUtilities.hpp:44-66).  Pointers are borrowed for the duration of the call only.
 *   - one host thread per context at a time (Realm runs one task at a time per GPU processor).
 *   - indices are signed 64-bit (`long long` = Legion::coord_t), the only index type the reference
 *     registers mat-vec tasks for (src/Initialize.cpp:463-482).  Values are fp64 (`_f64`) or fp32
 *     (`_f32`).  Field data is SoA and stride-1, any 8-byte (4-byte for f32) alignment is accepted;
 *     16/32-byte alignment (src/LegionSolversMapper.cpp:71-88) enables the 256-bit load paths.
 *   - scalars are DEVICE-RESIDENT: where a reference task folds `task->futures` into alpha with
 *     get_alpha (src/LegionUtilities.cpp:72-97), the kernel takes up to four device pointers and
 *     folds them in its prologue with the same association:
 *        n=0 -> 1;  n=1 -> f0;  n=2 -> f0/f1;  n=3 -> (f0*f1)/f2;  n=4 -> (f0*f1)/(f2*f3).
 *     Dot products are written to a device double; the host never waits on them.
 *   - there is no CPU fallback: without a CUDA device every entry point fails.
 */
#ifndef LSK_H
#define LSK_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LSK_VERSION 100

/* negative = contract violation detected on the host */
#define LSK_E_INVALID (-1)   /* null pointer, negative extent, bad enum */
#define LSK_E_NO_DEVICE (-2) /* no CUDA device / context creation failed */
#define LSK_E_CAPACITY (-3)  /* per-context scratch exhausted */
#define LSK_E_NCCL (-4)      /* NCCL failure (lsk_comm_*) */

typedef struct lsk_ctx lsk_ctx;
/* cudaStream_t, passed as an opaque pointer so the header needs no CUDA include */
typedef void *lsk_stream;

/* Legion::Rect<1, long long>: INCLUSIVE bounds (rowptr field of CSRMatrix, src/CSRMatrix.hpp:24-26) */
typedef struct {
    int64_t lo, hi;
} lsk_rect;

/* ------------------------------------------------------------------------------------------------
 * Context -- replaces CUDALibraryContext / LoadCUDALibsTask (src/CUDAUtilities.cpp:66-145,
 * src/CudaLibs.cu:11-66): one per GPU processor, created once, holds reduction scratch.
 * ---------------------------------------------------------------------------------------------- */
int lsk_version(void);
const char *lsk_error_string(int status);
int lsk_ctx_create(int device, lsk_ctx **out);
int lsk_ctx_destroy(lsk_ctx *ctx);
int lsk_ctx_device(const lsk_ctx *ctx);
int lsk_ctx_sm_count(const lsk_ctx *ctx);
/* number of kernels launched through this context since creation (bench.py's gpu_launches claim) */
uint64_t lsk_ctx_launch_count(const lsk_ctx *ctx);
/* device pointers to the constants 1.0, -1.0, 0.0 (Scalar(ctx, rt, value), src/Scalar.hpp:30-33) */
const double *lsk_ctx_const_f64(const lsk_ctx *ctx, int which /*0: 1.0, 1: -1.0, 2: 0.0*/);

/* ------------------------------------------------------------------------------------------------
 * BLAS-1 leaf tasks (src/LinearAlgebraTasks.cu).  `n` = domain.get_volume() as int64 -- the
 * reference narrows it to `int` for cuBLAS; these kernels do not.
 * ---------------------------------------------------------------------------------------------- */

/* ScalTask::cuda_task_body (cublasDscal, src/LinearAlgebraTasks.cu:14-56): x = alpha * x */
int lsk_scal_f64(lsk_ctx *ctx, lsk_stream s, int64_t n, int n_terms, const double *f0,
                 const double *f1, const double *f2, const double *f3, double *x);
/* AxpyTask::cuda_task_body (cublasDaxpy, :59-113): y = fma(alpha, x, y) */
int lsk_axpy_f64(lsk_ctx *ctx, lsk_stream s, int64_t n, int n_terms, const double *f0,
                 const double *f1, const double *f2, const double *f3, const double *x, double *y);
/* XpayTask::cuda_task_body (xpay_kernel, :118-176): y = fma(alpha, y, x) */
int lsk_xpay_f64(lsk_ctx *ctx, lsk_stream s, int64_t n, int n_terms, const double *f0,
                 const double *f1, const double *f2, const double *f3, const double *x, double *y);
/* DotTask::cuda_task_body (cublasDdot + cudaStreamSynchronize, :179-238): *out = sum v[i]*w[i].
 * `out` is a DEVICE pointer; the reduction is two-stage with a fixed order (deterministic). */
int lsk_dot_f64(lsk_ctx *ctx, lsk_stream s, int64_t n, const double *v, const double *w, double *out);

int lsk_scal_f32(lsk_ctx *ctx, lsk_stream s, int64_t n, int n_terms, const float *f0,
                 const float *f1, const float *f2, const float *f3, float *x);
int lsk_axpy_f32(lsk_ctx *ctx, lsk_stream s, int64_t n, int n_terms, const float *f0,
                 const float *f1, const float *f2, const float *f3, const float *x, float *y);
int lsk_xpay_f32(lsk_ctx *ctx, lsk_stream s, int64_t n, int n_terms, const float *f0,
                 const float *f1, const float *f2, const float *f3, const float *x, float *y);
int lsk_dot_f32(lsk_ctx *ctx, lsk_stream s, int64_t n, const float *v, const float *w, float *out);

/* IndexFill (PartitionedVector::constant_fill, src/PartitionedVector.cpp:150-173) */
int lsk_fill_f64(lsk_ctx *ctx, lsk_stream s, int64_t n, double value, double *x);
int lsk_fill_f32(lsk_ctx *ctx, lsk_stream s, int64_t n, float value, float *x);
/* fill from a device scalar (constant_fill(const Scalar&)) */
int lsk_fill_dev_f64(lsk_ctx *ctx, lsk_stream s, int64_t n, const double *value, double *x);
/* IndexCopy (PartitionedVector::operator=, src/PartitionedVector.cpp:176-192): dst = src */
int lsk_copy_f64(lsk_ctx *ctx, lsk_stream s, int64_t n, const double *src, double *dst);

/* ------------------------------------------------------------------------------------------------
 * Scalar futures (src/Scalar.cpp:15-93 -> tasks in src/UtilityTasks.cpp:34-99), device-resident:
 * one single-thread kernel instead of one CPU task launch.
 * ---------------------------------------------------------------------------------------------- */
enum lsk_scalar_op {
    LSK_OP_NEG = 0,  /* NegateScalarTask    */
    LSK_OP_ADD = 1,  /* AddScalarTask       */
    LSK_OP_SUB = 2,  /* SubtractScalarTask  */
    LSK_OP_MUL = 3,  /* MultiplyScalarTask  */
    LSK_OP_DIV = 4,  /* DivideScalarTask    */
    LSK_OP_SQRT = 5, /* SqrtScalarTask      */
    LSK_OP_RSQRT = 6,/* RSqrtScalarTask: 1 / sqrt(x) */
    LSK_OP_DUMMY = 7,/* DummyTask: returns 1 */
    LSK_OP_COPY = 8
};
/* *out = op(*a, *b); b is ignored (may be NULL) for the unary ops */
int lsk_scalar_op_f64(lsk_ctx *ctx, lsk_stream s, int op, const double *a, const double *b, double *out);
int lsk_scalar_op_f32(lsk_ctx *ctx, lsk_stream s, int op, const float *a, const float *b, float *out);
/* residual_norm_squared.push_back(...) (src/CGSolver.hpp:53) with a device-side length, so that a
 * recorded trace (CUDA graph) can be replayed: hist[*count % capacity] = *value; ++*count; and, when
 * `also` is not NULL, *also = *value.  The history is a circular buffer (the reference's TODO at
 * src/CGSolver.hpp:25-26). */
int lsk_scalar_append_f64(lsk_ctx *ctx, lsk_stream s, const double *value, double *hist, int64_t capacity,
                          int64_t *count, double *also);

/* ------------------------------------------------------------------------------------------------
 * CSR mat-vec -- replaces CSRMatvecTask::cuda_task_body (src/CSRMatrixTasks.cu:14-156) together
 * with convertGlobalRowptrToLocalIndPtr + makeCuSparseCSR (src/CuSPARSEHelpers.hpp:9-101): reads
 * the Legion layout directly, no indptr conversion, no descriptor churn, no workspace.
 *
 *   y[r] = sum_{k in rowptr[r]} entry[k - k_base] * x_shifted[col[k - k_base]],  r = 0..rows-1
 *   (beta = 0: y is overwritten, as cusparseSpMV is called with beta = 0, :118-119)
 *
 *   rows       output_domain.volume()                       (:73)
 *   nnz        csr_matrix.get_bounds().volume()             (kernel piece)
 *   entry,col  entry_reader.ptr(kernel.lo), col_reader.ptr(kernel.lo)
 *   rowptr     rowptr_reader.ptr(rowptr_domain.lo): `rows` inclusive rects of GLOBAL k
 *   k_base     kernel_domain.lo[0], the global k of entry[0]
 *   x_shifted  input_reader.ptr(input.lo) - input.lo[0]: indexable by GLOBAL column id -- the
 *              same shifted pointer makeShiftedCuSparseDnVec builds (src/CuSPARSEHelpers.hpp:188-201)
 *   y          output_writer.ptr(output.lo)
 *   dot_w/dot_out  optional fusion (both NULL to disable): *dot_out = sum_r y[r] * dot_w[r], the
 *              p.Ap of CGSolver::step (src/CGSolver.hpp:47-48) without a second pass over y.
 *   dot_yy_out optional (NULL to disable): *dot_yy_out = sum_r y[r] * y[r]  (BiCGStab's u.u,
 *              src/BiCGStabSolver.hpp:76).
 *   variant    LSK_SPMV_AUTO picks by mean row length nnz/rows.
 *
 * LSK_SPMV_STREAM adds the rounded products of a row in ascending k, exactly the order of the
 * reference's CPU body (src/CSRMatrixTasks.cpp:73-91): its result is bit-identical to it.
 * The LANES/VECTOR/WARP variants reduce a row across lanes (fixed tree order): <= 1e-12 relative.
 * AUTO: mean row length <= 12 -> STREAM (so the 5- and 7-point stencils stay bit-exact), <= 96 -> LANES, else WARP.
 * ---------------------------------------------------------------------------------------------- */
enum lsk_spmv_variant {
    LSK_SPMV_AUTO = 0,
    LSK_SPMV_STREAM = 1, /* block streams a contiguous run of non-zeros through shared memory */
    LSK_SPMV_VECTOR = 2, /* 2..16 lanes per row */
    LSK_SPMV_WARP = 3,   /* one warp per row */
    LSK_SPMV_LANES = 4,  /* the STREAM kernel's TMA-staged tiles with 2, 4 or 8 lanes per row (by mean row length) */
    /* flag, OR-ed into `variant`: y = y + A x instead of y = A x.  For a further operator block on rows that an earlier
     * block of the same planner mat-vec has already written (the reference's CPU bodies accumulate onto the zero-filled
     * destination through a sum-reduction accessor, src/CSRMatrixTasks.cpp:29-31; its cuSPARSE variant overwrites,
     * beta = 0, which is only right for one block per range space).  STREAM: the products are added one by one onto
     * the row's current value, in ascending k -- bit-identical to the CPU body. */
    LSK_SPMV_ACCUMULATE = 0x100
};
int lsk_csr_spmv_f64(lsk_ctx *ctx, lsk_stream s, int64_t rows, int64_t nnz, const double *entry,
                     const int64_t *col, const lsk_rect *rowptr, int64_t k_base,
                     const double *x_shifted, double *y, const double *dot_w, double *dot_out,
                     double *dot_yy_out, int variant);
int lsk_csr_spmv_f32(lsk_ctx *ctx, lsk_stream s, int64_t rows, int64_t nnz, const float *entry,
                     const int64_t *col, const lsk_rect *rowptr, int64_t k_base,
                     const float *x_shifted, float *y, const float *dot_w, float *dot_out,
                     float *dot_yy_out, int variant);
/* which variant AUTO resolves to for this shape (for reporting) */
int lsk_csr_spmv_pick(int64_t rows, int64_t nnz);

/* ------------------------------------------------------------------------------------------------
 * COO mat-vec -- replaces COOMatvecTask::cuda_task_body (src/COOMatrixTasks.cu:12-146):
 *   y_shifted[row[k]] += entry[k] * x_shifted[col[k]]   for k = 0..nnz-1   (beta = 1, :102-108)
 * guarded like the CPU body (src/COOMatrixTasks.cpp:70-73) by row in [row_lo,row_hi] and col in
 * [col_lo,col_hi].  Both vectors are shifted to global index 0 (:78-99).  Segmented warp-shuffle
 * reduction over runs of equal row; run ends are combined with fp64 atomics.
 * ---------------------------------------------------------------------------------------------- */
int lsk_coo_spmv_f64(lsk_ctx *ctx, lsk_stream s, int64_t nnz, const double *entry,
                     const int64_t *row, const int64_t *col, const double *x_shifted,
                     double *y_shifted, int64_t row_lo, int64_t row_hi, int64_t col_lo,
                     int64_t col_hi);
int lsk_coo_spmv_f32(lsk_ctx *ctx, lsk_stream s, int64_t nnz, const float *entry,
                     const int64_t *row, const int64_t *col, const float *x_shifted,
                     float *y_shifted, int64_t row_lo, int64_t row_hi, int64_t col_lo,
                     int64_t col_hi);

/* ------------------------------------------------------------------------------------------------
 * Transposed mat-vecs -- CSRRmatvecTask / COORmatvecTask (TaskIDs reserved at src/TaskIDs.hpp:40-45, bodies
 * `assert(false)` in the reference: src/CSRMatrixTasks.cpp:94-100, src/COOMatrixTasks.cpp:77-83):
 *   y_shifted[col[k]] += entry[k] * x[row of k]      for every stored non-zero k of the piece
 * accumulating (the caller zero-fills, as SquarePlanner::matvec does), guarded by col in [col_lo, col_hi] (and, COO,
 * row in [row_lo, row_hi]).  CSR: x is the piece's own rows (x[0] <-> first row of `rowptr`); COO: x_shifted is
 * indexable by GLOBAL row.  y_shifted is indexable by GLOBAL column.  fp64 atomics: <= 1e-12, not bit-reproducible.
 * ---------------------------------------------------------------------------------------------- */
int lsk_csr_rspmv_f64(lsk_ctx *ctx, lsk_stream s, int64_t rows, int64_t nnz, const double *entry, const int64_t *col,
                      const lsk_rect *rowptr, int64_t k_base, const double *x, double *y_shifted, int64_t col_lo,
                      int64_t col_hi);
int lsk_coo_rspmv_f64(lsk_ctx *ctx, lsk_stream s, int64_t nnz, const double *entry, const int64_t *row,
                      const int64_t *col, const double *x_shifted, double *y_shifted, int64_t row_lo, int64_t row_hi,
                      int64_t col_lo, int64_t col_hi);

/* ------------------------------------------------------------------------------------------------
 * The end of a GMRES restart cycle, finished.  The reference stops at a placeholder (DummyTask returns 1:
 * SOL += 1 * v_j, src/GMRESSolver.hpp:109-126); these two are what the algorithm needs instead:
 *   lsk_gmres_solve_f64   y = argmin || sqrt(*beta_sq) e1 - H y ||_2 for the (m + 1) x m upper Hessenberg H (row-major,
 *                         leading dimension ld, DEVICE memory) by Givens rotations; *resid (optional) = the minimum,
 *                         i.e. the norm of the new residual.  m <= 64.  One thread: it is ~m^2 flops.
 *   lsk_multi_axpy_f64    x = x + sum_j y[j] V[j], added in ascending j with one fma per term -- bit-identical to m
 *                         AxpyTasks -- in ONE pass: 8 (m + 2) instead of 24 m bytes per element.  y and the pointer
 *                         table V live on the device.
 * ---------------------------------------------------------------------------------------------- */
int lsk_gmres_solve_f64(lsk_ctx *ctx, lsk_stream s, int m, const double *H, int ld, const double *beta_sq, double *y,
                        double *resid);
int lsk_multi_axpy_f64(lsk_ctx *ctx, lsk_stream s, int64_t n, int m, const double *y, const double *const *V, double *x);

/* ------------------------------------------------------------------------------------------------
 * Fused solver passes.  Each equals a fixed sequence of the leaf tasks above with identical
 * element-wise arithmetic (same fma / rounding per element), saving HBM passes.
 * ---------------------------------------------------------------------------------------------- */

/* CGSolver::step lines src/CGSolver.hpp:50-52 in one pass:
 *   x = fma(rr_old/pq, p, x);  r = fma((-1*rr_old)/pq, q, r);  *rr_new = sum r*r            */
int lsk_cg_update_f64(lsk_ctx *ctx, lsk_stream s, int64_t n, const double *rr_old, const double *pq,
                      const double *p, const double *q, double *x, double *r, double *rr_new);

/* y = fma(alpha, x, y) then *out = sum y*w  (axpy followed by dot; w may alias y).
 * GMRES modified Gram-Schmidt (src/GMRESSolver.hpp:95-101) and BiCGStab use it. */
int lsk_axpy_dot_f64(lsk_ctx *ctx, lsk_stream s, int64_t n, int n_terms, const double *f0,
                     const double *f1, const double *f2, const double *f3, const double *x,
                     double *y, const double *w, double *out);

/* two dots sharing one pass: *out_vw = sum v*w, *out_ww = sum w*w
 * (BiCGStabSolver::step r.u and u.u, src/BiCGStabSolver.hpp:75-76) */
int lsk_dot2_f64(lsk_ctx *ctx, lsk_stream s, int64_t n, const double *v, const double *w,
                 double *out_vw, double *out_ww);

/* BiCGStab direction update (src/BiCGStabSolver.hpp:68-69) in one pass:
 *   p = fma(-omega, v, p);  p = fma(beta, p, r)   with beta = (rho_new/rho_old)*(alpha/omega)   */
int lsk_bicg_p_update_f64(lsk_ctx *ctx, lsk_stream s, int64_t n, const double *rho_new,
                          const double *rho_old, const double *alpha, const double *omega,
                          const double *v, const double *r, double *p);

/* BiCGStab tail (src/BiCGStabSolver.hpp:78-80) in one pass, omega = ru/uu:
 *   x = fma(alpha, p, x); x = fma(omega, r, x); r = fma(-omega, u, r); *rho_next = sum r*rt   */
int lsk_bicg_tail_f64(lsk_ctx *ctx, lsk_stream s, int64_t n, const double *alpha, const double *ru,
                      const double *uu, const double *p, const double *u, const double *rt,
                      double *x, double *r, double *rho_next);


/* ------------------------------------------------------------------------------------------------
 * Set-up side: the integer work that runs once before the hot path,
 * moved to the GPU.  All results are exact integers; parity with the oracle is bit-exact.
 * ---------------------------------------------------------------------------------------------- */

#define LSK_MAX_DIM 3
#define LSK_MAX_STENCIL 64

/* A stencil on a dense DIM-dimensional grid, as FillLinearizedCSRStencilTask receives it
 * (src/StencilGenerator.hpp:100-137): `shape[d]` points per dimension (bounds.lo = 0), `noff`
 * offsets with their entries ALREADY SORTED the way the task sorts them
 * (src/StencilGenerator.cpp:408-433; lsk_stencil_sort does it), order 0 = ROW_MAJOR, 1 = COLUMN_MAJOR. */
typedef struct {
    int dim;
    int order;
    int noff;
    int64_t shape[LSK_MAX_DIM];
    int64_t offsets[LSK_MAX_STENCIL][LSK_MAX_DIM];
    double values[LSK_MAX_STENCIL];
} lsk_stencil;

/* host-only helpers: the sort of the fill tasks, and calculate_stencil_size
 * (src/StencilGenerator.hpp:270-323, closed form: sum over offsets of prod_d max(0, shape_d - |o_d|)) */
int lsk_stencil_sort(lsk_stencil *st);
int64_t lsk_stencil_size(const lsk_stencil *st);

/* number of non-zeros in rows [r_lo, r_hi] (inclusive, linearised): *out is a DEVICE int64 */
int lsk_stencil_count_f64(lsk_ctx *ctx, lsk_stream s, const lsk_stencil *st, int64_t r_lo, int64_t r_hi,
                          int64_t *out);
/* FillLinearizedCSRStencilTask::task_body (src/StencilGenerator.cpp:380-543) for the slab of rows
 * [r_lo, r_hi]: k_first = global k of the slab's first non-zero (= count of rows [0, r_lo)).
 * Writes entry/col for k in [k_first, k_first + slab_nnz) at entry[0..], and rowptr[0..rows) with
 * inclusive GLOBAL-k rects.  scratch: rows + 1 device int64 (row start offsets). */
int lsk_stencil_fill_csr_f64(lsk_ctx *ctx, lsk_stream s, const lsk_stencil *st, int64_t r_lo, int64_t r_hi,
                             int64_t k_first, double *entry, int64_t *col, lsk_rect *rowptr,
                             int64_t *scratch);
/* row[k - k_base] = r for every k in rowptr[r - r_lo]  (CSR -> COO row field, as
 * FillLinearizedCOOStencilTask writes it, src/StencilGenerator.cpp:160-243) */
int lsk_csr_expand_rows(lsk_ctx *ctx, lsk_stream s, int64_t rows, int64_t r_lo, const lsk_rect *rowptr,
                        int64_t k_base, int64_t *row);

/* Dependent partitioning (Legion/Realm operations called at src/CSRMatrix.cpp:68-155 and
 * src/COOMatrix.cpp:56-141), on device-resident fields.  "span" results are 3 DEVICE int64:
 * {min, max, count}; min > max when empty.  "flags" results are one byte per point of the
 * parent space window, 1 = member. */
/* image_range over rowptr (src/CSRMatrix.cpp:89-109): union of `rows` rects */
int lsk_rect_span_i64(lsk_ctx *ctx, lsk_stream s, int64_t rows, const lsk_rect *rowptr, int64_t *out3);
int lsk_image_range_flags(lsk_ctx *ctx, lsk_stream s, int64_t rows, const lsk_rect *rowptr,
                          int64_t k_lo, int64_t k_n, uint8_t *kflags);
/* image of a point field over a kernel piece (src/CSRMatrix.cpp:112-132): values field[0..n) */
int lsk_minmax_i64(lsk_ctx *ctx, lsk_stream s, int64_t n, const int64_t *field, int64_t *out3);
int lsk_image_flags(lsk_ctx *ctx, lsk_stream s, int64_t n, const int64_t *field, const uint8_t *kflags,
                    int64_t out_lo, int64_t out_n, uint8_t *out_flags);
/* preimage of a point field (src/COOMatrix.cpp:77-96): { k : lo <= field[k] <= hi }, k = k_base + i */
int lsk_preimage_span_i64(lsk_ctx *ctx, lsk_stream s, int64_t n, const int64_t *field, int64_t lo,
                          int64_t hi, int64_t k_base, int64_t *out3);
int lsk_preimage_flags(lsk_ctx *ctx, lsk_stream s, int64_t n, const int64_t *field, int64_t lo,
                       int64_t hi, uint8_t *kflags);
/* preimage_range (src/CSRMatrix.cpp:135-155): rows whose rect meets the kernel piece kflags
 * (kflags[i] describes k = k_lo + i) */
int lsk_preimage_range_flags(lsk_ctx *ctx, lsk_stream s, int64_t rows, const lsk_rect *rowptr,
                             int64_t k_lo, int64_t k_n, const uint8_t *kflags, uint8_t *rflags);
/* create_equal_partition (host arithmetic): piece i of `pieces` over [0, n) */
int lsk_equal_partition(int64_t n, int pieces, int64_t *lo, int64_t *hi);
/* BlockingShardingFunctor::shard (src/LegionSolversMapper.cpp:140-151) */
int lsk_shard(int64_t point, int64_t volume, int64_t total_shards);


/* ------------------------------------------------------------------------------------------------
 * Collectives over NVLink / NVSwitch PEER MEMORY (one process per GPU, windows mapped with CUDA IPC).
 * The Krylov path has exactly two exchange steps: the ghost-x halo before a
 * mat-vec and the sum of per-rank dot partials.  Both are latency-bound (<= 512 KB, 8 bytes), so they
 * ride inside the kernels that produce the data and use NO FENCE AND NO FLAG: every 8-byte word that
 * crosses NVLink carries 32 data bits and the 32-bit number of the exchange it belongs to ("LL packet";
 * an aligned 8-byte store is atomic), so a packet both delivers its data and announces it.  A system-scope
 * fence costs 1.9 us on an idle B200 and 13-20 us while a PCIe copy is in flight (tools/probe_fence.cu);
 * an LL round trip costs 2.5 us either way.  Every rank must call the collectives the same number of
 * times in the same order (SPMD), on the stream that orders them with the producers / consumers.
 *
 * `lsk_peers.window[r]` is rank r's comm window (lsk_comm_window_bytes() bytes of zeroed device
 * memory, cudaMalloc'ed) as mapped INTO THE CALLING PROCESS; window[rank] is the local one.
 * ---------------------------------------------------------------------------------------------- */
/* Limits, stated rather than silent: at most LSK_MAX_RANKS ranks share peer-memory windows (one NVSwitch domain; the host
 * layer falls back to NCCL beyond that).  Packet tags are the low 32 bits of 64-bit exchange counters: a tag is only ever
 * compared for EQUALITY with the one expected next, and the slot it is looked for in holds the tag of two exchanges
 * earlier at worst, so the wrap after 2^32 exchanges (days of solver iterations) is harmless. */
#define LSK_MAX_RANKS 16
typedef struct {
    int rank, nranks;
    void *window[LSK_MAX_RANKS];
} lsk_peers;
size_t lsk_comm_window_bytes(void);
/* slots[0..count) (count <= 2) := sum over ranks, identical bits on every rank (rank-order sum) */
int lsk_allreduce_sum_f64(lsk_ctx *ctx, lsk_stream s, const lsk_peers *peers, double *slots, int count);
/* One halo move = what this rank trades with ONE peer in an exchange: `n` doubles at local `src` go to the peer, and
 * `recv_n` doubles from the peer end up at local `recv_dst` (this rank's ghost region).  The values travel as LL
 * packets into a LANDING BUFFER owned by the receiver -- lsk_halo_landing_bytes(count) bytes of zeroed device memory per
 * (receiver, sender) pair, 32-byte aligned, holding two exchanges (alternating) of count + 1 packets of 16 bytes -- and the RECEIVER copies
 * them into its ghost region: ghost values are only ever written by the rank that reads them, in stream order.
 *   ll_send  the peer's landing buffer for this rank's packets, as mapped into this process (sized for n)
 *   ll_recv  this rank's landing buffer for the peer's packets (sized for recv_n)
 * Both ranks of a pair must list each other in the same exchanges, even when one direction is empty (n == 0 or
 * recv_n == 0): the count + 1st packet is a token that travels in both directions, which bounds how far one rank can
 * run ahead of the other (at most one exchange: the two halves of a landing buffer are never overwritten while in use). */
typedef struct {
    int peer;
    int reserved;
    const double *src;
    int64_t n;
    void *ll_send;
    double *recv_dst;
    int64_t recv_n;
    void *ll_recv;
} lsk_halo_move;
#define LSK_MAX_HALO_MOVES 32
size_t lsk_halo_landing_bytes(int64_t count);
/* neighbour exchange with barrier semantics: when the kernel completes on this stream, every peer's data destined for
 * this rank is in place at recv_dst (and this rank's packets are on their way or have landed). */
int lsk_halo_exchange_f64(lsk_ctx *ctx, lsk_stream s, const lsk_peers *peers, const lsk_halo_move *moves,
                          int nmoves);
/* the same exchange, except that the values received are ADDED to recv_dst: the REVERSE halo exchange of a transposed
 * mat-vec (a rank's contributions to columns it does not own travel to their owners; src = its ghost region, recv_dst =
 * its boundary rows).  Needs landing buffers of its own, sized for the reversed counts. */
int lsk_halo_reduce_f64(lsk_ctx *ctx, lsk_stream s, const lsk_peers *peers, const lsk_halo_move *moves, int nmoves);
/* FUSED forms: no launch of their own.
 * lsk_ctx_set_peers(ctx, peers): from now on EVERY reducing kernel launched through ctx (dot, dot2,
 * cg_update, axpy_dot, bicg_tail, the fused SpMV dots) finishes with the cross-rank sum in the tail of
 * its last CTA: the value it writes is already the global one.  Every rank must then launch the same
 * reducing kernels in the same order.  NULL switches back to rank-local reductions. */
int lsk_ctx_set_peers(lsk_ctx *ctx, const lsk_peers *peers);
/* DEFERRED form of the fused all-reduce, for a reduction whose only consumer is the next kernel on the stream (the p.q and
 * r.r of a fused CG step).  lsk_ctx_defer_next_allreduce(ctx) applies to the NEXT reducing launch through ctx, if it is
 * lsk_csr_spmv_f64 with dot_out only, or lsk_cg_update_f64: its last CTA only SENDS the rank's sum to
 * the peers (the value it stores to the output slot is the rank-local one) and the kernel ends; the cross-rank sum is formed
 * by the next lsk_cg_update_f64 whose `pq`, or lsk_cg_direction_f64 whose `rr_new`, is that same slot -- every CTA of it polls
 * the packets in its own window at its start, and the global value is stored back to the slot.  The NVLink flight and the
 * wait for the slowest rank then overlap the kernel boundary.  Any other entry point that reads device scalars, called in
 * between, first finishes the reduction with a one-warp kernel (so does lsk_ctx_settle), i.e. a wrong guess costs a launch,
 * never a wrong number.  No-op without lsk_ctx_set_peers. */
int lsk_ctx_defer_next_allreduce(lsk_ctx *ctx);
int lsk_ctx_settle(lsk_ctx *ctx, lsk_stream s);
/* XpayTask fused with the halo exchange of its result: y = fma(alpha, y, x); the elements of y inside moves[i].src[0..n)
 * (sub-ranges of y) also leave as packets for the neighbours while the pass runs, and the CTAs unpack the neighbours'
 * packets into moves[i].recv_dst when they have finished their share of the pass; when the kernel completes this rank's
 * ghosts of y are current.  Needs lsk_ctx_set_peers.  nmoves <= 4. */
int lsk_xpay_halo_f64(lsk_ctx *ctx, lsk_stream s, int64_t n, int n_terms, const double *f0, const double *f1,
                      const double *f2, const double *f3, const double *x, double *y, const lsk_halo_move *moves,
                      int nmoves);
/* CGSolver::step lines src/CGSolver.hpp:53-54 in one launch:
 *   residual_norm_squared.push_back(rr_new);  p = fma(rr_new/rr_cur, p, r)   [xpay(P, rr_new, rr_cur, R)]
 * then *rr_cur = *rr_new for the next step.  `history` may be NULL (no append); otherwise circular like
 * lsk_scalar_append_f64.  With `moves` (nmoves <= 4; needs lsk_ctx_set_peers) the halo of p is exchanged inside the
 * kernel exactly as lsk_xpay_halo_f64 does (boundary chunks are processed first, the unpacking comes last: the packets
 * have landed by then).  TMA-streamed: r and p must be 32-byte congruent (lsk_cg_direction_supported). */
int lsk_cg_direction_supported(int64_t n, const double *r, const double *p);
int lsk_cg_direction_f64(lsk_ctx *ctx, lsk_stream s, int64_t n, double *rr_cur, const double *rr_new, const double *r,
                         double *p, const lsk_halo_move *moves, int nmoves, double *history,
                         int64_t history_capacity, int64_t *history_count);

/* CGSolver::step lines src/CGSolver.hpp:50-54 in ONE launch -- lsk_cg_update_f64 followed by lsk_cg_direction_f64 -- for
 * vectors small enough to live in the L2 (the slab of a multi-GPU run): one CTA per SM owns the same elements in both
 * passes (its new r waits in shared memory), and the kernel boundary between them becomes a grid-wide wait on one word
 * that carries the global r.r (cross-rank sum included, with lsk_ctx_set_peers).  Same element-wise arithmetic; r.r is
 * folded in a fixed order of its own.  `pq` may be a deferred reduction of the preceding mat-vec
 * (lsk_ctx_defer_next_allreduce); `moves` as for lsk_cg_direction_f64.  lsk_cg_tail_supported: p, q, x, r 32-byte
 * congruent and below the size where the TMA-streamed kernels take over. */
int lsk_cg_tail_supported(lsk_ctx *ctx, int64_t n, const double *p, const double *q, const double *x, const double *r);
/* accounting kept by lsk_cg_tail_f64 (CTA 0's view, accumulated since context creation): ns in {p.q resolve, phase 1,
 * wait for the global r.r, phase 2, unpacking the neighbours' halo} and the number of launches.  Synchronises. */
int lsk_cg_tail_stats(lsk_ctx *ctx, lsk_stream s, uint64_t *host_out6);
int lsk_cg_tail_f64(lsk_ctx *ctx, lsk_stream s, int64_t n, double *rr_cur, const double *pq, double *rr_new, double *p,
                    const double *q, double *x, double *r, const lsk_halo_move *moves, int nmoves, double *history,
                    int64_t history_capacity, int64_t *history_count);

/* accounting kept in the comm window: {all-reduce calls, ns inside them, halo closes, ns inside them}
 * (time between entering the collective and leaving it, on the thread that closes it).  Synchronises. */
int lsk_comm_stats(lsk_ctx *ctx, lsk_stream s, const lsk_peers *peers, uint64_t *host_out4);
/* non-zero if a spin-wait in one of the collectives gave up (protocol violation / dead peer) */
int lsk_comm_error(lsk_ctx *ctx, lsk_stream s, const lsk_peers *peers, int *host_out);

#ifdef __cplusplus
}
#endif
#endif /* LSK_H */
