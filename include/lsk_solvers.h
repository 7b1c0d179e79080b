/*
 * lsk_solvers.h -- C ABI of the C++ host layer (legionsolvers_b200/host/ headers): the planner-level
 * API of the reference (PartitionedVector, CSRMatrix / COOMatrix, SquarePlanner, CGSolver,
 * BiCGStabSolver, GMRESSolver) as opaque handles, for callers that are not C++ (the Python tests
 * and bench.py drive everything through this file and lsk.h).
 *
 * Conventions: plain C types; 0 = success, otherwise a status from lsk.h; after a failure
 * lsk_last_error() describes it (thread-local).  Handles are created and destroyed explicitly;
 * destroy solvers before planners, planners before vectors / matrices, everything before the runtime.
 * "global" host arrays are indexed by GLOBAL row; a rank reads / writes only the rows it owns.
 */
#ifndef LSK_SOLVERS_H
#define LSK_SOLVERS_H

#include "lsk.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct lsk_runtime lsk_runtime;
typedef struct lsk_vector lsk_vector;
typedef struct lsk_matrix lsk_matrix;
typedef struct lsk_planner lsk_planner;
typedef struct lsk_solver lsk_solver;

const char *lsk_last_error(void);

/* ---- halo plan: the host arithmetic of SquarePlanner::add_row_partitioned_matrix (src/SquarePlanner.hpp:209-235 derives the
 * ghost partition; Realm then moves the ghost instances implicitly -- here the moves are explicit).  ranges4[r] = {first owned
 * row, last owned row, ghost lo, ghost hi} of rank r (inclusive; ghost lo > ghost hi = none).  Output: moves5[i] = {peer,
 * send_lo, send_n, recv_lo, recv_n} for every peer this rank trades with, *nmoves of them (<= nranks - 1).  No CUDA, no
 * runtime needed: the multi-process CPU tests call it directly. */
int lsk_halo_plan(int rank, int nranks, const int64_t *ranges4, int64_t *moves5, int *nmoves);

/* ---- runtime: one per process / GPU (replaces Legion::Runtime + mapper for this path) ------------ */
/* external_stream: a cudaStream_t to enqueue on, or NULL for a private stream */
int lsk_rt_create(int device, int rank, int nranks, void *external_stream, lsk_runtime **out);
int lsk_rt_destroy(lsk_runtime *rt);
/* NCCL bootstrap: rank 0 calls lsk_rt_unique_id, ships the 128 bytes to every rank, all call comm_init */
int lsk_rt_unique_id(void *out128);
int lsk_rt_comm_init(lsk_runtime *rt, const void *uid128);
/* 0: collectives on NCCL (LSK_COMM=nccl, or no peer access); 1: stand-alone peer-memory kernels (CUDA IPC
 * over NVLink); 2: peer-memory collectives FUSED into the producing kernels' tails (one piece per rank) */
int lsk_rt_uses_peer_memory(lsk_runtime *rt);
/* lsk_comm_stats of this runtime's window (zeros when not on peer memory) */
int lsk_rt_comm_stats(lsk_runtime *rt, uint64_t *out4);
/* non-zero if a peer-memory collective gave up waiting (synchronises) */
int lsk_rt_comm_error(lsk_runtime *rt, int *out);
lsk_ctx *lsk_rt_ctx(lsk_runtime *rt);
void *lsk_rt_stream(lsk_runtime *rt);
int lsk_rt_fence(lsk_runtime *rt);                 /* issue_execution_fence */
uint64_t lsk_rt_kernel_launches(lsk_runtime *rt);  /* graph replays included */
/* Legion begin_trace / end_trace (test/BenchmarkStencil.cpp:219-241): first use of an id records the
 * enclosed launches into a CUDA graph, later uses replay it */
int lsk_rt_begin_trace(lsk_runtime *rt, int trace_id);
int lsk_rt_end_trace(lsk_runtime *rt, int trace_id);

/* ---- PartitionedVector<double> (src/PartitionedVector.hpp) --------------------------------------------- */
int lsk_vector_create(lsk_runtime *rt, const char *name, int64_t volume, int pieces, lsk_vector **out);
int lsk_vector_destroy(lsk_vector *v);
int lsk_vector_owned_range(lsk_vector *v, int64_t *lo, int64_t *hi);
int lsk_vector_constant_fill(lsk_vector *v, double value);
int lsk_vector_assign(lsk_vector *dst, const lsk_vector *src);           /* operator= (IndexCopy) */
int lsk_vector_scal(lsk_vector *v, double alpha);
int lsk_vector_axpy(lsk_vector *y, double alpha, const lsk_vector *x);   /* y.axpy(alpha, x) */
int lsk_vector_xpay(lsk_vector *y, double alpha, const lsk_vector *x);   /* y.xpay(alpha, x) */
int lsk_vector_dot(const lsk_vector *v, const lsk_vector *w, double *out); /* synchronises (get_value) */
int lsk_vector_copy_from_host(lsk_vector *v, const double *global);
int lsk_vector_copy_to_host(const lsk_vector *v, double *global);

/* ---- matrices (src/CSRMatrix.hpp, src/COOMatrix.hpp) ------------------------------------------------------ */
/* upload the slab rows [r_lo, r_hi] / non-zeros [k_lo, k_hi] this rank needs; arrays start at the slab */
int lsk_csr_create(lsk_runtime *rt, int64_t rows, int64_t cols, int64_t nnz_global, int64_t r_lo, int64_t r_hi,
                   int64_t k_lo, int64_t k_hi, const double *entry, const int64_t *col, const lsk_rect *rowptr,
                   lsk_matrix **out);
int lsk_coo_create(lsk_runtime *rt, int64_t rows, int64_t cols, int64_t nnz_global, int64_t k_lo, int64_t k_hi,
                   const double *entry, const int64_t *row, const int64_t *col, lsk_matrix **out);
/* create_linearized_csr_stencil_matrix (src/StencilGenerator.hpp:533-643), filled on the GPU */
int lsk_csr_create_stencil(lsk_runtime *rt, const lsk_stencil *stencil, int pieces, lsk_matrix **out);
/* the BenchmarkStencil matrices: dim_flag 1, 2, 3, 4 (= 27-point) (test/BenchmarkStencil.cpp:33-131) */
int lsk_benchmark_stencil(int dim_flag, int64_t nx, int64_t ny, int64_t nz, lsk_stencil *out);
int lsk_matrix_destroy(lsk_matrix *m);
/* rows, cols, nnz_global, slab r_lo, r_hi, k_lo, k_hi, is_csr */
int lsk_matrix_info(const lsk_matrix *m, int64_t *out8);
/* copy the slab back (generator parity tests): third array is rowptr (CSR) or row (COO) */
int lsk_matrix_slab_to_host(const lsk_matrix *m, double *entry, int64_t *col, void *rowptr_or_row);
/* device pointers of the slab fields: entry, col, rowptr-or-row (for direct lsk.h kernel calls) */
int lsk_matrix_device_fields(const lsk_matrix *m, void **entry, void **col, void **rowptr_or_row);

/* ---- SquarePlanner<double> (src/SquarePlanner.hpp) ------------------------------------------------------------ */
int lsk_planner_create(lsk_runtime *rt, lsk_planner **out);
int lsk_planner_destroy(lsk_planner *pl);
int lsk_planner_add_sol_vector(lsk_planner *pl, lsk_vector *v);
int lsk_planner_add_rhs_vector(lsk_planner *pl, lsk_vector *v);
int lsk_planner_add_row_partitioned_matrix(lsk_planner *pl, const lsk_matrix *m, int domain_index, int range_index);
int lsk_planner_allocate_workspace(lsk_planner *pl, int num_vectors);
/* which: 0 = canonical partition of space `index`, 1 = kernel partition of block `index`,
 * 2 = ghost partition of block `index`; bounds are inclusive; only local colours are meaningful for 1, 2 */
int lsk_planner_partition_bounds(lsk_planner *pl, int which, int index, int color, int64_t *lo, int64_t *hi);
int lsk_planner_local_colors(lsk_planner *pl, int space, int *first, int *end);
uint64_t lsk_planner_halo_bytes_per_matvec(lsk_planner *pl);
/* vector-id operations (0 SOL, 1 RHS, 2.. workspace); scalars given by value */
int lsk_planner_zero_fill(lsk_planner *pl, int vec);
int lsk_planner_copy(lsk_planner *pl, int dst, int src);
int lsk_planner_scal(lsk_planner *pl, int dst, double alpha);
int lsk_planner_axpy(lsk_planner *pl, int dst, double alpha, int src);
int lsk_planner_xpay(lsk_planner *pl, int dst, double alpha, int src);
int lsk_planner_dot(lsk_planner *pl, int v, int w, double *out); /* synchronises */
int lsk_planner_matvec(lsk_planner *pl, int dst, int src);
/* dst = A^T src over all registered blocks (CSRRmatvecTask / COORmatvecTask: reserved, unimplemented in the reference).
 * Fails on several ranks when a local piece references columns owned by another rank (no reverse halo exchange yet). */
int lsk_planner_rmatvec(lsk_planner *pl, int dst, int src);
/* fused dst = A src with out_yw = dst . vec(w) (and out_yy = dst . dst if not NULL); synchronises */
int lsk_planner_matvec_dot(lsk_planner *pl, int dst, int src, int w, double *out_yw, double *out_yy);
int lsk_planner_vector_to_host(lsk_planner *pl, int vec, int space, double *global);
int lsk_planner_vector_from_host(lsk_planner *pl, int vec, int space, const double *global);
/* Asynchronous copies of the OWNED rows, `global + own_lo` <-> the vector, with cudaMemcpyAsync(cudaMemcpyDefault) on
 * `stream` (a cudaStream_t; NULL = the runtime's stream) and no synchronisation: `global` may be pinned host memory or
 * device memory.  On a stream other than the runtime's the caller orders the copy against the solver's work with
 * events -- this is how the I/O of one solve overlaps the iterations of the next (bench.py's end-to-end loop). */
int lsk_planner_vector_to_async(lsk_planner *pl, int vec, int space, double *global, void *stream);
int lsk_planner_vector_from_async(lsk_planner *pl, int vec, int space, const double *global, void *stream);

/* ---- solvers (src/CGSolver.hpp, src/BiCGStabSolver.hpp, src/GMRESSolver.hpp) -------------------------------------- */
enum lsk_solver_kind { LSK_SOLVER_CG = 1, LSK_SOLVER_BICGSTAB = 2, LSK_SOLVER_GMRES = 3 }; /* BenchmarkStencil -solver */
/* fused = 0: the reference's call sequence, one launch per planner call; non-zero: the same arithmetic in the fewest
 * passes over memory (CG 3 launches per step, BiCGStab 5) */
int lsk_solver_create(lsk_planner *pl, int kind, int restart, int fused, lsk_solver **out);
int lsk_solver_destroy(lsk_solver *s);
int lsk_solver_step(lsk_solver *s);
/* start a new solve from the current RHS with SOL taken as 0: re-runs the constructor's initialisation (CG: P <- RHS,
 * R <- RHS, rr0; BiCGStab: R, R~ <- RHS, P, V <- 0, rho = alpha' = omega = 1/0/1; GMRES keeps no state: no-op) */
int lsk_solver_reset(lsk_solver *s);
/* options: LSK_OPT_GMRES_REAL_UPDATE (GMRES): 0 (default) = the reference's placeholder update SOL += 1 * v_j
 * (DummyTask, src/GMRESSolver.hpp:109-126); 1 = the finished algorithm: Givens least squares on the Hessenberg, SOL += V y
 * in one pass, and history `which` = 1 then holds || b - A x || after each cycle.  Set outside any trace. */
enum lsk_solver_option { LSK_OPT_GMRES_REAL_UPDATE = 1 };
int lsk_solver_set_option(lsk_solver *s, int option, int value);
/* the first n entries of history `which` (as lsk_solver_history), copied with cudaMemcpyAsync(cudaMemcpyDefault) on
 * `stream` (NULL = the runtime's stream) without synchronising: dst may be device or pinned host memory.  For callers
 * that pipeline solves and read the histories once at the end (bench.py's end-to-end loop). */
int lsk_solver_history_copy_async(lsk_solver *s, int which, double *dst, int64_t n, void *stream);
/* which: CG 0 = residual_norm_squared; BiCGStab 0 = rho, 1 = alpha, 2 = omega; GMRES 0 = the
 * (restart+1) x restart inner_products table, row-major, 1 = residual norms (real update only).  Copies up to `cap` values (oldest first),
 * *n = number available.  Synchronises. */
int lsk_solver_history(lsk_solver *s, int which, double *out, int64_t cap, int64_t *n);

/* ---- utilities: Matrix Market files ----------------------------------------------------------------------------------
 * The reference names "reading in matrices from file formats like MATLAB or Matrix Market and storing them in standard
 * formats (e.g. COO, CSR, ELL, ...)" as the first of its planned utilities (README.md:90-99) and ships no reader.
 * Host-only (no CUDA).  Supported: `matrix coordinate {real|integer|pattern} {general|symmetric|skew-symmetric}`;
 * indices become 0-based, symmetric storage is expanded (the mirrored entry follows its original), pattern entries get
 * the value 1.  Dense (`array`) and complex files are refused.  Errors: status LSK_E_INVALID + lsk_mm_last_error(). */
enum lsk_mm_field { LSK_MM_REAL = 0, LSK_MM_INTEGER = 1, LSK_MM_PATTERN = 2 };
enum lsk_mm_symmetry { LSK_MM_GENERAL = 0, LSK_MM_SYMMETRIC = 1, LSK_MM_SKEW_SYMMETRIC = 2 };
typedef struct {
    int64_t rows, cols, entries;   /* entries = lines stored in the file; the expanded matrix holds at most 2 * entries */
    int field, symmetry;
} lsk_mm_info;
const char *lsk_mm_last_error(void);
int lsk_mm_read_info(const char *path, lsk_mm_info *out);
/* fills entry / row / col (COOMatrix fields, what lsk_coo_create uploads) with the EXPANDED matrix, in file order */
int lsk_mm_read_coo_f64(const char *path, int64_t capacity, double *entry, int64_t *row, int64_t *col, int64_t *nnz_out);
int lsk_mm_write_coo_f64(const char *path, int64_t rows, int64_t cols, int64_t nnz, const double *entry, const int64_t *row,
                         const int64_t *col);
/* COO in any order -> the CSRMatrix fields (what lsk_csr_create uploads): entries grouped by row, a row's entries in input
 * order; rowptr[r] = INCLUSIVE {first k, last k}, empty rows lo > hi (src/CSRMatrix.hpp field layout) */
int lsk_coo_to_csr_f64(int64_t rows, int64_t nnz, const double *entry, const int64_t *row, const int64_t *col, double *entry_out,
                       int64_t *col_out, lsk_rect *rowptr_out);

#ifdef __cplusplus
}
#endif
#endif /* LSK_SOLVERS_H */
