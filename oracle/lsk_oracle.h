/*
 * lsk_oracle.h -- CPU ORACLE for the Krylov inner loop of dzhang314/LegionSolvers.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may build, load
 * or call it -- always as the checker or the reported CPU baseline, never as the thing shipped.
 *
 * Every function is a plain-C restatement of one reference routine; the reference file:line it
 * follows is cited at each definition in lsk_oracle.c (paths relative to the reference root).
 *
 * Pinning status (see DESIGN.md "Oracle"):
 *   pinned by the reference's own golden vectors (tests/test_oracle_golden.py):
 *     - CG residual^2 history, 1-D Laplacian n=100, P=4          (test_all.py:130-133)
 *     - range / kernel / ghost partitions, n=20, P=4, COO == CSR  (test_all.py:19-127)
 *     - BLAS-1 chain -> 0                                         (Test02VectorOperations.cpp:128-137)
 *     - scalar chain == 1.0                                       (Test01ScalarOperations.cpp:17-32)
 *   pinned by outputs of the importable Python reference scripts/krylov.py (tests/golden/).
 *   PARITY UNPINNED (Legion/Realm behaviour, third party, absent here): the equal-partition
 *   split for N not divisible by P, and the order in which per-piece dot futures are summed.
 */
#ifndef LSK_ORACLE_H
#define LSK_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Legion::Rect<1, long long>: inclusive bounds, 16 bytes. */
typedef struct {
    int64_t lo, hi;
} orc_rect;

/* ---- thread control for the CPU-baseline legs (1 piece per thread, like 1 Legion CPU proc per piece) */
void orc_set_threads(int n);
int orc_get_threads(void);

/* ---- scalar futures ---------------------------------------------------------------------- */
double orc_get_alpha(int n, const double *f);
double orc_scalar_neg(double x);
double orc_scalar_add(double x, double y);
double orc_scalar_sub(double x, double y);
double orc_scalar_mul(double x, double y);
double orc_scalar_div(double x, double y);
double orc_scalar_sqrt(double x);
double orc_scalar_rsqrt(double x);
double orc_scalar_dummy(void);

float orc_get_alpha_f32(int n, const float *f);

/* ---- BLAS-1 leaf tasks (one piece) ---------------------------------------------------------- */
void orc_scal(int64_t n, double alpha, double *x);
void orc_axpy(int64_t n, double alpha, const double *x, double *y);
void orc_xpay(int64_t n, double alpha, const double *x, double *y);
double orc_dot(int64_t n, const double *v, const double *w);
/* test-only: 0 = the reference's sequential order (default), 1 = pairwise tree (sensitivity measurements) */
void orc_set_dot_order(int order);

void orc_scal_f32(int64_t n, float alpha, float *x);
void orc_axpy_f32(int64_t n, float alpha, const float *x, float *y);
void orc_xpay_f32(int64_t n, float alpha, const float *x, float *y);
float orc_dot_f32(int64_t n, const float *v, const float *w);

/* ---- mat-vec leaf tasks (one piece).  All index arguments are GLOBAL indices; `entry`, `col`,
 *      `row`, `rowptr`, `x` and `y` point at global index 0 of their arrays. ------------------- */
void orc_csr_matvec_literal(int64_t k_lo, int64_t k_hi, int64_t r_lo, int64_t r_hi, int64_t in_lo,
                            int64_t in_hi, const double *entry, const int64_t *col,
                            const orc_rect *rowptr, const double *x, double *y);
int orc_csr_matvec(int64_t k_lo, int64_t k_hi, int64_t r_lo, int64_t r_hi, int64_t in_lo,
                   int64_t in_hi, const double *entry, const int64_t *col, const orc_rect *rowptr,
                   const double *x, double *y);
void orc_coo_matvec(int64_t k_lo, int64_t k_hi, int64_t r_lo, int64_t r_hi, int64_t in_lo,
                    int64_t in_hi, const double *entry, const int64_t *row, const int64_t *col,
                    const double *x, double *y);

/* transposed products y[col_k] += entry_k * x[row_k] (no reference body exists: see lsk_oracle.c) */
void orc_coo_rmatvec(int64_t k_lo, int64_t k_hi, int64_t r_lo, int64_t r_hi, int64_t out_lo, int64_t out_hi,
                     const double *entry, const int64_t *row, const int64_t *col, const double *x, double *y);
void orc_csr_rmatvec(int64_t r_lo, int64_t r_hi, int64_t out_lo, int64_t out_hi, const double *entry,
                     const int64_t *col, const orc_rect *rowptr, const double *x, double *y);

/* ---- problem generators ---------------------------------------------------------------------- */
void orc_laplacian_1d_coo(int64_t k_lo, int64_t k_hi, double *entry, int64_t *row, int64_t *col);
void orc_laplacian_1d_csr(int64_t k_lo, int64_t k_hi, double *entry, int64_t *col);
void orc_laplacian_1d_rowptr(int64_t n, int64_t r_lo, int64_t r_hi, orc_rect *rowptr);
int64_t orc_laplacian_2d_kernel_size(int64_t height, int64_t width);

/* offsets: noff x dim int64 (row-major), values: noff doubles.  order: 0 row-major, 1 column-major */
int64_t orc_stencil_size(int dim, const int64_t *lo, const int64_t *hi, int noff,
                         const int64_t *offsets);
void orc_sort_stencil(int dim, int noff, int64_t *offsets, double *values, int order);
void orc_fill_linearized_csr_stencil(int dim, const int64_t *lo, const int64_t *hi, int noff,
                                     const int64_t *offsets, const double *values, int order,
                                     int64_t k_lo, int64_t k_hi, int64_t r_lo, int64_t r_hi,
                                     double *entry, int64_t *col, orc_rect *rowptr);
void orc_fill_linearized_coo_stencil(int dim, const int64_t *lo, const int64_t *hi, int noff,
                                     const int64_t *offsets, const double *values, int order,
                                     int64_t k_lo, int64_t k_hi, double *entry, int64_t *row,
                                     int64_t *col);

/* ---- partitions (third-party Legion/Realm operations, restated) ------------------------------ */
void orc_equal_partition(int64_t n, int pieces, int64_t *lo, int64_t *hi);
/* flags[] arrays are bytes over the parent space; 1 = index belongs to the piece */
void orc_image_range(const orc_rect *rowptr, int64_t r_lo, int64_t r_hi, int64_t nnz,
                     uint8_t *kernel_flags);
void orc_image(const int64_t *field, const uint8_t *kernel_flags, int64_t nnz, int64_t n,
               uint8_t *out_flags);
void orc_preimage(const int64_t *field, int64_t nnz, int64_t lo, int64_t hi,
                  uint8_t *kernel_flags);
void orc_preimage_range(const orc_rect *rowptr, int64_t n_rows, const uint8_t *kernel_flags,
                        uint8_t *range_flags);
int orc_shard(int64_t point, int64_t volume, int64_t total_shards);

/* ---- planner + solvers ------------------------------------------------------------------------ */
typedef struct orc_planner orc_planner;

orc_planner *orc_planner_create(int nspaces, const int64_t *n, const int *pieces);
void orc_planner_destroy(orc_planner *pl);
/* matrix arrays are borrowed; row == NULL means CSR (rowptr required), else COO */
int orc_planner_add_matrix(orc_planner *pl, int domain_idx, int range_idx, int64_t nnz,
                           const double *entry, const int64_t *col, const orc_rect *rowptr,
                           const int64_t *row);
void orc_planner_allocate_workspace(orc_planner *pl, int nvec);
double *orc_planner_vector(orc_planner *pl, int vec_idx, int space);
void orc_planner_piece_bounds(orc_planner *pl, int space, int piece, int64_t *lo, int64_t *hi);
void orc_planner_kernel_bounds(orc_planner *pl, int matrix, int piece, int64_t *lo, int64_t *hi);
void orc_planner_ghost_bounds(orc_planner *pl, int matrix, int piece, int64_t *lo, int64_t *hi);
void orc_planner_use_literal_csr(orc_planner *pl, int flag);

void orc_planner_fill(orc_planner *pl, int vec, double value);
void orc_planner_copy(orc_planner *pl, int dst, int src);
void orc_planner_scal(orc_planner *pl, int dst, int nterms, const double *terms);
void orc_planner_axpy(orc_planner *pl, int dst, int nterms, const double *terms, int src);
void orc_planner_xpay(orc_planner *pl, int dst, int nterms, const double *terms, int src);
double orc_planner_dot(orc_planner *pl, int v, int w);
void orc_planner_matvec(orc_planner *pl, int dst, int src);

typedef struct orc_cg orc_cg;
orc_cg *orc_cg_create(orc_planner *pl);
void orc_cg_step(orc_cg *s);
int64_t orc_cg_history(orc_cg *s, double *out, int64_t cap);
void orc_cg_destroy(orc_cg *s);

typedef struct orc_bicgstab orc_bicgstab;
orc_bicgstab *orc_bicgstab_create(orc_planner *pl);
void orc_bicgstab_step(orc_bicgstab *s);
/* which: 0 rho, 1 alpha, 2 omega */
int64_t orc_bicgstab_history(orc_bicgstab *s, int which, double *out, int64_t cap);
void orc_bicgstab_destroy(orc_bicgstab *s);

typedef struct orc_gmres orc_gmres;
orc_gmres *orc_gmres_create(orc_planner *pl, int restart);
void orc_gmres_step(orc_gmres *s);
/* (restart+1) x restart row-major "inner_products" table of the last cycle */
void orc_gmres_hessenberg(orc_gmres *s, double *out);
void orc_gmres_destroy(orc_gmres *s);

#ifdef __cplusplus
}
#endif
#endif /* LSK_ORACLE_H */
