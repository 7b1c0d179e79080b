/*
 * lsk_oracle.c -- CPU ORACLE (test infrastructure; see lsk_oracle.h for the rules of use).
 *
 * Plain-C restatement of the reference's CPU task variants and of the Legion operations the hot
 * path relies on.  Citations are file:line under the reference root (dzhang314/LegionSolvers).
 * Compile WITHOUT -ffast-math and WITHOUT fp contraction (-ffp-contract=off): the reference's
 * CPU bodies round every product before adding, except where they call std::fma explicitly.
 */
#include "lsk_oracle.h"

#include <assert.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

static int g_threads = 1;

void orc_set_threads(int n) { g_threads = n < 1 ? 1 : n; }
int orc_get_threads(void) { return g_threads; }

/* ============================================================================================
 * Scalar futures
 * ============================================================================================ */

/* src/LegionUtilities.cpp:72-97 -- 0 futures -> 1; 1 -> f0; 2 -> f0/f1; 3 -> (f0*f1)/f2;
 * 4 -> (f0*f1)/(f2*f3), in exactly that association. */
double orc_get_alpha(int n, const double *f) {
    switch (n) {
    case 0: return 1.0;
    case 1: return f[0];
    case 2: return f[0] / f[1];
    case 3: return (f[0] * f[1]) / f[2];
    case 4: return (f[0] * f[1]) / (f[2] * f[3]);
    default: assert(0); return 0.0;
    }
}

float orc_get_alpha_f32(int n, const float *f) {
    switch (n) {
    case 0: return 1.0f;
    case 1: return f[0];
    case 2: return f[0] / f[1];
    case 3: return (f[0] * f[1]) / f[2];
    case 4: return (f[0] * f[1]) / (f[2] * f[3]);
    default: assert(0); return 0.0f;
    }
}

/* src/UtilityTasks.cpp:34-99 (Negate/Add/Subtract/Multiply/Divide/Sqrt/RSqrt/Dummy). */
double orc_scalar_neg(double x) { return -x; }
double orc_scalar_add(double x, double y) { return x + y; }
double orc_scalar_sub(double x, double y) { return x - y; }
double orc_scalar_mul(double x, double y) { return x * y; }
double orc_scalar_div(double x, double y) { return x / y; }
double orc_scalar_sqrt(double x) { return sqrt(x); }
double orc_scalar_rsqrt(double x) { return 1.0 / sqrt(x); }
double orc_scalar_dummy(void) { return 1.0; }

/* ============================================================================================
 * BLAS-1 leaf tasks -- src/LinearAlgebraTasks.cpp
 * ============================================================================================ */

/* ScalTask::task_body, src/LinearAlgebraTasks.cpp:15-44: x = alpha * x */
void orc_scal(int64_t n, double alpha, double *x) {
    for (int64_t i = 0; i < n; ++i) x[i] = alpha * x[i];
}

/* AxpyTask::task_body, src/LinearAlgebraTasks.cpp:47-88: y = fma(alpha, x, y) */
void orc_axpy(int64_t n, double alpha, const double *x, double *y) {
    for (int64_t i = 0; i < n; ++i) y[i] = fma(alpha, x[i], y[i]);
}

/* XpayTask::task_body, src/LinearAlgebraTasks.cpp:91-132: y = fma(alpha, y, x) */
void orc_xpay(int64_t n, double alpha, const double *x, double *y) {
    for (int64_t i = 0; i < n; ++i) y[i] = fma(alpha, y[i], x[i]);
}

/* DotTask::task_body, src/LinearAlgebraTasks.cpp:135-175: sequential result += v*w */
/* Test-only switch (orc_set_dot_order): 0 = the reference's sequential sum below (default); 1 = a pairwise tree over
 * blocks of 256 sequential products -- ANOTHER valid evaluation order of the same dot product.  The tests use the
 * spread between the two to measure how strongly a Lanczos-type recurrence (BiCGStab, GMRES on symmetric matrices)
 * amplifies a change of summation order, and bound the GPU's deviation by that measured sensitivity. */
static int g_dot_order = 0;
void orc_set_dot_order(int order) { g_dot_order = order; }

static double dot_pairwise(int64_t n, const double *v, const double *w) {
    if (n <= 256) {
        double r = 0.0;
        for (int64_t i = 0; i < n; ++i) r += v[i] * w[i];
        return r;
    }
    const int64_t h = n / 2;
    return dot_pairwise(h, v, w) + dot_pairwise(n - h, v + h, w + h);
}

double orc_dot(int64_t n, const double *v, const double *w) {
    if (g_dot_order == 1) return dot_pairwise(n, v, w);
    double result = 0.0;
    for (int64_t i = 0; i < n; ++i) result += v[i] * w[i];
    return result;
}

void orc_scal_f32(int64_t n, float alpha, float *x) {
    for (int64_t i = 0; i < n; ++i) x[i] = alpha * x[i];
}
void orc_axpy_f32(int64_t n, float alpha, const float *x, float *y) {
    for (int64_t i = 0; i < n; ++i) y[i] = fmaf(alpha, x[i], y[i]);
}
void orc_xpay_f32(int64_t n, float alpha, const float *x, float *y) {
    for (int64_t i = 0; i < n; ++i) y[i] = fmaf(alpha, y[i], x[i]);
}
float orc_dot_f32(int64_t n, const float *v, const float *w) {
    float result = 0.0f;
    for (int64_t i = 0; i < n; ++i) result += v[i] * w[i];
    return result;
}

/* ============================================================================================
 * Mat-vec leaf tasks
 * ============================================================================================ */

/* CSRMatvecTask::task_body AS WRITTEN, src/CSRMatrixTasks.cpp:73-91: for every non-zero k of the
 * kernel piece, scan EVERY row of the rowptr piece for the rect that contains k (the last match
 * wins), then y[row] += entry[k] * x[col[k]] guarded by the input/output domains.
 * O(nnz_piece * rows_piece): usable only at golden-vector sizes. */
void orc_csr_matvec_literal(int64_t k_lo, int64_t k_hi, int64_t r_lo, int64_t r_hi, int64_t in_lo,
                            int64_t in_hi, const double *entry, const int64_t *col,
                            const orc_rect *rowptr, const double *x, double *y) {
    for (int64_t k = k_lo; k <= k_hi; ++k) {
        int64_t row = -1;
        for (int64_t r = r_lo; r <= r_hi; ++r) {
            if (rowptr[r].lo <= k && k <= rowptr[r].hi) row = r;
        }
        assert(row >= 0);
        const int64_t c = col[k];
        if (in_lo <= c && c <= in_hi && r_lo <= row && row <= r_hi) {
            y[row] += entry[k] * x[c];
        }
    }
}

/* The same k-ascending accumulation in linear time ("reference-equivalent"):
 * a row cursor advances with k, which is valid whenever the rects are ordered and disjoint (true
 * of every generator in the reference).  Visits the non-zeros in the same order and performs the
 * same rounded multiply-then-add, so the result is bit-identical to the literal scan.
 * Returns 0 on success, -1 if some k of the piece is covered by no row of the piece. */
int orc_csr_matvec(int64_t k_lo, int64_t k_hi, int64_t r_lo, int64_t r_hi, int64_t in_lo,
                   int64_t in_hi, const double *entry, const int64_t *col, const orc_rect *rowptr,
                   const double *x, double *y) {
    int64_t r = r_lo;
    for (int64_t k = k_lo; k <= k_hi; ++k) {
        while (r <= r_hi && rowptr[r].hi < k) ++r;
        if (r > r_hi || rowptr[r].lo > k) return -1;
        const int64_t c = col[k];
        if (in_lo <= c && c <= in_hi) y[r] += entry[k] * x[c];
    }
    return 0;
}

/* COOMatvecTask::task_body, src/COOMatrixTasks.cpp:66-74 */
void orc_coo_matvec(int64_t k_lo, int64_t k_hi, int64_t r_lo, int64_t r_hi, int64_t in_lo,
                    int64_t in_hi, const double *entry, const int64_t *row, const int64_t *col,
                    const double *x, double *y) {
    for (int64_t k = k_lo; k <= k_hi; ++k) {
        const int64_t r = row[k], c = col[k];
        if (in_lo <= c && c <= in_hi && r_lo <= r && r <= r_hi) y[r] += entry[k] * x[c];
    }
}

/* CSRRmatvecTask / COORmatvecTask: the reference reserves the tasks (src/TaskIDs.hpp:40-45) but their bodies are
 * `assert(false)` (src/CSRMatrixTasks.cpp:94-100, src/COOMatrixTasks.cpp:77-83) -- there is no reference arithmetic to
 * restate.  The transposed products are defined by analogy with the forward bodies above: the same loop over the
 * piece's stored non-zeros in ascending k, the same guards, product rounded then added, with the roles of row and
 * column exchanged:  y[col_k] += entry_k * x[row_k].  PARITY UNPINNED by any reference vector; pinned in the tests by
 * an independent check against scipy's A^T x. */
void orc_coo_rmatvec(int64_t k_lo, int64_t k_hi, int64_t r_lo, int64_t r_hi, int64_t out_lo, int64_t out_hi,
                     const double *entry, const int64_t *row, const int64_t *col, const double *x, double *y) {
    for (int64_t k = k_lo; k <= k_hi; ++k) {
        const int64_t r = row[k], c = col[k];
        if (out_lo <= c && c <= out_hi && r_lo <= r && r <= r_hi) y[c] += entry[k] * x[r];
    }
}

void orc_csr_rmatvec(int64_t r_lo, int64_t r_hi, int64_t out_lo, int64_t out_hi, const double *entry,
                     const int64_t *col, const orc_rect *rowptr, const double *x, double *y) {
    for (int64_t r = r_lo; r <= r_hi; ++r)
        for (int64_t k = rowptr[r].lo; k <= rowptr[r].hi; ++k) {
            const int64_t c = col[k];
            if (out_lo <= c && c <= out_hi) y[c] += entry[k] * x[r];
        }
}

/* ============================================================================================
 * Problem generators
 * ============================================================================================ */

/* FillCOONegativeLaplacianTask (1-D), src/ExampleSystems.cpp:311-321 */
void orc_laplacian_1d_coo(int64_t k_lo, int64_t k_hi, double *entry, int64_t *row, int64_t *col) {
    for (int64_t k = k_lo; k <= k_hi; ++k) {
        row[k] = (k + 1) / 3;
        col[k] = k - 2 * ((k + 1) / 3);
        entry[k] = (k % 3) ? -1.0 : +2.0;
    }
}

/* FillCSRNegativeLaplacianTask (1-D), src/ExampleSystems.cpp:382-389 */
void orc_laplacian_1d_csr(int64_t k_lo, int64_t k_hi, double *entry, int64_t *col) {
    for (int64_t k = k_lo; k <= k_hi; ++k) {
        col[k] = k - 2 * ((k + 1) / 3);
        entry[k] = (k % 3) ? -1.0 : +2.0;
    }
}

/* FillCSRNegativeLaplacianRowptrTask (1-D), src/ExampleSystems.cpp:448-466 */
void orc_laplacian_1d_rowptr(int64_t n, int64_t r_lo, int64_t r_hi, orc_rect *rowptr) {
    for (int64_t k = r_lo; k <= r_hi; ++k) {
        if (k == 0) {
            rowptr[k].lo = 0;
            rowptr[k].hi = 1;
        } else if (k == n - 1) {
            rowptr[k].lo = 3 * n - 4;
            rowptr[k].hi = 3 * n - 3;
        } else {
            rowptr[k].lo = 3 * k - 1;
            rowptr[k].hi = 3 * k + 1;
        }
    }
}

/* laplacian_2d_kernel_size, src/ExampleSystems.hpp:33-42 */
int64_t orc_laplacian_2d_kernel_size(int64_t height, int64_t width) {
    return 4 * 2 + (height - 2) * 2 * 3 + (width - 2) * 2 * 3 + (height - 2) * (width - 2) * 4 +
           width * height;
}

#define ORC_MAX_DIM 3

/* calculate_stencil_size, src/StencilGenerator.hpp:270-323: exact nnz by induction on DIM.  The
 * reference fast-forwards over the run of interior slices (all have the same sub-stencil); here
 * every slice is simply evaluated -- same sum. */
static int64_t stencil_size_rec(int dim, const int64_t *lo, const int64_t *hi, int noff,
                                const int64_t *offsets, int stride) {
    if (noff == 0) return 0;
    int64_t result = 0;
    if (dim == 1) {
        const int64_t length = hi[0] - lo[0] + 1;
        for (int j = 0; j < noff; ++j) {
            const int64_t d = llabs(offsets[(size_t) j * stride]);
            if (d <= length) result += length - d;
        }
        return result;
    }
    int64_t *sub = (int64_t *) malloc(sizeof(int64_t) * (size_t) noff * stride);
    int64_t memo_full = -1;
    for (int64_t i = lo[0]; i <= hi[0]; ++i) {
        int nsub = 0;
        for (int j = 0; j < noff; ++j) {
            const int64_t s = i + offsets[(size_t) j * stride];
            if (lo[0] <= s && s <= hi[0]) {
                memcpy(sub + (size_t) nsub * stride, offsets + (size_t) j * stride + 1,
                       sizeof(int64_t) * (size_t) (dim - 1));
                ++nsub;
            }
        }
        if (nsub == noff) {
            if (memo_full < 0) memo_full = stencil_size_rec(dim - 1, lo + 1, hi + 1, nsub, sub, stride);
            result += memo_full;
        } else {
            result += stencil_size_rec(dim - 1, lo + 1, hi + 1, nsub, sub, stride);
        }
    }
    free(sub);
    return result;
}

int64_t orc_stencil_size(int dim, const int64_t *lo, const int64_t *hi, int noff,
                         const int64_t *offsets) {
    assert(dim >= 1 && dim <= ORC_MAX_DIM);
    /* repack with a fixed stride so the recursion can slice tails in place */
    int64_t *packed = (int64_t *) malloc(sizeof(int64_t) * (size_t) (noff > 0 ? noff : 1) * ORC_MAX_DIM);
    for (int j = 0; j < noff; ++j)
        for (int d = 0; d < dim; ++d) packed[(size_t) j * ORC_MAX_DIM + d] = offsets[(size_t) j * dim + d];
    const int64_t r = stencil_size_rec(dim, lo, hi, noff, packed, ORC_MAX_DIM);
    free(packed);
    return r;
}

/* compare_row_major / compare_column_major, src/StencilGenerator.hpp:172-194 */
static int point_less(int dim, const int64_t *p, const int64_t *q, int order) {
    if (order == 0) {
        for (int i = 0; i < dim; ++i) {
            if (p[i] < q[i]) return 1;
            if (p[i] > q[i]) return 0;
        }
    } else {
        for (int i = dim - 1; i >= 0; --i) {
            if (p[i] < q[i]) return 1;
            if (p[i] > q[i]) return 0;
        }
    }
    return 0;
}

/* The std::sort of the (offset, entry) pairs at the top of every fill task,
 * src/StencilGenerator.cpp:408-433: by point in the chosen order, ties by entry. */
void orc_sort_stencil(int dim, int noff, int64_t *offsets, double *values, int order) {
    for (int i = 1; i < noff; ++i) { /* insertion sort: the comparator is a strict total order */
        int64_t p[ORC_MAX_DIM];
        memcpy(p, offsets + (size_t) i * dim, sizeof(int64_t) * (size_t) dim);
        const double v = values[i];
        int j = i - 1;
        while (j >= 0) {
            const int64_t *q = offsets + (size_t) j * dim;
            const int less = point_less(dim, p, q, order)
                                 ? 1
                                 : (point_less(dim, q, p, order) ? 0 : (v < values[j]));
            if (!less) break;
            memcpy(offsets + (size_t) (j + 1) * dim, q, sizeof(int64_t) * (size_t) dim);
            values[j + 1] = values[j];
            --j;
        }
        memcpy(offsets + (size_t) (j + 1) * dim, p, sizeof(int64_t) * (size_t) dim);
        values[j + 1] = v;
    }
}

/* linearize_row_major / linearize_column_major, src/StencilGenerator.hpp:230-259 */
static int64_t linearize(int dim, const int64_t *p, const int64_t *lo, const int64_t *hi, int order) {
    int64_t result = 0, acc = 1;
    if (order == 0) {
        for (int i = dim - 1; i >= 0; --i) {
            result += acc * (p[i] - lo[i]);
            acc *= hi[i] - lo[i] + 1;
        }
    } else {
        for (int i = 0; i < dim; ++i) {
            result += acc * (p[i] - lo[i]);
            acc *= hi[i] - lo[i] + 1;
        }
    }
    return result;
}

/* increment_row_major / increment_column_major, src/StencilGenerator.hpp:197-227 */
static int increment(int dim, int64_t *p, const int64_t *lo, const int64_t *hi, int order) {
    if (order == 0) {
        for (int i = dim - 1; i >= 0; --i) {
            if (p[i] >= hi[i]) p[i] = lo[i];
            else { ++p[i]; return 1; }
        }
    } else {
        for (int i = 0; i < dim; ++i) {
            if (p[i] >= hi[i]) p[i] = lo[i];
            else { ++p[i]; return 1; }
        }
    }
    return 0;
}

static int in_bounds(int dim, const int64_t *p, const int64_t *lo, const int64_t *hi) {
    for (int i = 0; i < dim; ++i)
        if (p[i] < lo[i] || p[i] > hi[i]) return 0;
    return 1;
}

/* FillLinearizedCSRStencilTask::task_body, src/StencilGenerator.cpp:380-543.  Walks the whole grid
 * in the chosen order, emitting one non-zero per in-bounds (sorted) offset; writes only the k in
 * [k_lo,k_hi] and the rows in [r_lo,r_hi].  rowptr[row] = inclusive {first k, last k}.  The
 * reference's bulk-slice fast-forward (:472-489) only skips grid slices that lie wholly before
 * the piece; the values written are the same, so it is omitted here. */
void orc_fill_linearized_csr_stencil(int dim, const int64_t *lo, const int64_t *hi, int noff,
                                     const int64_t *offsets_in, const double *values_in, int order,
                                     int64_t k_lo, int64_t k_hi, int64_t r_lo, int64_t r_hi,
                                     double *entry, int64_t *col, orc_rect *rowptr) {
    int64_t *offsets = (int64_t *) malloc(sizeof(int64_t) * (size_t) noff * dim);
    double *values = (double *) malloc(sizeof(double) * (size_t) noff);
    memcpy(offsets, offsets_in, sizeof(int64_t) * (size_t) noff * dim);
    memcpy(values, values_in, sizeof(double) * (size_t) noff);
    orc_sort_stencil(dim, noff, offsets, values, order);

    int64_t point[ORC_MAX_DIM], shifted[ORC_MAX_DIM];
    memcpy(point, lo, sizeof(int64_t) * (size_t) dim);
    int64_t k = 0;
    do {
        const int64_t point_lin = linearize(dim, point, lo, hi, order);
        const int64_t row_begin = k;
        for (int j = 0; j < noff; ++j) {
            for (int d = 0; d < dim; ++d) shifted[d] = point[d] + offsets[(size_t) j * dim + d];
            if (in_bounds(dim, shifted, lo, hi)) {
                if (k_lo <= k && k <= k_hi) {
                    col[k] = linearize(dim, shifted, lo, hi, order);
                    entry[k] = values[j];
                }
                ++k;
            }
        }
        if (r_lo <= point_lin && point_lin <= r_hi) {
            rowptr[point_lin].lo = row_begin;
            rowptr[point_lin].hi = k - 1;
        }
        if (order == 0 && k > k_hi && point_lin > r_hi) break; /* :524-527 early exit */
    } while (increment(dim, point, lo, hi, order));
    free(offsets);
    free(values);
}

/* FillLinearizedCOOStencilTask::task_body, src/StencilGenerator.cpp:160-243 */
void orc_fill_linearized_coo_stencil(int dim, const int64_t *lo, const int64_t *hi, int noff,
                                     const int64_t *offsets_in, const double *values_in, int order,
                                     int64_t k_lo, int64_t k_hi, double *entry, int64_t *row,
                                     int64_t *col) {
    int64_t *offsets = (int64_t *) malloc(sizeof(int64_t) * (size_t) noff * dim);
    double *values = (double *) malloc(sizeof(double) * (size_t) noff);
    memcpy(offsets, offsets_in, sizeof(int64_t) * (size_t) noff * dim);
    memcpy(values, values_in, sizeof(double) * (size_t) noff);
    orc_sort_stencil(dim, noff, offsets, values, order);

    int64_t point[ORC_MAX_DIM], shifted[ORC_MAX_DIM];
    memcpy(point, lo, sizeof(int64_t) * (size_t) dim);
    int64_t k = 0;
    do {
        for (int j = 0; j < noff; ++j) {
            for (int d = 0; d < dim; ++d) shifted[d] = point[d] + offsets[(size_t) j * dim + d];
            if (in_bounds(dim, shifted, lo, hi)) {
                if (k_lo <= k && k <= k_hi) {
                    row[k] = linearize(dim, point, lo, hi, order);
                    col[k] = linearize(dim, shifted, lo, hi, order);
                    entry[k] = values[j];
                }
                ++k;
            }
        }
        if (k > k_hi) break;
    } while (increment(dim, point, lo, hi, order));
    free(offsets);
    free(values);
}

/* ============================================================================================
 * Partitions (Legion/Realm, third party: legion-24.12.0 / master, build_legion.py:29-38)
 * ============================================================================================ */

/* Runtime::create_equal_partition on a dense 1-D space (call sites test/Test06CSRSolveCG.cpp:59-60,
 * src/StencilGenerator.hpp:582-588).  Realm splits a dense rect of `n` points into `pieces`
 * sub-rects with piece i = [floor(n*i/P), floor(n*(i+1)/P) - 1].  PINNED only for n % P == 0
 * (golden partition n=20, P=4); the non-divisible rule is PARITY UNPINNED. */
void orc_equal_partition(int64_t n, int pieces, int64_t *lo, int64_t *hi) {
    for (int i = 0; i < pieces; ++i) {
        lo[i] = (int64_t) (((__int128) n * i) / pieces);
        hi[i] = (int64_t) (((__int128) n * (i + 1)) / pieces) - 1;
    }
}

/* create_partition_by_image_range over the rowptr field (src/CSRMatrix.cpp:89-109):
 * kernel piece = union of rowptr[r] for r in the range piece. */
void orc_image_range(const orc_rect *rowptr, int64_t r_lo, int64_t r_hi, int64_t nnz,
                     uint8_t *kernel_flags) {
    memset(kernel_flags, 0, (size_t) nnz);
    for (int64_t r = r_lo; r <= r_hi; ++r)
        for (int64_t k = rowptr[r].lo; k <= rowptr[r].hi; ++k)
            if (0 <= k && k < nnz) kernel_flags[k] = 1;
}

/* create_partition_by_image over a point field (src/CSRMatrix.cpp:112-132, src/COOMatrix.cpp:98-141):
 * out piece = { field[k] : k in kernel piece } intersected with the parent space [0,n). */
void orc_image(const int64_t *field, const uint8_t *kernel_flags, int64_t nnz, int64_t n,
               uint8_t *out_flags) {
    memset(out_flags, 0, (size_t) n);
    for (int64_t k = 0; k < nnz; ++k)
        if (kernel_flags[k] && 0 <= field[k] && field[k] < n) out_flags[field[k]] = 1;
}

/* create_partition_by_preimage over a point field (src/COOMatrix.cpp:56-96, src/CSRMatrix.cpp:68-86):
 * kernel piece = { k : field[k] in [lo,hi] }. */
void orc_preimage(const int64_t *field, int64_t nnz, int64_t lo, int64_t hi, uint8_t *kernel_flags) {
    for (int64_t k = 0; k < nnz; ++k) kernel_flags[k] = (lo <= field[k] && field[k] <= hi) ? 1 : 0;
}

/* create_partition_by_preimage_range (src/CSRMatrix.cpp:135-155):
 * range piece = { r : rowptr[r] intersects the kernel piece }. */
void orc_preimage_range(const orc_rect *rowptr, int64_t n_rows, const uint8_t *kernel_flags,
                        uint8_t *range_flags) {
    for (int64_t r = 0; r < n_rows; ++r) {
        uint8_t hit = 0;
        for (int64_t k = rowptr[r].lo; k <= rowptr[r].hi && !hit; ++k) hit = kernel_flags[k];
        range_flags[r] = hit;
    }
}

/* BlockingShardingFunctor::shard, src/LegionSolversMapper.cpp:140-151 */
int orc_shard(int64_t point, int64_t volume, int64_t total_shards) {
    const int64_t per = (volume + total_shards - 1) / total_shards;
    return (int) (point / per);
}

/* ============================================================================================
 * SquarePlanner (src/SquarePlanner.hpp) over PartitionedVector ops (src/PartitionedVector.cpp)
 * ============================================================================================ */

typedef struct {
    int domain_idx, range_idx;
    int64_t nnz;
    const double *entry;
    const int64_t *col;
    const orc_rect *rowptr; /* CSR */
    const int64_t *row;     /* COO */
    int64_t *k_lo, *k_hi;   /* kernel partition, per range piece (bounding interval) */
    int64_t *g_lo, *g_hi;   /* ghost partition, per range piece (bounding interval)  */
} orc_block;

struct orc_planner {
    int nspaces;
    int64_t *n;
    int *pieces;
    int64_t **p_lo, **p_hi; /* canonical partitions */
    int nvec;               /* 2 + workspace */
    double ***vec;          /* vec[v][space] -> n[space] doubles */
    int nblocks;
    orc_block *blocks;
    int literal_csr;
};

orc_planner *orc_planner_create(int nspaces, const int64_t *n, const int *pieces) {
    orc_planner *pl = (orc_planner *) calloc(1, sizeof(*pl));
    pl->nspaces = nspaces;
    pl->n = (int64_t *) malloc(sizeof(int64_t) * (size_t) nspaces);
    pl->pieces = (int *) malloc(sizeof(int) * (size_t) nspaces);
    pl->p_lo = (int64_t **) malloc(sizeof(int64_t *) * (size_t) nspaces);
    pl->p_hi = (int64_t **) malloc(sizeof(int64_t *) * (size_t) nspaces);
    for (int s = 0; s < nspaces; ++s) {
        pl->n[s] = n[s];
        pl->pieces[s] = pieces[s];
        pl->p_lo[s] = (int64_t *) malloc(sizeof(int64_t) * (size_t) pieces[s]);
        pl->p_hi[s] = (int64_t *) malloc(sizeof(int64_t) * (size_t) pieces[s]);
        orc_equal_partition(n[s], pieces[s], pl->p_lo[s], pl->p_hi[s]);
    }
    /* vectors 0 (SOL) and 1 (RHS) always exist: add_sol_vector / add_rhs_vector, :99-151 */
    pl->nvec = 2;
    pl->vec = (double ***) malloc(sizeof(double **) * 2);
    for (int v = 0; v < 2; ++v) {
        pl->vec[v] = (double **) malloc(sizeof(double *) * (size_t) nspaces);
        for (int s = 0; s < nspaces; ++s) pl->vec[v][s] = (double *) calloc((size_t) n[s], sizeof(double));
    }
    return pl;
}

void orc_planner_destroy(orc_planner *pl) {
    if (!pl) return;
    for (int v = 0; v < pl->nvec; ++v) {
        for (int s = 0; s < pl->nspaces; ++s) free(pl->vec[v][s]);
        free(pl->vec[v]);
    }
    free(pl->vec);
    for (int b = 0; b < pl->nblocks; ++b) {
        free(pl->blocks[b].k_lo); free(pl->blocks[b].k_hi);
        free(pl->blocks[b].g_lo); free(pl->blocks[b].g_hi);
    }
    free(pl->blocks);
    for (int s = 0; s < pl->nspaces; ++s) { free(pl->p_lo[s]); free(pl->p_hi[s]); }
    free(pl->p_lo); free(pl->p_hi); free(pl->n); free(pl->pieces);
    free(pl);
}

void orc_planner_use_literal_csr(orc_planner *pl, int flag) { pl->literal_csr = flag; }

/* add_row_partitioned_matrix, src/SquarePlanner.hpp:209-235: kernel partition from the range
 * partition (CSR: image_range over rowptr, src/CSRMatrix.cpp:89-109; COO: preimage of row,
 * src/COOMatrix.cpp:77-96), then ghost partition = image of col over the kernel partition.
 * The oracle keeps the bounding interval of each piece (what a physical instance covers) and
 * requires the kernel pieces to be contiguous, which holds for row-sorted matrices.
 * Returns the block index, or -1 if a kernel piece is not a contiguous interval. */
int orc_planner_add_matrix(orc_planner *pl, int domain_idx, int range_idx, int64_t nnz,
                           const double *entry, const int64_t *col, const orc_rect *rowptr,
                           const int64_t *row) {
    pl->blocks = (orc_block *) realloc(pl->blocks, sizeof(orc_block) * (size_t) (pl->nblocks + 1));
    orc_block *b = &pl->blocks[pl->nblocks];
    memset(b, 0, sizeof(*b));
    b->domain_idx = domain_idx; b->range_idx = range_idx; b->nnz = nnz;
    b->entry = entry; b->col = col; b->rowptr = rowptr; b->row = row;
    const int P = pl->pieces[range_idx];
    b->k_lo = (int64_t *) malloc(sizeof(int64_t) * (size_t) P);
    b->k_hi = (int64_t *) malloc(sizeof(int64_t) * (size_t) P);
    b->g_lo = (int64_t *) malloc(sizeof(int64_t) * (size_t) P);
    b->g_hi = (int64_t *) malloc(sizeof(int64_t) * (size_t) P);
    for (int c = 0; c < P; ++c) {
        const int64_t r_lo = pl->p_lo[range_idx][c], r_hi = pl->p_hi[range_idx][c];
        int64_t klo = INT64_MAX, khi = INT64_MIN, count = 0;
        if (row == NULL) {
            for (int64_t r = r_lo; r <= r_hi; ++r) {
                if (rowptr[r].hi < rowptr[r].lo) continue;
                if (rowptr[r].lo < klo) klo = rowptr[r].lo;
                if (rowptr[r].hi > khi) khi = rowptr[r].hi;
                count += rowptr[r].hi - rowptr[r].lo + 1;
            }
        } else {
            for (int64_t k = 0; k < nnz; ++k) {
                if (r_lo <= row[k] && row[k] <= r_hi) {
                    if (k < klo) klo = k;
                    if (k > khi) khi = k;
                    ++count;
                }
            }
        }
        if (count == 0) { klo = 0; khi = -1; }
        else if (khi - klo + 1 != count) return -1;
        int64_t glo = INT64_MAX, ghi = INT64_MIN;
        for (int64_t k = klo; k <= khi; ++k) {
            if (col[k] < glo) glo = col[k];
            if (col[k] > ghi) ghi = col[k];
        }
        if (count == 0) { glo = 0; ghi = -1; }
        b->k_lo[c] = klo; b->k_hi[c] = khi; b->g_lo[c] = glo; b->g_hi[c] = ghi;
    }
    return pl->nblocks++;
}

/* allocate_workspace, src/SquarePlanner.hpp:153-190 */
void orc_planner_allocate_workspace(orc_planner *pl, int nvec) {
    const int total = pl->nvec + nvec;
    pl->vec = (double ***) realloc(pl->vec, sizeof(double **) * (size_t) total);
    for (int v = pl->nvec; v < total; ++v) {
        pl->vec[v] = (double **) malloc(sizeof(double *) * (size_t) pl->nspaces);
        for (int s = 0; s < pl->nspaces; ++s)
            pl->vec[v][s] = (double *) calloc((size_t) pl->n[s], sizeof(double));
    }
    pl->nvec = total;
}

double *orc_planner_vector(orc_planner *pl, int vec_idx, int space) { return pl->vec[vec_idx][space]; }

void orc_planner_piece_bounds(orc_planner *pl, int space, int piece, int64_t *lo, int64_t *hi) {
    *lo = pl->p_lo[space][piece]; *hi = pl->p_hi[space][piece];
}
void orc_planner_kernel_bounds(orc_planner *pl, int m, int piece, int64_t *lo, int64_t *hi) {
    *lo = pl->blocks[m].k_lo[piece]; *hi = pl->blocks[m].k_hi[piece];
}
void orc_planner_ghost_bounds(orc_planner *pl, int m, int piece, int64_t *lo, int64_t *hi) {
    *lo = pl->blocks[m].g_lo[piece]; *hi = pl->blocks[m].g_hi[piece];
}

/* Flattened (space, piece) iteration: one "point task" per piece, run on g_threads host threads --
 * the stand-in for one Legion CPU processor per piece. */
typedef struct { int space, piece; } orc_sp;

static int planner_points(const orc_planner *pl, orc_sp **out) {
    int total = 0;
    for (int s = 0; s < pl->nspaces; ++s) total += pl->pieces[s];
    orc_sp *pts = (orc_sp *) malloc(sizeof(orc_sp) * (size_t) total);
    int i = 0;
    for (int s = 0; s < pl->nspaces; ++s)
        for (int c = 0; c < pl->pieces[s]; ++c) { pts[i].space = s; pts[i].piece = c; ++i; }
    *out = pts;
    return total;
}

#define ORC_FOR_PIECES(pl, ...)                                                                    \
    do {                                                                                           \
        orc_sp *pts_;                                                                              \
        const int npts_ = planner_points(pl, &pts_);                                               \
        _Pragma("omp parallel for schedule(static) num_threads(g_threads) if (g_threads > 1)")     \
        for (int i_ = 0; i_ < npts_; ++i_) {                                                       \
            const int s = pts_[i_].space, c = pts_[i_].piece;                                      \
            const int64_t lo = pl->p_lo[s][c], cnt = pl->p_hi[s][c] - lo + 1;                      \
            (void) cnt;                                                                            \
            __VA_ARGS__                                                                            \
        }                                                                                          \
        free(pts_);                                                                                \
    } while (0)

/* PartitionedVector::constant_fill (IndexFill), src/PartitionedVector.cpp:150-161 */
void orc_planner_fill(orc_planner *pl, int v, double value) {
    ORC_FOR_PIECES(pl, { double *x = pl->vec[v][s] + lo; for (int64_t i = 0; i < cnt; ++i) x[i] = value; });
}

/* PartitionedVector::operator= (IndexCopy), src/PartitionedVector.cpp:176-192; planner.copy :269-274 */
void orc_planner_copy(orc_planner *pl, int dst, int src) {
    ORC_FOR_PIECES(pl, { memcpy(pl->vec[dst][s] + lo, pl->vec[src][s] + lo, sizeof(double) * (size_t) cnt); });
}

/* planner.scal :276-281 -> PartitionedVector::scal :195-208 -> ScalTask */
void orc_planner_scal(orc_planner *pl, int dst, int nterms, const double *terms) {
    ORC_FOR_PIECES(pl, { orc_scal(cnt, orc_get_alpha(nterms, terms), pl->vec[dst][s] + lo); });
}

/* planner.axpy :283-322 -> PartitionedVector::axpy (1, 2 or 3 futures) :211-287 -> AxpyTask */
void orc_planner_axpy(orc_planner *pl, int dst, int nterms, const double *terms, int src) {
    ORC_FOR_PIECES(pl, { orc_axpy(cnt, orc_get_alpha(nterms, terms), pl->vec[src][s] + lo, pl->vec[dst][s] + lo); });
}

/* planner.xpay :324-346 -> PartitionedVector::xpay (1 or 2 futures) :290-335 -> XpayTask */
void orc_planner_xpay(orc_planner *pl, int dst, int nterms, const double *terms, int src) {
    ORC_FOR_PIECES(pl, { orc_xpay(cnt, orc_get_alpha(nterms, terms), pl->vec[src][s] + lo, pl->vec[dst][s] + lo); });
}

/* planner.dot :331-338: per space, PartitionedVector::dot (src/PartitionedVector.cpp:337-358) =
 * one DotTask per piece, future-map sum-reduction; spaces are then chained with AddScalarTask.
 * The order in which Legion folds the per-piece futures is PARITY UNPINNED; the oracle folds them
 * in colour order starting from the first piece's value. */
double orc_planner_dot(orc_planner *pl, int v, int w) {
    int total = 0;
    for (int s = 0; s < pl->nspaces; ++s) total += pl->pieces[s];
    double *partial = (double *) malloc(sizeof(double) * (size_t) total);
    int *base = (int *) malloc(sizeof(int) * (size_t) pl->nspaces);
    for (int s = 0, acc = 0; s < pl->nspaces; ++s) { base[s] = acc; acc += pl->pieces[s]; }
    ORC_FOR_PIECES(pl, { partial[base[s] + c] = orc_dot(cnt, pl->vec[v][s] + lo, pl->vec[w][s] + lo); });
    double result = 0.0;
    for (int s = 0; s < pl->nspaces; ++s) {
        double space_sum = partial[base[s]];
        for (int c = 1; c < pl->pieces[s]; ++c) space_sum += partial[base[s] + c];
        result = (s == 0) ? space_sum : result + space_sum;
    }
    free(partial);
    free(base);
    return result;
}

/* planner.matvec, src/SquarePlanner.hpp:340-357: zero_fill(dst), then one mat-vec index launch per
 * registered block accumulating into dst (READ_WRITE sum accessor, src/CSRMatrix.cpp:158-214,
 * src/COOMatrix.cpp:144-191). */
void orc_planner_matvec(orc_planner *pl, int dst, int src) {
    orc_planner_fill(pl, dst, 0.0);
    for (int m = 0; m < pl->nblocks; ++m) {
        const orc_block *b = &pl->blocks[m];
        const int rs = b->range_idx, ds = b->domain_idx;
        const int P = pl->pieces[rs];
        double *y = pl->vec[dst][rs];
        const double *x = pl->vec[src][ds];
        _Pragma("omp parallel for schedule(static) num_threads(g_threads) if (g_threads > 1)")
        for (int c = 0; c < P; ++c) {
            const int64_t r_lo = pl->p_lo[rs][c], r_hi = pl->p_hi[rs][c];
            if (b->k_hi[c] < b->k_lo[c]) continue;
            if (b->row != NULL) {
                orc_coo_matvec(b->k_lo[c], b->k_hi[c], r_lo, r_hi, b->g_lo[c], b->g_hi[c], b->entry,
                               b->row, b->col, x, y);
            } else if (pl->literal_csr) {
                orc_csr_matvec_literal(b->k_lo[c], b->k_hi[c], r_lo, r_hi, b->g_lo[c], b->g_hi[c],
                                       b->entry, b->col, b->rowptr, x, y);
            } else {
                const int rc = orc_csr_matvec(b->k_lo[c], b->k_hi[c], r_lo, r_hi, b->g_lo[c],
                                              b->g_hi[c], b->entry, b->col, b->rowptr, x, y);
                assert(rc == 0);
                (void) rc;
            }
        }
    }
}

/* ============================================================================================
 * Solvers
 * ============================================================================================ */

enum { V_SOL = 0, V_RHS = 1 };

/* ---- CGSolver, src/CGSolver.hpp --------------------------------------------------------------- */
struct orc_cg {
    orc_planner *pl;
    double *rr;
    int64_t len, cap;
};
enum { CG_P = 2, CG_Q = 3, CG_R = 4 };

static void cg_push(orc_cg *s, double v) {
    if (s->len == s->cap) {
        s->cap = s->cap ? 2 * s->cap : 64;
        s->rr = (double *) realloc(s->rr, sizeof(double) * (size_t) s->cap);
    }
    s->rr[s->len++] = v;
}

/* constructor :32-44: workspace(3); P <- RHS; R <- RHS (x0 assumed 0); rr[0] = R.R */
orc_cg *orc_cg_create(orc_planner *pl) {
    orc_cg *s = (orc_cg *) calloc(1, sizeof(*s));
    s->pl = pl;
    orc_planner_allocate_workspace(pl, 3);
    orc_planner_copy(pl, CG_P, V_RHS);
    orc_planner_copy(pl, CG_R, V_RHS);
    cg_push(s, orc_planner_dot(pl, CG_R, CG_R));
    return s;
}

/* step :46-55 */
void orc_cg_step(orc_cg *s) {
    orc_planner *pl = s->pl;
    orc_planner_matvec(pl, CG_Q, CG_P);
    const double p_norm = orc_planner_dot(pl, CG_P, CG_Q);
    const double rr_old = s->rr[s->len - 1];
    const double t2[2] = {rr_old, p_norm};
    orc_planner_axpy(pl, V_SOL, 2, t2, CG_P);
    const double t3[3] = {-1.0, rr_old, p_norm};
    orc_planner_axpy(pl, CG_R, 3, t3, CG_Q);
    const double rr_new = orc_planner_dot(pl, CG_R, CG_R);
    cg_push(s, rr_new);
    const double tb[2] = {rr_new, rr_old};
    orc_planner_xpay(pl, CG_P, 2, tb, CG_R);
}

int64_t orc_cg_history(orc_cg *s, double *out, int64_t cap) {
    const int64_t n = s->len < cap ? s->len : cap;
    if (out) memcpy(out, s->rr, sizeof(double) * (size_t) n);
    return s->len;
}

void orc_cg_destroy(orc_cg *s) {
    if (s) { free(s->rr); free(s); }
}

/* ---- BiCGStabSolver, src/BiCGStabSolver.hpp ----------------------------------------------------- */
struct orc_bicgstab {
    orc_planner *pl;
    double *h[3]; /* rho, alpha, omega */
    int64_t len[3], cap[3];
};
enum { BI_P = 2, BI_R = 3, BI_RT = 4, BI_U = 5, BI_V = 6 };

static void bi_push(orc_bicgstab *s, int w, double v) {
    if (s->len[w] == s->cap[w]) {
        s->cap[w] = s->cap[w] ? 2 * s->cap[w] : 64;
        s->h[w] = (double *) realloc(s->h[w], sizeof(double) * (size_t) s->cap[w]);
    }
    s->h[w][s->len[w]++] = v;
}

/* constructor :36-60 */
orc_bicgstab *orc_bicgstab_create(orc_planner *pl) {
    orc_bicgstab *s = (orc_bicgstab *) calloc(1, sizeof(*s));
    s->pl = pl;
    orc_planner_allocate_workspace(pl, 5);
    orc_planner_copy(pl, BI_R, V_RHS);
    orc_planner_copy(pl, BI_RT, V_RHS);
    bi_push(s, 0, 1.0);
    bi_push(s, 1, 0.0);
    bi_push(s, 2, 1.0);
    orc_planner_fill(pl, BI_P, 0.0);
    orc_planner_fill(pl, BI_V, 0.0);
    return s;
}

#define BI_BACK(s, w) ((s)->h[w][(s)->len[w] - 1])

/* step :62-82 */
void orc_bicgstab_step(orc_bicgstab *s) {
    orc_planner *pl = s->pl;
    const double rho_new = orc_planner_dot(pl, BI_R, BI_RT);
    const double beta = orc_scalar_mul(orc_scalar_div(rho_new, BI_BACK(s, 0)),
                                       orc_scalar_div(BI_BACK(s, 1), BI_BACK(s, 2)));
    bi_push(s, 0, rho_new);
    double t1[1];
    t1[0] = orc_scalar_neg(BI_BACK(s, 2));
    orc_planner_axpy(pl, BI_P, 1, t1, BI_V);
    t1[0] = beta;
    orc_planner_xpay(pl, BI_P, 1, t1, BI_R);
    orc_planner_matvec(pl, BI_V, BI_P);
    const double temp = orc_planner_dot(pl, BI_RT, BI_V);
    const double t3[3] = {-1.0, BI_BACK(s, 0), temp};
    orc_planner_axpy(pl, BI_R, 3, t3, BI_V);
    bi_push(s, 1, orc_scalar_div(BI_BACK(s, 0), temp));
    orc_planner_matvec(pl, BI_U, BI_R);
    const double r_anorm2 = orc_planner_dot(pl, BI_R, BI_U);
    const double u_norm2 = orc_planner_dot(pl, BI_U, BI_U);
    bi_push(s, 2, orc_scalar_div(r_anorm2, u_norm2));
    t1[0] = BI_BACK(s, 1);
    orc_planner_axpy(pl, V_SOL, 1, t1, BI_P);
    t1[0] = BI_BACK(s, 2);
    orc_planner_axpy(pl, V_SOL, 1, t1, BI_R);
    t1[0] = orc_scalar_neg(BI_BACK(s, 2));
    orc_planner_axpy(pl, BI_R, 1, t1, BI_U);
}

int64_t orc_bicgstab_history(orc_bicgstab *s, int which, double *out, int64_t cap) {
    const int64_t n = s->len[which] < cap ? s->len[which] : cap;
    if (out) memcpy(out, s->h[which], sizeof(double) * (size_t) n);
    return s->len[which];
}

void orc_bicgstab_destroy(orc_bicgstab *s) {
    if (s) { free(s->h[0]); free(s->h[1]); free(s->h[2]); free(s); }
}

/* ---- GMRESSolver, src/GMRESSolver.hpp ------------------------------------------------------------- */
struct orc_gmres {
    orc_planner *pl;
    int restart;
    double *ip; /* (restart+1) x restart "inner_products" */
};
#define KRYLOV(i) ((i) + 2)

/* constructor :32-77 */
orc_gmres *orc_gmres_create(orc_planner *pl, int restart) {
    orc_gmres *s = (orc_gmres *) calloc(1, sizeof(*s));
    s->pl = pl;
    s->restart = restart;
    orc_planner_allocate_workspace(pl, restart + 1);
    s->ip = (double *) calloc((size_t) (restart + 1) * (size_t) restart, sizeof(double));
    return s;
}

/* step :83-127 -- one restart cycle.  The least-squares solve is the reference's placeholder:
 * DummyTask returns 1 (src/UtilityTasks.cpp:96-99) and SOL += 1 * V_j for every j < restart. */
void orc_gmres_step(orc_gmres *s) {
    orc_planner *pl = s->pl;
    const int m = s->restart;
    double t1[1];
    orc_planner_matvec(pl, KRYLOV(0), V_SOL);
    t1[0] = -1.0;
    orc_planner_xpay(pl, KRYLOV(0), 1, t1, V_RHS);
    t1[0] = orc_scalar_rsqrt(orc_planner_dot(pl, KRYLOV(0), KRYLOV(0)));
    orc_planner_scal(pl, KRYLOV(0), 1, t1);
    for (int j = 0; j < m; ++j) {
        orc_planner_matvec(pl, KRYLOV(j + 1), KRYLOV(j));
        for (int k = 0; k <= j; ++k) {
            const double h = orc_planner_dot(pl, KRYLOV(k), KRYLOV(j + 1));
            s->ip[(size_t) k * m + j] = h;
            t1[0] = orc_scalar_neg(h);
            orc_planner_axpy(pl, KRYLOV(j + 1), 1, t1, KRYLOV(k));
        }
        const double d = orc_planner_dot(pl, KRYLOV(j + 1), KRYLOV(j + 1));
        s->ip[(size_t) (j + 1) * m + j] = orc_scalar_sqrt(d);
        if (j + 1 < m) {
            t1[0] = orc_scalar_rsqrt(d);
            orc_planner_scal(pl, KRYLOV(j + 1), 1, t1);
        }
    }
    t1[0] = orc_scalar_dummy();
    for (int j = 0; j < m; ++j) orc_planner_axpy(pl, V_SOL, 1, t1, KRYLOV(j));
}

void orc_gmres_hessenberg(orc_gmres *s, double *out) {
    memcpy(out, s->ip, sizeof(double) * (size_t) (s->restart + 1) * (size_t) s->restart);
}

void orc_gmres_destroy(orc_gmres *s) {
    if (s) { free(s->ip); free(s); }
}
