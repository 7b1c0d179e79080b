// lsk_rspmv.cu -- SURVEY.md section 8(f) rank 2: the transposed mat-vecs and the pieces of a REAL GMRES update.
//
// CSRRmatvecTask / COORmatvecTask are declared and given TaskIDs by the reference (src/TaskIDs.hpp:40-45; fp64 / 1-D /
// long long: 598624 and 583072) but their bodies are `assert(false)` (src/CSRMatrixTasks.cpp:94-100,
// src/COOMatrixTasks.cpp:77-83).  Semantics here follow the forward tasks: for every stored non-zero k of the piece
//     y[col_k] += entry_k * x[row_k]                     (y = y + A^T x, accumulating like the reference's sum-reduction
// accessor; guarded by the same row / column windows), with fp64 atomics -- the output index is the column, so several
// rows update one element.  Likewise GMRESSolver::step ends in a placeholder (DummyTask returns 1, SOL += 1 * v_j,
// src/GMRESSolver.hpp:109-126, src/UtilityTasks.cpp:96-99); the two kernels at the end of this file are what the
// finished algorithm needs: the small least-squares problem min || beta e1 - H y || by Givens rotations, and
// SOL += V y in ONE pass over the basis vectors.
#include "lsk_common.cuh"
#include "lsk_spmv_tma.cuh"

namespace lsk {

// V lanes per row walk the row's non-zeros; x[row] is read once per lane
template <int V>
__global__ void __launch_bounds__(kBlock)
csr_rspmv_kernel(int64_t rows, const double *__restrict__ entry, const long long *__restrict__ col, const lsk_rect *__restrict__ rowptr,
                 int64_t k_base, const double *__restrict__ x, double *y_shifted, long long col_lo, long long col_hi) {
    constexpr int RPC = kBlock / V;
    const int sub = threadIdx.x % V;
    for (int64_t base = (int64_t) blockIdx.x * RPC; base < rows; base += (int64_t) gridDim.x * RPC) {
        const int64_t row = base + threadIdx.x / V;
        if (row >= rows) continue;
        const longlong2 rc = __ldg(reinterpret_cast<const longlong2 *>(rowptr + row));
        const double xr = __ldg(x + row);
        for (long long j = rc.x - k_base + sub; j <= rc.y - k_base; j += V) {
            const long long c = load1_stream(col + j);
            if (c >= col_lo && c <= col_hi) atomicAdd(y_shifted + c, mul_rn(load1_stream(entry + j), xr));
        }
    }
}

__global__ void __launch_bounds__(kBlock)
coo_rspmv_kernel(int64_t nnz, const double *__restrict__ entry, const long long *__restrict__ row, const long long *__restrict__ col,
                 const double *__restrict__ x_shifted, double *y_shifted, long long row_lo, long long row_hi, long long col_lo, long long col_hi) {
    for (int64_t k = (int64_t) blockIdx.x * kBlock + threadIdx.x; k < nnz; k += (int64_t) gridDim.x * kBlock) {
        const long long r = load1_stream(row + k), c = load1_stream(col + k);
        if (r >= row_lo && r <= row_hi && c >= col_lo && c <= col_hi)
            atomicAdd(y_shifted + c, mul_rn(load1_stream(entry + k), __ldg(x_shifted + r)));
    }
}

// ---- GMRES: min || beta e1 - H y ||_2 for the (m + 1) x m upper Hessenberg H, by Givens rotations -----------------
// One thread: m <= 64, ~m^2 flops.  H is row-major with leading dimension ld (the solver's inner_products table).
constexpr int kMaxRestart = 64;
__global__ void gmres_solve_kernel(int m, const double *__restrict__ H, int ld, const double *beta_sq, double *y, double *resid) {
    __shared__ double R[kMaxRestart + 1][kMaxRestart];
    __shared__ double g[kMaxRestart + 1], cs[kMaxRestart], sn[kMaxRestart];
    if (threadIdx.x != 0) return;
    for (int i = 0; i <= m; ++i)
        for (int j = 0; j < m; ++j) R[i][j] = H[(size_t) i * ld + j];
    for (int i = 0; i <= m; ++i) g[i] = 0.0;
    g[0] = sqrt(*beta_sq);
    for (int j = 0; j < m; ++j) {
        for (int i = 0; i < j; ++i) {  // earlier rotations on column j
            const double a = R[i][j], b = R[i + 1][j];
            R[i][j] = cs[i] * a + sn[i] * b;
            R[i + 1][j] = -sn[i] * a + cs[i] * b;
        }
        const double a = R[j][j], b = R[j + 1][j];
        const double rho = hypot(a, b);
        cs[j] = rho > 0.0 ? a / rho : 1.0;
        sn[j] = rho > 0.0 ? b / rho : 0.0;
        R[j][j] = rho;
        R[j + 1][j] = 0.0;
        const double gj = g[j];
        g[j] = cs[j] * gj;
        g[j + 1] = -sn[j] * gj;
    }
    for (int i = m - 1; i >= 0; --i) {  // back substitution
        double s = g[i];
        for (int j = i + 1; j < m; ++j) s -= R[i][j] * y[j];
        y[i] = R[i][i] != 0.0 ? s / R[i][i] : 0.0;
    }
    if (resid) *resid = fabs(g[m]);  // the least-squares residual = || b - A x_new ||
}

// x[i] = fma(y[m-1], V[m-1][i], ... fma(y[0], V[0][i], x[i])): the m axpys of the update in one pass (same order, same
// rounding as m AxpyTasks with alpha = y[j]); the pointer table and the coefficients live on the device
__global__ void __launch_bounds__(kBlock)
multi_axpy_kernel(int64_t n, int m, const double *__restrict__ y, const double *const *__restrict__ V, double *__restrict__ x) {
    __shared__ double s_y[kMaxRestart];
    __shared__ const double *s_v[kMaxRestart];
    for (int j = threadIdx.x; j < m; j += kBlock) {
        s_y[j] = y[j];
        s_v[j] = V[j];
    }
    __syncthreads();
    for (int64_t i = (int64_t) blockIdx.x * kBlock + threadIdx.x; i < n; i += (int64_t) gridDim.x * kBlock) {
        double acc = x[i];
        int j = 0;
        for (; j + 4 <= m; j += 4) {  // four loads in flight; the fma chain stays in ascending j
            const double v0 = __ldg(s_v[j] + i), v1 = __ldg(s_v[j + 1] + i), v2 = __ldg(s_v[j + 2] + i), v3 = __ldg(s_v[j + 3] + i);
            acc = fma_rn(s_y[j], v0, acc);
            acc = fma_rn(s_y[j + 1], v1, acc);
            acc = fma_rn(s_y[j + 2], v2, acc);
            acc = fma_rn(s_y[j + 3], v3, acc);
        }
        for (; j < m; ++j) acc = fma_rn(s_y[j], __ldg(s_v[j] + i), acc);
        x[i] = acc;
    }
}

}  // namespace lsk

using namespace lsk;

extern "C" {

int lsk_csr_rspmv_f64(lsk_ctx *ctx, lsk_stream s, int64_t rows, int64_t nnz, const double *entry, const int64_t *col, const lsk_rect *rowptr,
                      int64_t k_base, const double *x, double *y_shifted, int64_t col_lo, int64_t col_hi) {
    if (!ctx || rows < 0 || nnz < 0) return LSK_E_INVALID;
    if (rows == 0 || nnz == 0) return 0;
    if (!entry || !col || !rowptr || !x || !y_shifted) return LSK_E_INVALID;
    const double mean = (double) nnz / (double) rows;
    const int V = mean <= 4.0 ? 2 : mean <= 8.0 ? 4 : mean <= 16.0 ? 8 : mean <= 48.0 ? 16 : 32;
    const int64_t passes = (rows * V + kBlock - 1) / kBlock;
    const int grid = stream_grid(ctx, passes * kBlock, 8);
    const long long *c = reinterpret_cast<const long long *>(col);
    const cudaStream_t st = (cudaStream_t) s;
    switch (V) {
    case 2: csr_rspmv_kernel<2><<<grid, kBlock, 0, st>>>(rows, entry, c, rowptr, k_base, x, y_shifted, col_lo, col_hi); break;
    case 4: csr_rspmv_kernel<4><<<grid, kBlock, 0, st>>>(rows, entry, c, rowptr, k_base, x, y_shifted, col_lo, col_hi); break;
    case 8: csr_rspmv_kernel<8><<<grid, kBlock, 0, st>>>(rows, entry, c, rowptr, k_base, x, y_shifted, col_lo, col_hi); break;
    case 16: csr_rspmv_kernel<16><<<grid, kBlock, 0, st>>>(rows, entry, c, rowptr, k_base, x, y_shifted, col_lo, col_hi); break;
    default: csr_rspmv_kernel<32><<<grid, kBlock, 0, st>>>(rows, entry, c, rowptr, k_base, x, y_shifted, col_lo, col_hi); break;
    }
    return after_launch(ctx);
}

int lsk_coo_rspmv_f64(lsk_ctx *ctx, lsk_stream s, int64_t nnz, const double *entry, const int64_t *row, const int64_t *col,
                      const double *x_shifted, double *y_shifted, int64_t row_lo, int64_t row_hi, int64_t col_lo, int64_t col_hi) {
    if (!ctx || nnz < 0) return LSK_E_INVALID;
    if (nnz == 0) return 0;
    if (!entry || !row || !col || !x_shifted || !y_shifted) return LSK_E_INVALID;
    const int grid = stream_grid(ctx, nnz, 8);
    coo_rspmv_kernel<<<grid, kBlock, 0, (cudaStream_t) s>>>(nnz, entry, reinterpret_cast<const long long *>(row), reinterpret_cast<const long long *>(col),
                                                            x_shifted, y_shifted, row_lo, row_hi, col_lo, col_hi);
    return after_launch(ctx);
}

int lsk_gmres_solve_f64(lsk_ctx *ctx, lsk_stream s, int m, const double *H, int ld, const double *beta_sq, double *y, double *resid) {
    if (!ctx || m < 1 || m > kMaxRestart || !H || ld < m || !beta_sq || !y) return LSK_E_INVALID;
    gmres_solve_kernel<<<1, 32, 0, (cudaStream_t) s>>>(m, H, ld, beta_sq, y, resid);
    return after_launch(ctx);
}

int lsk_multi_axpy_f64(lsk_ctx *ctx, lsk_stream s, int64_t n, int m, const double *y, const double *const *V, double *x) {
    if (!ctx || n < 0 || m < 0 || m > kMaxRestart || (m > 0 && (!y || !V)) || (n > 0 && !x)) return LSK_E_INVALID;
    if (n == 0 || m == 0) return 0;
    const int grid = stream_grid(ctx, n, 8);
    multi_axpy_kernel<<<grid, kBlock, 0, (cudaStream_t) s>>>(n, m, y, V, x);
    return after_launch(ctx);
}

}  // extern "C"
