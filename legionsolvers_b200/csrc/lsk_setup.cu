// lsk_setup.cu -- the integer work that runs once before the hot path, on the GPU:
//   * the linearized stencil generator (FillLinearizedCSRStencilTask, src/StencilGenerator.cpp:380-543,
//     which the reference runs as an O(N) sequential CPU walk PER PIECE), and
//   * the dependent-partitioning primitives behind CSRMatrix / COOMatrix
//     (image_range, image, preimage, preimage_range: src/CSRMatrix.cpp:68-155, src/COOMatrix.cpp:56-141).
// Everything here is exact integer arithmetic; results are bit-identical to the oracle.
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <limits.h>

#include "lsk_common.cuh"

namespace lsk {

struct StencilDev {
    int dim, noff;
    long long shape[LSK_MAX_DIM];
    long long stride[LSK_MAX_DIM];  // linearisation stride of each dimension in the chosen order
    long long off[LSK_MAX_STENCIL][LSK_MAX_DIM];
    double val[LSK_MAX_STENCIL];
};

static bool make_dev(const lsk_stencil *st, StencilDev &d) {
    if (!st || st->dim < 1 || st->dim > LSK_MAX_DIM || st->noff < 0 || st->noff > LSK_MAX_STENCIL) return false;
    d.dim = st->dim;
    d.noff = st->noff;
    for (int i = 0; i < LSK_MAX_DIM; ++i) {
        d.shape[i] = i < st->dim ? st->shape[i] : 1;
        d.stride[i] = 0;
        if (i < st->dim && st->shape[i] <= 0) return false;
    }
    // linearize_row_major / linearize_column_major, src/StencilGenerator.hpp:230-259
    long long acc = 1;
    if (st->order == 0) {
        for (int i = st->dim - 1; i >= 0; --i) { d.stride[i] = acc; acc *= st->shape[i]; }
    } else {
        for (int i = 0; i < st->dim; ++i) { d.stride[i] = acc; acc *= st->shape[i]; }
    }
    for (int j = 0; j < st->noff; ++j) {
        for (int i = 0; i < LSK_MAX_DIM; ++i) d.off[j][i] = i < st->dim ? st->offsets[j][i] : 0;
        d.val[j] = st->values[j];
    }
    return true;
}

__device__ __forceinline__ void delinearize(const StencilDev &s, long long r, long long (&p)[LSK_MAX_DIM]) {
#pragma unroll
    for (int i = 0; i < LSK_MAX_DIM; ++i) p[i] = (i < s.dim) ? (r / s.stride[i]) % s.shape[i] : 0;
}

__device__ __forceinline__ int row_count(const StencilDev &s, const long long (&p)[LSK_MAX_DIM]) {
    int n = 0;
    for (int j = 0; j < s.noff; ++j) {
        bool in = true;
#pragma unroll
        for (int i = 0; i < LSK_MAX_DIM; ++i) {
            const long long q = p[i] + s.off[j][i];
            in = in && (i >= s.dim || (q >= 0 && q < s.shape[i]));
        }
        n += in ? 1 : 0;
    }
    return n;
}

// counts[i] = nnz of row r_lo + i (or only their sum when counts == nullptr)
__global__ void __launch_bounds__(kBlock)
stencil_count_kernel(StencilDev s, long long r_lo, long long nrows, long long *counts,
                     unsigned long long *total) {
    long long local = 0;
    for (long long i = (long long) blockIdx.x * kBlock + threadIdx.x; i < nrows; i += (long long) gridDim.x * kBlock) {
        long long p[LSK_MAX_DIM];
        delinearize(s, r_lo + i, p);
        const int c = row_count(s, p);
        if (counts) counts[i] = c;
        local += c;
    }
    if (total) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
        if ((threadIdx.x & 31) == 0 && local) atomicAdd(total, (unsigned long long) local);  // integer: exact
    }
}

__global__ void __launch_bounds__(kBlock)
stencil_fill_kernel(StencilDev s, long long r_lo, long long nrows, long long k_first,
                    const long long *__restrict__ start, double *__restrict__ entry,
                    long long *__restrict__ col, lsk_rect *__restrict__ rowptr) {
    for (long long i = (long long) blockIdx.x * kBlock + threadIdx.x; i < nrows; i += (long long) gridDim.x * kBlock) {
        long long p[LSK_MAX_DIM];
        delinearize(s, r_lo + i, p);
        long long k = start[i];  // slab-local position of the row's first non-zero
        const long long begin = k;
        for (int j = 0; j < s.noff; ++j) {
            bool in = true;
            long long lin = 0;
#pragma unroll
            for (int d = 0; d < LSK_MAX_DIM; ++d) {
                const long long q = p[d] + s.off[j][d];
                in = in && (d >= s.dim || (q >= 0 && q < s.shape[d]));
                lin += s.stride[d] * q;
            }
            if (in) {
                col[k] = lin;
                entry[k] = s.val[j];
                ++k;
            }
        }
        rowptr[i].lo = k_first + begin;
        rowptr[i].hi = k_first + k - 1;
    }
}

__global__ void __launch_bounds__(kBlock)
expand_rows_kernel(long long rows, long long r_lo, const lsk_rect *__restrict__ rowptr, long long k_base,
                   long long *__restrict__ row) {
    for (long long i = (long long) blockIdx.x * kBlock + threadIdx.x; i < rows; i += (long long) gridDim.x * kBlock) {
        const lsk_rect rc = rowptr[i];
        for (long long k = rc.lo; k <= rc.hi; ++k) row[k - k_base] = r_lo + i;
    }
}

// ---- {min, max, count} reductions over integer fields (atomics on integers are exact and order-free)
__device__ __forceinline__ void span_commit(long long mn, long long mx, long long cnt, long long *out3) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const long long a = __shfl_xor_sync(0xffffffffu, mn, o);
        const long long b = __shfl_xor_sync(0xffffffffu, mx, o);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        mn = a < mn ? a : mn;
        mx = b > mx ? b : mx;
    }
    if ((threadIdx.x & 31) == 0 && cnt > 0) {
        atomicMin(out3 + 0, mn);
        atomicMax(out3 + 1, mx);
        atomicAdd(reinterpret_cast<unsigned long long *>(out3 + 2), (unsigned long long) cnt);
    }
}

__global__ void span_init_kernel(long long *out3) {
    out3[0] = LLONG_MAX;
    out3[1] = LLONG_MIN;
    out3[2] = 0;
}

__global__ void __launch_bounds__(kBlock)
rect_span_kernel(long long rows, const lsk_rect *__restrict__ rowptr, long long *out3) {
    long long mn = LLONG_MAX, mx = LLONG_MIN, cnt = 0;
    for (long long i = (long long) blockIdx.x * kBlock + threadIdx.x; i < rows; i += (long long) gridDim.x * kBlock) {
        const lsk_rect rc = rowptr[i];
        if (rc.hi >= rc.lo) {
            mn = rc.lo < mn ? rc.lo : mn;
            mx = rc.hi > mx ? rc.hi : mx;
            cnt += rc.hi - rc.lo + 1;
        }
    }
    span_commit(mn, mx, cnt, out3);
}

__global__ void __launch_bounds__(kBlock)
minmax_kernel(long long n, const long long *__restrict__ f, long long *out3) {
    long long mn = LLONG_MAX, mx = LLONG_MIN, cnt = 0;
    for (long long i = (long long) blockIdx.x * kBlock + threadIdx.x; i < n; i += (long long) gridDim.x * kBlock) {
        const long long v = f[i];
        mn = v < mn ? v : mn;
        mx = v > mx ? v : mx;
        ++cnt;
    }
    span_commit(mn, mx, cnt, out3);
}

__global__ void __launch_bounds__(kBlock)
preimage_span_kernel(long long n, const long long *__restrict__ f, long long lo, long long hi, long long k_base,
                     long long *out3) {
    long long mn = LLONG_MAX, mx = LLONG_MIN, cnt = 0;
    for (long long i = (long long) blockIdx.x * kBlock + threadIdx.x; i < n; i += (long long) gridDim.x * kBlock) {
        const long long v = f[i];
        if (v >= lo && v <= hi) {
            const long long k = k_base + i;
            mn = k < mn ? k : mn;
            mx = k > mx ? k : mx;
            ++cnt;
        }
    }
    span_commit(mn, mx, cnt, out3);
}

__global__ void __launch_bounds__(kBlock)
image_range_flags_kernel(long long rows, const lsk_rect *__restrict__ rowptr, long long k_lo, long long k_n,
                         uint8_t *kflags) {
    for (long long i = (long long) blockIdx.x * kBlock + threadIdx.x; i < rows; i += (long long) gridDim.x * kBlock) {
        const lsk_rect rc = rowptr[i];
        for (long long k = rc.lo; k <= rc.hi; ++k)
            if (k >= k_lo && k < k_lo + k_n) kflags[k - k_lo] = 1;
    }
}

__global__ void __launch_bounds__(kBlock)
image_flags_kernel(long long n, const long long *__restrict__ f, const uint8_t *__restrict__ kflags,
                   long long out_lo, long long out_n, uint8_t *out_flags) {
    for (long long i = (long long) blockIdx.x * kBlock + threadIdx.x; i < n; i += (long long) gridDim.x * kBlock) {
        if (kflags && !kflags[i]) continue;
        const long long v = f[i] - out_lo;
        if (v >= 0 && v < out_n) out_flags[v] = 1;
    }
}

__global__ void __launch_bounds__(kBlock)
preimage_flags_kernel(long long n, const long long *__restrict__ f, long long lo, long long hi, uint8_t *kflags) {
    for (long long i = (long long) blockIdx.x * kBlock + threadIdx.x; i < n; i += (long long) gridDim.x * kBlock)
        kflags[i] = (f[i] >= lo && f[i] <= hi) ? 1 : 0;
}

__global__ void __launch_bounds__(kBlock)
preimage_range_flags_kernel(long long rows, const lsk_rect *__restrict__ rowptr, long long k_lo, long long k_n,
                            const uint8_t *__restrict__ kflags, uint8_t *rflags) {
    for (long long i = (long long) blockIdx.x * kBlock + threadIdx.x; i < rows; i += (long long) gridDim.x * kBlock) {
        const lsk_rect rc = rowptr[i];
        uint8_t hit = 0;
        for (long long k = rc.lo; k <= rc.hi && !hit; ++k)
            if (k >= k_lo && k < k_lo + k_n) hit = kflags[k - k_lo];
        rflags[i] = hit;
    }
}

static int grid_for(const lsk_ctx *ctx, int64_t n) { return stream_grid(ctx, n, 8); }

}  // namespace lsk

using namespace lsk;

extern "C" {

int lsk_stencil_sort(lsk_stencil *st) {
    if (!st || st->dim < 1 || st->dim > LSK_MAX_DIM || st->noff < 0 || st->noff > LSK_MAX_STENCIL) return LSK_E_INVALID;
    struct Item { int64_t o[LSK_MAX_DIM]; double v; };
    Item items[LSK_MAX_STENCIL];
    for (int j = 0; j < st->noff; ++j) {
        for (int d = 0; d < LSK_MAX_DIM; ++d) items[j].o[d] = d < st->dim ? st->offsets[j][d] : 0;
        items[j].v = st->values[j];
    }
    const int dim = st->dim, order = st->order;
    // compare_row_major / compare_column_major, ties broken by the entry (src/StencilGenerator.cpp:408-433)
    auto less_pt = [&](const Item &p, const Item &q) {
        if (order == 0) {
            for (int i = 0; i < dim; ++i) { if (p.o[i] < q.o[i]) return true; if (p.o[i] > q.o[i]) return false; }
        } else {
            for (int i = dim - 1; i >= 0; --i) { if (p.o[i] < q.o[i]) return true; if (p.o[i] > q.o[i]) return false; }
        }
        return false;
    };
    std::stable_sort(items, items + st->noff, [&](const Item &p, const Item &q) {
        if (less_pt(p, q)) return true;
        if (less_pt(q, p)) return false;
        return p.v < q.v;
    });
    for (int j = 0; j < st->noff; ++j) {
        for (int d = 0; d < LSK_MAX_DIM; ++d) st->offsets[j][d] = items[j].o[d];
        st->values[j] = items[j].v;
    }
    return 0;
}

int64_t lsk_stencil_size(const lsk_stencil *st) {
    if (!st || st->dim < 1 || st->dim > LSK_MAX_DIM) return -1;
    int64_t total = 0;
    for (int j = 0; j < st->noff; ++j) {
        int64_t prod = 1;
        for (int d = 0; d < st->dim; ++d) {
            int64_t a = st->offsets[j][d] < 0 ? -st->offsets[j][d] : st->offsets[j][d];
            int64_t len = st->shape[d] - a;
            prod *= len > 0 ? len : 0;
        }
        total += prod;
    }
    return total;
}

int lsk_stencil_count_f64(lsk_ctx *ctx, lsk_stream s, const lsk_stencil *st, int64_t r_lo, int64_t r_hi, int64_t *out) {
    StencilDev d;
    if (!ctx || !out || !make_dev(st, d) || r_lo < 0) return LSK_E_INVALID;
    const cudaStream_t cs = (cudaStream_t) s;
    LSK_RETURN_IF_CUDA(cudaMemsetAsync(out, 0, sizeof(int64_t), cs));
    const int64_t n = r_hi - r_lo + 1;
    if (n <= 0) return 0;
    stencil_count_kernel<<<grid_for(ctx, n), kBlock, 0, cs>>>(d, r_lo, n, nullptr, reinterpret_cast<unsigned long long *>(out));
    return after_launch(ctx);
}

int lsk_stencil_fill_csr_f64(lsk_ctx *ctx, lsk_stream s, const lsk_stencil *st, int64_t r_lo, int64_t r_hi,
                             int64_t k_first, double *entry, int64_t *col, lsk_rect *rowptr, int64_t *scratch) {
    StencilDev d;
    if (!ctx || !make_dev(st, d) || r_lo < 0) return LSK_E_INVALID;
    const int64_t n = r_hi - r_lo + 1;
    if (n <= 0) return 0;
    if (!entry || !col || !rowptr || !scratch) return LSK_E_INVALID;
    const cudaStream_t cs = (cudaStream_t) s;
    long long *counts = reinterpret_cast<long long *>(scratch);
    stencil_count_kernel<<<grid_for(ctx, n), kBlock, 0, cs>>>(d, r_lo, n, counts, nullptr);
    int rc = after_launch(ctx);
    if (rc) return rc;
    // exclusive prefix sum of the row lengths, in place (library scan: set-up path only)
    size_t temp_bytes = 0;
    LSK_RETURN_IF_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, temp_bytes, counts, counts, n, cs));
    void *temp = nullptr;
    LSK_RETURN_IF_CUDA(cudaMallocAsync(&temp, temp_bytes ? temp_bytes : 1, cs));
    cudaError_t e = cub::DeviceScan::ExclusiveSum(temp, temp_bytes, counts, counts, n, cs);
    cudaFreeAsync(temp, cs);
    if (e != cudaSuccess) return (int) e;
    stencil_fill_kernel<<<grid_for(ctx, n), kBlock, 0, cs>>>(d, r_lo, n, k_first, counts, entry,
                                                              reinterpret_cast<long long *>(col), rowptr);
    return after_launch(ctx);
}

int lsk_csr_expand_rows(lsk_ctx *ctx, lsk_stream s, int64_t rows, int64_t r_lo, const lsk_rect *rowptr,
                        int64_t k_base, int64_t *row) {
    if (!ctx || rows < 0 || (rows > 0 && (!rowptr || !row))) return LSK_E_INVALID;
    if (rows == 0) return 0;
    expand_rows_kernel<<<grid_for(ctx, rows), kBlock, 0, (cudaStream_t) s>>>(rows, r_lo, rowptr, k_base,
                                                                            reinterpret_cast<long long *>(row));
    return after_launch(ctx);
}

int lsk_rect_span_i64(lsk_ctx *ctx, lsk_stream s, int64_t rows, const lsk_rect *rowptr, int64_t *out3) {
    if (!ctx || !out3 || rows < 0 || (rows > 0 && !rowptr)) return LSK_E_INVALID;
    span_init_kernel<<<1, 1, 0, (cudaStream_t) s>>>(reinterpret_cast<long long *>(out3));
    if (rows > 0)
        rect_span_kernel<<<grid_for(ctx, rows), kBlock, 0, (cudaStream_t) s>>>(rows, rowptr, reinterpret_cast<long long *>(out3));
    return after_launch(ctx);
}

int lsk_minmax_i64(lsk_ctx *ctx, lsk_stream s, int64_t n, const int64_t *field, int64_t *out3) {
    if (!ctx || !out3 || n < 0 || (n > 0 && !field)) return LSK_E_INVALID;
    span_init_kernel<<<1, 1, 0, (cudaStream_t) s>>>(reinterpret_cast<long long *>(out3));
    if (n > 0)
        minmax_kernel<<<grid_for(ctx, n), kBlock, 0, (cudaStream_t) s>>>(n, reinterpret_cast<const long long *>(field),
                                                                        reinterpret_cast<long long *>(out3));
    return after_launch(ctx);
}

int lsk_preimage_span_i64(lsk_ctx *ctx, lsk_stream s, int64_t n, const int64_t *field, int64_t lo, int64_t hi,
                          int64_t k_base, int64_t *out3) {
    if (!ctx || !out3 || n < 0 || (n > 0 && !field)) return LSK_E_INVALID;
    span_init_kernel<<<1, 1, 0, (cudaStream_t) s>>>(reinterpret_cast<long long *>(out3));
    if (n > 0)
        preimage_span_kernel<<<grid_for(ctx, n), kBlock, 0, (cudaStream_t) s>>>(
            n, reinterpret_cast<const long long *>(field), lo, hi, k_base, reinterpret_cast<long long *>(out3));
    return after_launch(ctx);
}

int lsk_image_range_flags(lsk_ctx *ctx, lsk_stream s, int64_t rows, const lsk_rect *rowptr, int64_t k_lo,
                          int64_t k_n, uint8_t *kflags) {
    if (!ctx || rows < 0 || k_n < 0 || (rows > 0 && !rowptr) || (k_n > 0 && !kflags)) return LSK_E_INVALID;
    if (k_n > 0) LSK_RETURN_IF_CUDA(cudaMemsetAsync(kflags, 0, (size_t) k_n, (cudaStream_t) s));
    if (rows == 0 || k_n == 0) return 0;
    image_range_flags_kernel<<<grid_for(ctx, rows), kBlock, 0, (cudaStream_t) s>>>(rows, rowptr, k_lo, k_n, kflags);
    return after_launch(ctx);
}

int lsk_image_flags(lsk_ctx *ctx, lsk_stream s, int64_t n, const int64_t *field, const uint8_t *kflags,
                    int64_t out_lo, int64_t out_n, uint8_t *out_flags) {
    if (!ctx || n < 0 || out_n < 0 || (n > 0 && !field) || (out_n > 0 && !out_flags)) return LSK_E_INVALID;
    if (out_n > 0) LSK_RETURN_IF_CUDA(cudaMemsetAsync(out_flags, 0, (size_t) out_n, (cudaStream_t) s));
    if (n == 0 || out_n == 0) return 0;
    image_flags_kernel<<<grid_for(ctx, n), kBlock, 0, (cudaStream_t) s>>>(n, reinterpret_cast<const long long *>(field),
                                                                          kflags, out_lo, out_n, out_flags);
    return after_launch(ctx);
}

int lsk_preimage_flags(lsk_ctx *ctx, lsk_stream s, int64_t n, const int64_t *field, int64_t lo, int64_t hi,
                       uint8_t *kflags) {
    if (!ctx || n < 0 || (n > 0 && (!field || !kflags))) return LSK_E_INVALID;
    if (n == 0) return 0;
    preimage_flags_kernel<<<grid_for(ctx, n), kBlock, 0, (cudaStream_t) s>>>(n, reinterpret_cast<const long long *>(field),
                                                                             lo, hi, kflags);
    return after_launch(ctx);
}

int lsk_preimage_range_flags(lsk_ctx *ctx, lsk_stream s, int64_t rows, const lsk_rect *rowptr, int64_t k_lo,
                             int64_t k_n, const uint8_t *kflags, uint8_t *rflags) {
    if (!ctx || rows < 0 || (rows > 0 && (!rowptr || !rflags)) || (k_n > 0 && !kflags)) return LSK_E_INVALID;
    if (rows == 0) return 0;
    preimage_range_flags_kernel<<<grid_for(ctx, rows), kBlock, 0, (cudaStream_t) s>>>(rows, rowptr, k_lo, k_n, kflags, rflags);
    return after_launch(ctx);
}

/* Realm's dense equal split: piece i = [floor(n*i/P), floor(n*(i+1)/P) - 1]; pinned by the reference
 * only for n % P == 0 (see oracle/lsk_oracle.c orc_equal_partition). */
int lsk_equal_partition(int64_t n, int pieces, int64_t *lo, int64_t *hi) {
    if (n < 0 || pieces <= 0 || !lo || !hi) return LSK_E_INVALID;
    for (int i = 0; i < pieces; ++i) {
        lo[i] = (int64_t) (((__int128) n * i) / pieces);
        hi[i] = (int64_t) (((__int128) n * (i + 1)) / pieces) - 1;
    }
    return 0;
}

int lsk_shard(int64_t point, int64_t volume, int64_t total_shards) {
    if (total_shards <= 0) return LSK_E_INVALID;
    const int64_t per = (volume + total_shards - 1) / total_shards;
    return per > 0 ? (int) (point / per) : 0;
}

}  // extern "C"
