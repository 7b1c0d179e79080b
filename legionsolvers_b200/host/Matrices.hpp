// Matrices.hpp -- AbstractLinearOperator / AbstractMatrix / CSRMatrix / COOMatrix of the reference
// (src/AbstractLinearOperator.hpp, src/AbstractMatrix.hpp, src/CSRMatrix.{hpp,cpp},
// src/COOMatrix.{hpp,cpp}) over device memory.
//
// Layout is the reference's, field for field: CSR = kernel region {entry: ENTRY_T, col: int64} over
// [0, nnz) + rowptr region {rowptr: Rect<1,int64> INCLUSIVE, global k} over [0, rows); COO = kernel
// region {entry, row, col}.  A rank holds the SLAB of the matrix its colours need: rows
// [slab_r_lo, slab_r_hi] and their non-zeros [slab_k_lo, slab_k_hi].
//
// The partition derivations keep the reference's names.  They are computed on the GPU from the
// device-resident fields (lsk_rect_span_i64 / lsk_minmax_i64 / lsk_preimage_span_i64) and return the
// bounding interval of each LOCAL colour's piece, which is what a physical instance covers and what
// the mat-vec task receives; the exact index sets are available through the *_flags kernels.
#pragma once

#include <memory>
#include <vector>

#include "PartitionedVector.hpp"

namespace LegionSolvers {

// optional fusion of the mat-vec with the dot products that follow it in the solvers
template <typename T>
struct MatvecFusion {
    const PartitionedVector<T> *w = nullptr;  // y.w wanted (per-piece partial slots in yw)
    std::vector<T *> yw, yy;                  // indexed by colour; null = not wanted
};

template <typename T>
class AbstractLinearOperator {
public:
    virtual ~AbstractLinearOperator() = default;
    // domain_partition_from_range_partition (src/AbstractMatrix.cpp): ghost pieces of the domain
    virtual IntervalPartition domain_partition_from_range_partition(int64_t domain_volume,
                                                                    const IndexPartition &range_partition) const = 0;
};

template <typename T>
class AbstractMatrix : public AbstractLinearOperator<T> {
public:
    virtual int64_t get_kernel_volume() const = 0;
    virtual int64_t rows() const = 0;
    virtual int64_t cols() const = 0;
    virtual IntervalPartition create_kernel_partition_from_range_partition(const IndexPartition &range_partition) const = 0;
    virtual IntervalPartition create_domain_partition_from_kernel_partition(int64_t domain_volume,
                                                                            const IntervalPartition &kernel_partition,
                                                                            const IndexPartition &colors) const = 0;
    // accumulate: dst += A src instead of dst = A src (matrices that overwrite, i.e. CSR: a further block on the same rows)
    virtual void matvec(PartitionedVector<T> &dst, const PartitionedVector<T> &src, const IntervalPartition &kernel_partition,
                        const IntervalPartition &ghost_partition, const MatvecFusion<T> *fusion = nullptr,
                        bool accumulate = false) const = 0;
    // dst += A^T src (CSRRmatvecTask / COORmatvecTask, reserved but unimplemented in the reference): dst lives on the DOMAIN
    // space, src on the range space; accumulates.  dst must hold every column the local pieces reference.
    virtual void rmatvec(PartitionedVector<T> &dst, const PartitionedVector<T> &src, const IntervalPartition &kernel_partition,
                         const IntervalPartition &ghost_partition) const = 0;
    virtual bool overwrites_output() const = 0;  // CSR: beta = 0; COO: beta = 1 (reference GPU variants)

    IntervalPartition domain_partition_from_range_partition(int64_t domain_volume,
                                                            const IndexPartition &range_partition) const override {
        return create_domain_partition_from_kernel_partition(
            domain_volume, create_kernel_partition_from_range_partition(range_partition), range_partition);
    }
};

namespace detail {
struct Span3 {
    int64_t mn, mx, count;
};
template <class F>
inline Span3 device_span(Runtime *rt, F &&launch) {
    DeviceBuffer<int64_t> out(rt, 3);
    int64_t *o = out.ptr;
    rt->enqueue("span", [&] { return launch(o); });
    Span3 h;
    rt->check_cuda(cudaMemcpyAsync(&h, o, sizeof(h), cudaMemcpyDeviceToHost, rt->stream()), "span D2H");
    rt->fence();
    return h;
}
}  // namespace detail

template <typename T>
struct SpmvKernels;
template <>
struct SpmvKernels<double> {
    static int csr(lsk_ctx *c, cudaStream_t s, int64_t rows, int64_t nnz, const double *e, const int64_t *col,
                   const lsk_rect *rp, int64_t kb, const double *x, double *y, const double *w, double *o, double *oyy, bool acc = false) {
        return lsk_csr_spmv_f64(c, s, rows, nnz, e, col, rp, kb, x, y, w, o, oyy, LSK_SPMV_AUTO | (acc ? LSK_SPMV_ACCUMULATE : 0));
    }
    static int coo(lsk_ctx *c, cudaStream_t s, int64_t nnz, const double *e, const int64_t *row, const int64_t *col,
                   const double *x, double *y, int64_t rl, int64_t rh, int64_t cl, int64_t ch) {
        return lsk_coo_spmv_f64(c, s, nnz, e, row, col, x, y, rl, rh, cl, ch);
    }
};
template <>
struct SpmvKernels<float> {
    static int csr(lsk_ctx *c, cudaStream_t s, int64_t rows, int64_t nnz, const float *e, const int64_t *col,
                   const lsk_rect *rp, int64_t kb, const float *x, float *y, const float *w, float *o, float *oyy, bool acc = false) {
        return lsk_csr_spmv_f32(c, s, rows, nnz, e, col, rp, kb, x, y, w, o, oyy, LSK_SPMV_AUTO | (acc ? LSK_SPMV_ACCUMULATE : 0));
    }
    static int coo(lsk_ctx *c, cudaStream_t s, int64_t nnz, const float *e, const int64_t *row, const int64_t *col,
                   const float *x, float *y, int64_t rl, int64_t rh, int64_t cl, int64_t ch) {
        return lsk_coo_spmv_f32(c, s, nnz, e, row, col, x, y, rl, rh, cl, ch);
    }
};

// ====================================================================================================
// CSRMatrix
// ====================================================================================================
template <typename T>
class CSRMatrix : public AbstractMatrix<T> {
    Runtime *rt;
    int64_t n_rows, n_cols, nnz_global;
    int64_t slab_r_lo, slab_r_hi, slab_k_lo, slab_k_hi;
    DeviceBuffer<T> entry;        // fid_entry, element 0 <-> k = slab_k_lo
    DeviceBuffer<int64_t> col;    // fid_col
    DeviceBuffer<lsk_rect> rowptr;  // fid_rowptr, element 0 <-> row slab_r_lo

public:
    // adopt device arrays (generator) or upload host arrays (h_* non-null) of the slab
    CSRMatrix(Runtime *rt_, int64_t rows_, int64_t cols_, int64_t nnz_global_, int64_t r_lo, int64_t r_hi, int64_t k_lo,
              int64_t k_hi, const T *h_entry = nullptr, const int64_t *h_col = nullptr, const lsk_rect *h_rowptr = nullptr)
        : rt(rt_), n_rows(rows_), n_cols(cols_), nnz_global(nnz_global_), slab_r_lo(r_lo), slab_r_hi(r_hi),
          slab_k_lo(k_lo), slab_k_hi(k_hi), entry(rt_, (size_t) std::max<int64_t>(0, k_hi - k_lo + 1)),
          col(rt_, (size_t) std::max<int64_t>(0, k_hi - k_lo + 1)), rowptr(rt_, (size_t) std::max<int64_t>(0, r_hi - r_lo + 1)) {
        if (h_entry) rt->check_cuda(cudaMemcpyAsync(entry.ptr, h_entry, sizeof(T) * entry.count, cudaMemcpyHostToDevice, rt->stream()), "entry H2D");
        if (h_col) rt->check_cuda(cudaMemcpyAsync(col.ptr, h_col, sizeof(int64_t) * col.count, cudaMemcpyHostToDevice, rt->stream()), "col H2D");
        if (h_rowptr) rt->check_cuda(cudaMemcpyAsync(rowptr.ptr, h_rowptr, sizeof(lsk_rect) * rowptr.count, cudaMemcpyHostToDevice, rt->stream()), "rowptr H2D");
        if (h_entry || h_col || h_rowptr) rt->fence();
    }

    Runtime *runtime() const { return rt; }
    int64_t get_kernel_volume() const override { return nnz_global; }
    int64_t rows() const override { return n_rows; }
    int64_t cols() const override { return n_cols; }
    bool overwrites_output() const override { return true; }
    T *entry_ptr() const { return entry.ptr; }
    int64_t *col_ptr() const { return col.ptr; }
    lsk_rect *rowptr_ptr() const { return rowptr.ptr; }
    int64_t slab_rows_lo() const { return slab_r_lo; }
    int64_t slab_rows_hi() const { return slab_r_hi; }
    int64_t slab_kernel_lo() const { return slab_k_lo; }
    int64_t slab_kernel_hi() const { return slab_k_hi; }

    // create_partition_by_image_range over fid_rowptr (src/CSRMatrix.cpp:89-109)
    IntervalPartition create_kernel_partition_from_range_partition(const IndexPartition &range) const override {
        IntervalPartition kp(range.pieces);
        for (int c = range.first_color; c < range.end_color; ++c) {
            const int64_t lo = range.lo[(size_t) c], hi = range.hi[(size_t) c];
            if (hi < lo) continue;
            if (lo < slab_r_lo || hi > slab_r_hi) rt->fail(LSK_E_INVALID, "CSR slab does not hold the rows of a local colour");
            const lsk_rect *rp = rowptr.ptr + (lo - slab_r_lo);
            const detail::Span3 s = detail::device_span(rt, [&](int64_t *o) { return lsk_rect_span_i64(rt->ctx(), rt->stream(), hi - lo + 1, rp, o); });
            if (s.count == 0) continue;
            if (s.mx - s.mn + 1 != s.count) rt->fail(LSK_E_INVALID, "CSR kernel piece is not a contiguous run of non-zeros");
            kp.lo[(size_t) c] = s.mn;
            kp.hi[(size_t) c] = s.mx;
        }
        return kp;
    }

    // create_partition_by_image over fid_col (src/CSRMatrix.cpp:112-132)
    IntervalPartition create_domain_partition_from_kernel_partition(int64_t domain_volume, const IntervalPartition &kp,
                                                                    const IndexPartition &colors) const override {
        IntervalPartition gp(colors.pieces);
        for (int c = colors.first_color; c < colors.end_color; ++c) {
            const int64_t lo = kp.lo[(size_t) c], hi = kp.hi[(size_t) c];
            if (hi < lo) continue;
            if (lo < slab_k_lo || hi > slab_k_hi) rt->fail(LSK_E_INVALID, "CSR slab does not hold a local kernel piece");
            const int64_t *cp = col.ptr + (lo - slab_k_lo);
            const detail::Span3 s = detail::device_span(rt, [&](int64_t *o) { return lsk_minmax_i64(rt->ctx(), rt->stream(), hi - lo + 1, cp, o); });
            gp.lo[(size_t) c] = std::max<int64_t>(0, s.mn);  // image is intersected with the parent space
            gp.hi[(size_t) c] = std::min<int64_t>(domain_volume - 1, s.mx);
        }
        return gp;
    }

    void rmatvec(PartitionedVector<T> &dst, const PartitionedVector<T> &src, const IntervalPartition &kp,
                 const IntervalPartition &gp) const override {
        if constexpr (!std::is_same<T, double>::value) {
            rt->fail(LSK_E_INVALID, "transposed mat-vec is instantiated for fp64");
        } else {
            const IndexPartition &p = src.partition();  // rows
            for (int c = p.first_color; c < p.end_color; ++c) {
                const int64_t r_lo = p.lo[(size_t) c], nrow = p.piece_size(c);
                const int64_t k_lo = kp.lo[(size_t) c], nk = kp.hi[(size_t) c] - k_lo + 1;
                if (nk <= 0 || nrow <= 0) continue;
                if (gp.lo[(size_t) c] < dst.buf_lo() || gp.hi[(size_t) c] > dst.buf_hi())
                    rt->fail(LSK_E_INVALID, "destination vector does not hold the columns of the piece");
                const T *e = entry.ptr + (k_lo - slab_k_lo);
                const int64_t *cc = col.ptr + (k_lo - slab_k_lo);
                const lsk_rect *rp = rowptr.ptr + (r_lo - slab_r_lo);
                const T *x = src.ptr(r_lo);
                T *y = dst.shifted();
                const int64_t cl = gp.lo[(size_t) c], ch = gp.hi[(size_t) c];
                rt->enqueue("csr rmatvec", [&] { return lsk_csr_rspmv_f64(rt->ctx(), rt->stream(), nrow, nk, e, cc, rp, k_lo, x, y, cl, ch); });
            }
        }
    }

    // CSRMatrix::matvec (src/CSRMatrix.cpp:158-214): one CSRMatvecTask per piece of dst with regions
    // {dst piece, kernel piece, rowptr piece, ghost piece of src}
    void matvec(PartitionedVector<T> &dst, const PartitionedVector<T> &src, const IntervalPartition &kp,
                const IntervalPartition &gp, const MatvecFusion<T> *fusion = nullptr, bool accumulate = false) const override {
        const IndexPartition &p = dst.partition();
        for (int c = p.first_color; c < p.end_color; ++c) {
            const int64_t r_lo = p.lo[(size_t) c], nrow = p.piece_size(c);
            const int64_t k_lo = kp.lo[(size_t) c], nk = kp.hi[(size_t) c] - k_lo + 1;
            if (gp.hi[(size_t) c] >= gp.lo[(size_t) c] && (gp.lo[(size_t) c] < src.buf_lo() || gp.hi[(size_t) c] > src.buf_hi()))
                rt->fail(LSK_E_INVALID, "source vector does not hold the ghost piece");
            const T *e = entry.ptr + (nk > 0 ? k_lo - slab_k_lo : 0);
            const int64_t *cc = col.ptr + (nk > 0 ? k_lo - slab_k_lo : 0);
            const lsk_rect *rp = rowptr.ptr + (r_lo - slab_r_lo);
            const T *x = src.shifted();
            T *y = dst.ptr(r_lo);
            const T *w = (fusion && fusion->w && fusion->yw[(size_t) c]) ? fusion->w->ptr(r_lo) : nullptr;
            T *ow = w ? fusion->yw[(size_t) c] : nullptr;
            T *oyy = (fusion && !fusion->yy.empty()) ? fusion->yy[(size_t) c] : nullptr;
            rt->enqueue("csr matvec", [&] {
                return SpmvKernels<T>::csr(rt->ctx(), rt->stream(), nrow, nk > 0 ? nk : 0, e, cc, rp, nk > 0 ? k_lo : 0, x, y, w, ow, oyy, accumulate);
            });
        }
    }
};

// ====================================================================================================
// COOMatrix
// ====================================================================================================
template <typename T>
class COOMatrix : public AbstractMatrix<T> {
    Runtime *rt;
    int64_t n_rows, n_cols, nnz_global;
    int64_t slab_k_lo, slab_k_hi;
    DeviceBuffer<T> entry;
    DeviceBuffer<int64_t> row, col;

public:
    COOMatrix(Runtime *rt_, int64_t rows_, int64_t cols_, int64_t nnz_global_, int64_t k_lo, int64_t k_hi,
              const T *h_entry = nullptr, const int64_t *h_row = nullptr, const int64_t *h_col = nullptr)
        : rt(rt_), n_rows(rows_), n_cols(cols_), nnz_global(nnz_global_), slab_k_lo(k_lo), slab_k_hi(k_hi),
          entry(rt_, (size_t) std::max<int64_t>(0, k_hi - k_lo + 1)), row(rt_, (size_t) std::max<int64_t>(0, k_hi - k_lo + 1)),
          col(rt_, (size_t) std::max<int64_t>(0, k_hi - k_lo + 1)) {
        if (h_entry) rt->check_cuda(cudaMemcpyAsync(entry.ptr, h_entry, sizeof(T) * entry.count, cudaMemcpyHostToDevice, rt->stream()), "entry H2D");
        if (h_row) rt->check_cuda(cudaMemcpyAsync(row.ptr, h_row, sizeof(int64_t) * row.count, cudaMemcpyHostToDevice, rt->stream()), "row H2D");
        if (h_col) rt->check_cuda(cudaMemcpyAsync(col.ptr, h_col, sizeof(int64_t) * col.count, cudaMemcpyHostToDevice, rt->stream()), "col H2D");
        if (h_entry || h_row || h_col) rt->fence();
    }

    int64_t get_kernel_volume() const override { return nnz_global; }
    int64_t rows() const override { return n_rows; }
    int64_t cols() const override { return n_cols; }
    bool overwrites_output() const override { return false; }
    T *entry_ptr() const { return entry.ptr; }
    int64_t *row_ptr() const { return row.ptr; }
    int64_t *col_ptr() const { return col.ptr; }
    int64_t slab_kernel_lo() const { return slab_k_lo; }
    int64_t slab_kernel_hi() const { return slab_k_hi; }

    // create_partition_by_preimage over fid_row (src/COOMatrix.cpp:77-96)
    IntervalPartition create_kernel_partition_from_range_partition(const IndexPartition &range) const override {
        IntervalPartition kp(range.pieces);
        const int64_t n = slab_k_hi - slab_k_lo + 1;
        for (int c = range.first_color; c < range.end_color; ++c) {
            const int64_t lo = range.lo[(size_t) c], hi = range.hi[(size_t) c];
            if (hi < lo || n <= 0) continue;
            const detail::Span3 s = detail::device_span(rt, [&](int64_t *o) {
                return lsk_preimage_span_i64(rt->ctx(), rt->stream(), n, row.ptr, lo, hi, slab_k_lo, o);
            });
            if (s.count == 0) continue;
            if (s.mx - s.mn + 1 != s.count)
                rt->fail(LSK_E_INVALID, "COO kernel piece is not a contiguous run (entries must be grouped by row block)");
            kp.lo[(size_t) c] = s.mn;
            kp.hi[(size_t) c] = s.mx;
        }
        return kp;
    }

    // create_partition_by_image over fid_col (src/COOMatrix.cpp:98-118)
    IntervalPartition create_domain_partition_from_kernel_partition(int64_t domain_volume, const IntervalPartition &kp,
                                                                    const IndexPartition &colors) const override {
        IntervalPartition gp(colors.pieces);
        for (int c = colors.first_color; c < colors.end_color; ++c) {
            const int64_t lo = kp.lo[(size_t) c], hi = kp.hi[(size_t) c];
            if (hi < lo) continue;
            const int64_t *cp = col.ptr + (lo - slab_k_lo);
            const detail::Span3 s = detail::device_span(rt, [&](int64_t *o) { return lsk_minmax_i64(rt->ctx(), rt->stream(), hi - lo + 1, cp, o); });
            gp.lo[(size_t) c] = std::max<int64_t>(0, s.mn);
            gp.hi[(size_t) c] = std::min<int64_t>(domain_volume - 1, s.mx);
        }
        return gp;
    }

    void rmatvec(PartitionedVector<T> &dst, const PartitionedVector<T> &src, const IntervalPartition &kp,
                 const IntervalPartition &gp) const override {
        if constexpr (!std::is_same<T, double>::value) {
            rt->fail(LSK_E_INVALID, "transposed mat-vec is instantiated for fp64");
        } else {
            const IndexPartition &p = src.partition();  // rows
            for (int c = p.first_color; c < p.end_color; ++c) {
                const int64_t k_lo = kp.lo[(size_t) c], nk = kp.hi[(size_t) c] - k_lo + 1;
                if (nk <= 0) continue;
                if (gp.lo[(size_t) c] < dst.buf_lo() || gp.hi[(size_t) c] > dst.buf_hi())
                    rt->fail(LSK_E_INVALID, "destination vector does not hold the columns of the piece");
                const T *e = entry.ptr + (k_lo - slab_k_lo);
                const int64_t *rr = row.ptr + (k_lo - slab_k_lo), *cc = col.ptr + (k_lo - slab_k_lo);
                const T *x = src.shifted();
                T *y = dst.shifted();
                const int64_t rl = p.lo[(size_t) c], rh = p.hi[(size_t) c], cl = gp.lo[(size_t) c], ch = gp.hi[(size_t) c];
                rt->enqueue("coo rmatvec", [&] { return lsk_coo_rspmv_f64(rt->ctx(), rt->stream(), nk, e, rr, cc, x, y, rl, rh, cl, ch); });
            }
        }
    }

    // COOMatrix::matvec (src/COOMatrix.cpp:144-191): accumulates into dst (beta = 1)
    void matvec(PartitionedVector<T> &dst, const PartitionedVector<T> &src, const IntervalPartition &kp,
                const IntervalPartition &gp, const MatvecFusion<T> * = nullptr, bool = false) const override {
        const IndexPartition &p = dst.partition();
        for (int c = p.first_color; c < p.end_color; ++c) {
            const int64_t k_lo = kp.lo[(size_t) c], nk = kp.hi[(size_t) c] - k_lo + 1;
            if (nk <= 0) continue;
            if (gp.lo[(size_t) c] < src.buf_lo() || gp.hi[(size_t) c] > src.buf_hi())
                rt->fail(LSK_E_INVALID, "source vector does not hold the ghost piece");
            const T *e = entry.ptr + (k_lo - slab_k_lo);
            const int64_t *rr = row.ptr + (k_lo - slab_k_lo), *cc = col.ptr + (k_lo - slab_k_lo);
            const T *x = src.shifted();
            T *y = dst.shifted();
            const int64_t rl = p.lo[(size_t) c], rh = p.hi[(size_t) c], cl = gp.lo[(size_t) c], ch = gp.hi[(size_t) c];
            rt->enqueue("coo matvec", [&] { return SpmvKernels<T>::coo(rt->ctx(), rt->stream(), nk, e, rr, cc, x, y, rl, rh, cl, ch); });
        }
    }
};

}  // namespace LegionSolvers
