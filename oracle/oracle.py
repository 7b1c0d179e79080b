"""ctypes/numpy front end of the CPU ORACLE (oracle/lsk_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs -- as the checker or the reported CPU baseline, never by
the product package (legionsolvers_b200/ must not import this module).

The reference itself (C++ over Legion/Realm) cannot be compiled here: every source includes
<legion.h> (e.g. src/LegionUtilities.hpp:10) and Legion is neither installed nor fetchable, so
there is no oracle/_ref.  See lsk_oracle.h for what pins this restatement.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "_build" / "liblsk_oracle.so"

RECT_DTYPE = np.dtype([("lo", np.int64), ("hi", np.int64)])  # Legion::Rect<1, long long>


def build(force: bool = False) -> Path:
    src = [_HERE / "lsk_oracle.c", _HERE / "lsk_oracle.h", _HERE / "Makefile"]
    stale = (not _LIB_PATH.exists()) or any(s.stat().st_mtime > _LIB_PATH.stat().st_mtime for s in src)
    if force or stale:
        env = {k: v for k, v in os.environ.items() if k not in ("CC", "CFLAGS")}
        subprocess.run(["make", "-C", str(_HERE)], check=True, capture_output=True, env=env)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(str(build()))
        _declare(_lib)
    return _lib


def _p(a, dtype=None):
    if a is None:
        return None
    assert isinstance(a, np.ndarray) and a.flags["C_CONTIGUOUS"]
    if dtype is not None:
        assert a.dtype == dtype, (a.dtype, dtype)
    return a.ctypes.data_as(C.c_void_p)


def _declare(L):
    i64, dbl, vp, ci = C.c_int64, C.c_double, C.c_void_p, C.c_int
    L.orc_set_threads.argtypes = [ci]
    L.orc_set_dot_order.argtypes = [ci]
    L.orc_get_threads.restype = ci
    L.orc_get_alpha.argtypes = [ci, vp]
    L.orc_get_alpha.restype = dbl
    for name in ("neg", "sqrt", "rsqrt"):
        f = getattr(L, f"orc_scalar_{name}")
        f.argtypes, f.restype = [dbl], dbl
    for name in ("add", "sub", "mul", "div"):
        f = getattr(L, f"orc_scalar_{name}")
        f.argtypes, f.restype = [dbl, dbl], dbl
    L.orc_scalar_dummy.restype = dbl
    L.orc_scal.argtypes = [i64, dbl, vp]
    L.orc_axpy.argtypes = [i64, dbl, vp, vp]
    L.orc_xpay.argtypes = [i64, dbl, vp, vp]
    L.orc_dot.argtypes = [i64, vp, vp]
    L.orc_dot.restype = dbl
    L.orc_scal_f32.argtypes = [i64, C.c_float, vp]
    L.orc_axpy_f32.argtypes = [i64, C.c_float, vp, vp]
    L.orc_xpay_f32.argtypes = [i64, C.c_float, vp, vp]
    L.orc_dot_f32.argtypes = [i64, vp, vp]
    L.orc_dot_f32.restype = C.c_float
    L.orc_csr_matvec_literal.argtypes = [i64] * 6 + [vp] * 5
    L.orc_csr_matvec.argtypes = [i64] * 6 + [vp] * 5
    L.orc_csr_matvec.restype = ci
    L.orc_coo_matvec.argtypes = [i64] * 6 + [vp] * 5
    L.orc_coo_rmatvec.argtypes = [i64] * 6 + [vp] * 5
    L.orc_csr_rmatvec.argtypes = [i64] * 4 + [vp] * 5
    L.orc_laplacian_1d_coo.argtypes = [i64, i64, vp, vp, vp]
    L.orc_laplacian_1d_csr.argtypes = [i64, i64, vp, vp]
    L.orc_laplacian_1d_rowptr.argtypes = [i64, i64, i64, vp]
    L.orc_laplacian_2d_kernel_size.argtypes = [i64, i64]
    L.orc_laplacian_2d_kernel_size.restype = i64
    L.orc_stencil_size.argtypes = [ci, vp, vp, ci, vp]
    L.orc_stencil_size.restype = i64
    L.orc_sort_stencil.argtypes = [ci, ci, vp, vp, ci]
    L.orc_fill_linearized_csr_stencil.argtypes = [ci, vp, vp, ci, vp, vp, ci, i64, i64, i64, i64, vp, vp, vp]
    L.orc_fill_linearized_coo_stencil.argtypes = [ci, vp, vp, ci, vp, vp, ci, i64, i64, vp, vp, vp]
    L.orc_equal_partition.argtypes = [i64, ci, vp, vp]
    L.orc_image_range.argtypes = [vp, i64, i64, i64, vp]
    L.orc_image.argtypes = [vp, vp, i64, i64, vp]
    L.orc_preimage.argtypes = [vp, i64, i64, i64, vp]
    L.orc_preimage_range.argtypes = [vp, i64, vp, vp]
    L.orc_shard.argtypes = [i64, i64, i64]
    L.orc_shard.restype = ci
    L.orc_planner_create.argtypes = [ci, vp, vp]
    L.orc_planner_create.restype = vp
    L.orc_planner_destroy.argtypes = [vp]
    L.orc_planner_add_matrix.argtypes = [vp, ci, ci, i64, vp, vp, vp, vp]
    L.orc_planner_add_matrix.restype = ci
    L.orc_planner_allocate_workspace.argtypes = [vp, ci]
    L.orc_planner_vector.argtypes = [vp, ci, ci]
    L.orc_planner_vector.restype = C.POINTER(C.c_double)
    L.orc_planner_piece_bounds.argtypes = [vp, ci, ci, vp, vp]
    L.orc_planner_kernel_bounds.argtypes = [vp, ci, ci, vp, vp]
    L.orc_planner_ghost_bounds.argtypes = [vp, ci, ci, vp, vp]
    L.orc_planner_use_literal_csr.argtypes = [vp, ci]
    L.orc_planner_fill.argtypes = [vp, ci, dbl]
    L.orc_planner_copy.argtypes = [vp, ci, ci]
    L.orc_planner_scal.argtypes = [vp, ci, ci, vp]
    L.orc_planner_axpy.argtypes = [vp, ci, ci, vp, ci]
    L.orc_planner_xpay.argtypes = [vp, ci, ci, vp, ci]
    L.orc_planner_dot.argtypes = [vp, ci, ci]
    L.orc_planner_dot.restype = dbl
    L.orc_planner_matvec.argtypes = [vp, ci, ci]
    L.orc_cg_create.argtypes = [vp]
    L.orc_cg_create.restype = vp
    L.orc_cg_step.argtypes = [vp]
    L.orc_cg_history.argtypes = [vp, vp, i64]
    L.orc_cg_history.restype = i64
    L.orc_cg_destroy.argtypes = [vp]
    L.orc_bicgstab_create.argtypes = [vp]
    L.orc_bicgstab_create.restype = vp
    L.orc_bicgstab_step.argtypes = [vp]
    L.orc_bicgstab_history.argtypes = [vp, ci, vp, i64]
    L.orc_bicgstab_history.restype = i64
    L.orc_bicgstab_destroy.argtypes = [vp]
    L.orc_gmres_create.argtypes = [vp, ci]
    L.orc_gmres_create.restype = vp
    L.orc_gmres_step.argtypes = [vp]
    L.orc_gmres_hessenberg.argtypes = [vp, vp]
    L.orc_gmres_destroy.argtypes = [vp]


# --------------------------------------------------------------------------------------------
# scalars / BLAS-1
# --------------------------------------------------------------------------------------------
def set_threads(n: int) -> None:
    lib().orc_set_threads(int(n))


def set_dot_order(order: int) -> None:
    """0 = the reference's sequential dot (default); 1 = pairwise tree -- another valid order, for sensitivity measurements."""
    lib().orc_set_dot_order(int(order))


def get_alpha(terms) -> float:
    t = np.ascontiguousarray(terms, dtype=np.float64)
    return lib().orc_get_alpha(len(t), _p(t))


def scalar(op: str, x: float, y: float | None = None) -> float:
    f = getattr(lib(), f"orc_scalar_{op}")
    return f(x) if y is None else f(x, y)


def scal(alpha, x):
    lib().orc_scal(x.size, alpha, _p(x, np.float64))


def axpy(alpha, x, y):
    lib().orc_axpy(y.size, alpha, _p(x, np.float64), _p(y, np.float64))


def xpay(alpha, x, y):
    lib().orc_xpay(y.size, alpha, _p(x, np.float64), _p(y, np.float64))


def dot(v, w) -> float:
    return lib().orc_dot(v.size, _p(v, np.float64), _p(w, np.float64))


def scal_f32(alpha, x):
    lib().orc_scal_f32(x.size, alpha, _p(x, np.float32))


def axpy_f32(alpha, x, y):
    lib().orc_axpy_f32(y.size, alpha, _p(x, np.float32), _p(y, np.float32))


def xpay_f32(alpha, x, y):
    lib().orc_xpay_f32(y.size, alpha, _p(x, np.float32), _p(y, np.float32))


def dot_f32(v, w) -> float:
    return lib().orc_dot_f32(v.size, _p(v, np.float32), _p(w, np.float32))


# --------------------------------------------------------------------------------------------
# matrices
# --------------------------------------------------------------------------------------------
class Matrix:
    """Global CSR (rowptr = inclusive rects) or COO (row) arrays in the reference's layout."""

    def __init__(self, n_rows, n_cols, entry, col, rowptr=None, row=None):
        self.n_rows, self.n_cols = int(n_rows), int(n_cols)
        self.entry = np.ascontiguousarray(entry, dtype=np.float64)
        self.col = np.ascontiguousarray(col, dtype=np.int64)
        self.rowptr = None if rowptr is None else np.ascontiguousarray(rowptr, dtype=RECT_DTYPE)
        self.row = None if row is None else np.ascontiguousarray(row, dtype=np.int64)
        assert (self.rowptr is None) != (self.row is None)

    @property
    def nnz(self):
        return self.entry.size

    @property
    def is_csr(self):
        return self.rowptr is not None

    def to_coo(self) -> "Matrix":
        assert self.is_csr
        counts = self.rowptr["hi"] - self.rowptr["lo"] + 1
        row = np.repeat(np.arange(self.n_rows, dtype=np.int64), counts)
        return Matrix(self.n_rows, self.n_cols, self.entry, self.col, row=row)

    def to_scipy(self):
        import scipy.sparse as sp

        if self.is_csr:
            indptr = np.concatenate([self.rowptr["lo"], [self.rowptr["hi"][-1] + 1]])
            return sp.csr_matrix((self.entry, self.col, indptr), shape=(self.n_rows, self.n_cols))
        return sp.coo_matrix((self.entry, (self.row, self.col)), shape=(self.n_rows, self.n_cols)).tocsr()


def laplacian_1d_csr(n: int) -> Matrix:
    """csr_negative_laplacian_1d: kernel space [0, 3n-2), src/ExampleSystems.cpp (CSR creator)."""
    nnz = 3 * n - 2
    entry, col = np.empty(nnz), np.empty(nnz, dtype=np.int64)
    rowptr = np.empty(n, dtype=RECT_DTYPE)
    lib().orc_laplacian_1d_csr(0, nnz - 1, _p(entry), _p(col))
    lib().orc_laplacian_1d_rowptr(n, 0, n - 1, _p(rowptr))
    return Matrix(n, n, entry, col, rowptr=rowptr)


def laplacian_1d_coo(n: int) -> Matrix:
    nnz = 3 * n - 2
    entry, col, row = np.empty(nnz), np.empty(nnz, dtype=np.int64), np.empty(nnz, dtype=np.int64)
    lib().orc_laplacian_1d_coo(0, nnz - 1, _p(entry), _p(row), _p(col))
    return Matrix(n, n, entry, col, row=row)


def benchmark_stencil(dim_flag: int):
    """(offsets, values) of test/BenchmarkStencil.cpp:33-131 for -dim 1..4 (4 = 3-D 27-point)."""
    if dim_flag == 1:
        off = [(0,), (-1,), (1,)]
        val = [2.0, -1.0, -1.0]
    elif dim_flag == 2:
        off = [(0, 0), (-1, 0), (1, 0), (0, -1), (0, 1)]
        val = [4.0, -1.0, -1.0, -1.0, -1.0]
    elif dim_flag == 3:
        off = [(0, 0, 0), (-1, 0, 0), (1, 0, 0), (0, -1, 0), (0, 1, 0), (0, 0, -1), (0, 0, 1)]
        val = [6.0] + [-1.0] * 6
    elif dim_flag == 4:
        off, val = [(0, 0, 0)], [88.0 / 26.0]
        for o in [(-1, 0, 0), (1, 0, 0), (0, -1, 0), (0, 1, 0), (0, 0, -1), (0, 0, 1)]:
            off.append(o)
            val.append(-6.0 / 26.0)
        for o in [(-1, -1, 0), (-1, 1, 0), (1, -1, 0), (1, 1, 0), (-1, 0, -1), (-1, 0, 1),
                  (1, 0, -1), (1, 0, 1), (0, -1, -1), (0, -1, 1), (0, 1, -1), (0, 1, 1)]:
            off.append(o)
            val.append(-3.0 / 26.0)
        for o in [(-1, -1, -1), (-1, -1, 1), (-1, 1, -1), (-1, 1, 1), (1, -1, -1), (1, -1, 1),
                  (1, 1, -1), (1, 1, 1)]:
            off.append(o)
            val.append(-2.0 / 26.0)
    else:
        raise ValueError("INVALID DIM")
    return np.array(off, dtype=np.int64), np.array(val, dtype=np.float64)


def stencil_size(shape, offsets) -> int:
    dim = len(shape)
    lo = np.zeros(dim, dtype=np.int64)
    hi = np.array(shape, dtype=np.int64) - 1
    offsets = np.ascontiguousarray(offsets, dtype=np.int64).reshape(-1, dim)
    return lib().orc_stencil_size(dim, _p(lo), _p(hi), len(offsets), _p(offsets))


def stencil_csr(shape, offsets, values, order: int = 0, k_range=None, r_range=None) -> Matrix:
    """create_linearized_csr_stencil_matrix, src/StencilGenerator.hpp:533-643."""
    dim = len(shape)
    lo = np.zeros(dim, dtype=np.int64)
    hi = np.array(shape, dtype=np.int64) - 1
    offsets = np.ascontiguousarray(offsets, dtype=np.int64).reshape(-1, dim)
    values = np.ascontiguousarray(values, dtype=np.float64)
    n = int(np.prod(shape))
    nnz = stencil_size(shape, offsets)
    entry, col = np.zeros(nnz), np.zeros(nnz, dtype=np.int64)
    rowptr = np.zeros(n, dtype=RECT_DTYPE)
    k_lo, k_hi = k_range if k_range else (0, nnz - 1)
    r_lo, r_hi = r_range if r_range else (0, n - 1)
    lib().orc_fill_linearized_csr_stencil(dim, _p(lo), _p(hi), len(offsets), _p(offsets), _p(values),
                                          order, k_lo, k_hi, r_lo, r_hi, _p(entry), _p(col), _p(rowptr))
    return Matrix(n, n, entry, col, rowptr=rowptr)


def stencil_coo(shape, offsets, values, order: int = 0) -> Matrix:
    dim = len(shape)
    lo = np.zeros(dim, dtype=np.int64)
    hi = np.array(shape, dtype=np.int64) - 1
    offsets = np.ascontiguousarray(offsets, dtype=np.int64).reshape(-1, dim)
    values = np.ascontiguousarray(values, dtype=np.float64)
    n = int(np.prod(shape))
    nnz = stencil_size(shape, offsets)
    entry, col, row = np.zeros(nnz), np.zeros(nnz, dtype=np.int64), np.zeros(nnz, dtype=np.int64)
    lib().orc_fill_linearized_coo_stencil(dim, _p(lo), _p(hi), len(offsets), _p(offsets), _p(values),
                                          order, 0, nnz - 1, _p(entry), _p(row), _p(col))
    return Matrix(n, n, entry, col, row=row)


def csr_matvec(m: Matrix, x, y, k=None, r=None, cols=None, literal=False):
    """One CSRMatvecTask point task: accumulates into y (caller zero-fills, like the planner)."""
    k_lo, k_hi = k if k else (0, m.nnz - 1)
    r_lo, r_hi = r if r else (0, m.n_rows - 1)
    c_lo, c_hi = cols if cols else (0, m.n_cols - 1)
    args = (k_lo, k_hi, r_lo, r_hi, c_lo, c_hi, _p(m.entry), _p(m.col), _p(m.rowptr), _p(x, np.float64),
            _p(y, np.float64))
    if literal:
        lib().orc_csr_matvec_literal(*args)
    else:
        rc = lib().orc_csr_matvec(*args)
        assert rc == 0


def coo_matvec(m: Matrix, x, y, k=None, r=None, cols=None):
    k_lo, k_hi = k if k else (0, m.nnz - 1)
    r_lo, r_hi = r if r else (0, m.n_rows - 1)
    c_lo, c_hi = cols if cols else (0, m.n_cols - 1)
    lib().orc_coo_matvec(k_lo, k_hi, r_lo, r_hi, c_lo, c_hi, _p(m.entry), _p(m.row), _p(m.col),
                         _p(x, np.float64), _p(y, np.float64))


def rmatvec(m: Matrix, x, y):
    """y += A^T x over the whole matrix (CSRRmatvecTask / COORmatvecTask: no reference body, see lsk_oracle.c)."""
    if m.is_csr:
        lib().orc_csr_rmatvec(0, m.n_rows - 1, 0, m.n_cols - 1, _p(m.entry), _p(m.col), _p(m.rowptr), _p(x, np.float64), _p(y, np.float64))
    else:
        lib().orc_coo_rmatvec(0, m.nnz - 1, 0, m.n_rows - 1, 0, m.n_cols - 1, _p(m.entry), _p(m.row), _p(m.col), _p(x, np.float64),
                              _p(y, np.float64))


# --------------------------------------------------------------------------------------------
# partitions
# --------------------------------------------------------------------------------------------
def equal_partition(n: int, pieces: int):
    lo, hi = np.empty(pieces, dtype=np.int64), np.empty(pieces, dtype=np.int64)
    lib().orc_equal_partition(n, pieces, _p(lo), _p(hi))
    return lo, hi


def image_range(m: Matrix, r_lo, r_hi):
    f = np.zeros(m.nnz, dtype=np.uint8)
    lib().orc_image_range(_p(m.rowptr), r_lo, r_hi, m.nnz, _p(f))
    return f


def image(field, kernel_flags, n):
    out = np.zeros(n, dtype=np.uint8)
    lib().orc_image(_p(field, np.int64), _p(kernel_flags, np.uint8), field.size, n, _p(out))
    return out


def preimage(field, lo, hi):
    f = np.zeros(field.size, dtype=np.uint8)
    lib().orc_preimage(_p(field, np.int64), field.size, lo, hi, _p(f))
    return f


def preimage_range(m: Matrix, kernel_flags):
    f = np.zeros(m.n_rows, dtype=np.uint8)
    lib().orc_preimage_range(_p(m.rowptr), m.n_rows, _p(kernel_flags, np.uint8), _p(f))
    return f


def shard(point, volume, total_shards) -> int:
    return lib().orc_shard(point, volume, total_shards)


# --------------------------------------------------------------------------------------------
# planner + solvers
# --------------------------------------------------------------------------------------------
class Planner:
    """SquarePlanner (src/SquarePlanner.hpp) on the CPU: vector ids 0 SOL, 1 RHS, 2.. workspace."""

    def __init__(self, sizes, pieces):
        self.sizes = np.ascontiguousarray(sizes, dtype=np.int64)
        self.pieces = np.ascontiguousarray(pieces, dtype=np.int32)
        self.h = lib().orc_planner_create(len(self.sizes), _p(self.sizes), _p(self.pieces))
        self._keep = []

    def close(self):
        if self.h:
            lib().orc_planner_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def add_matrix(self, m: Matrix, domain_idx=0, range_idx=0) -> int:
        self._keep.append(m)
        idx = lib().orc_planner_add_matrix(self.h, domain_idx, range_idx, m.nnz, _p(m.entry), _p(m.col),
                                           _p(m.rowptr) if m.is_csr else None,
                                           None if m.is_csr else _p(m.row))
        assert idx >= 0, "kernel piece is not contiguous"
        return idx

    def vector(self, vec_idx, space=0):
        ptr = lib().orc_planner_vector(self.h, vec_idx, space)
        return np.ctypeslib.as_array(ptr, shape=(int(self.sizes[space]),))

    def _bounds(self, fn, a, b):
        lo, hi = C.c_int64(), C.c_int64()
        fn(self.h, a, b, C.byref(lo), C.byref(hi))
        return lo.value, hi.value

    def piece_bounds(self, space, piece):
        return self._bounds(lib().orc_planner_piece_bounds, space, piece)

    def kernel_bounds(self, matrix, piece):
        return self._bounds(lib().orc_planner_kernel_bounds, matrix, piece)

    def ghost_bounds(self, matrix, piece):
        return self._bounds(lib().orc_planner_ghost_bounds, matrix, piece)

    def use_literal_csr(self, flag=True):
        lib().orc_planner_use_literal_csr(self.h, int(flag))

    def allocate_workspace(self, n):
        lib().orc_planner_allocate_workspace(self.h, n)

    def fill(self, v, value):
        lib().orc_planner_fill(self.h, v, value)

    def copy(self, dst, src):
        lib().orc_planner_copy(self.h, dst, src)

    def scal(self, dst, terms):
        t = np.ascontiguousarray(terms, dtype=np.float64)
        lib().orc_planner_scal(self.h, dst, len(t), _p(t))

    def axpy(self, dst, terms, src):
        t = np.ascontiguousarray(terms, dtype=np.float64)
        lib().orc_planner_axpy(self.h, dst, len(t), _p(t), src)

    def xpay(self, dst, terms, src):
        t = np.ascontiguousarray(terms, dtype=np.float64)
        lib().orc_planner_xpay(self.h, dst, len(t), _p(t), src)

    def dot(self, v, w) -> float:
        return lib().orc_planner_dot(self.h, v, w)

    def matvec(self, dst, src):
        lib().orc_planner_matvec(self.h, dst, src)


class CGSolver:
    def __init__(self, planner: Planner):
        self.planner = planner
        self.h = lib().orc_cg_create(planner.h)

    def step(self):
        lib().orc_cg_step(self.h)

    @property
    def residual_norm_squared(self):
        n = lib().orc_cg_history(self.h, None, 0)
        out = np.empty(n)
        lib().orc_cg_history(self.h, _p(out), n)
        return out

    def __del__(self):
        if self.h:
            lib().orc_cg_destroy(self.h)
            self.h = None


class BiCGStabSolver:
    def __init__(self, planner: Planner):
        self.planner = planner
        self.h = lib().orc_bicgstab_create(planner.h)

    def step(self):
        lib().orc_bicgstab_step(self.h)

    def _hist(self, which):
        n = lib().orc_bicgstab_history(self.h, which, None, 0)
        out = np.empty(n)
        lib().orc_bicgstab_history(self.h, which, _p(out), n)
        return out

    rho = property(lambda self: self._hist(0))
    alpha = property(lambda self: self._hist(1))
    omega = property(lambda self: self._hist(2))

    def __del__(self):
        if self.h:
            lib().orc_bicgstab_destroy(self.h)
            self.h = None


class GMRESSolver:
    def __init__(self, planner: Planner, restart: int):
        self.planner, self.restart = planner, restart
        self.h = lib().orc_gmres_create(planner.h, restart)

    def step(self):
        lib().orc_gmres_step(self.h)

    @property
    def inner_products(self):
        out = np.empty((self.restart + 1, self.restart))
        lib().orc_gmres_hessenberg(self.h, _p(out))
        return out

    def __del__(self):
        if self.h:
            lib().orc_gmres_destroy(self.h)
            self.h = None
