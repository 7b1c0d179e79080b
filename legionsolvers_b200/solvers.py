"""Python mirror of the reference's planner-level API, a thin ctypes driver over the C++ host layer
(include/lsk_solvers.h): Runtime, PartitionedVector, CSRMatrix, COOMatrix, SquarePlanner, CGSolver,
BiCGStabSolver, GMRESSolver -- same names and argument meaning as src/*.hpp of the reference, so the
parity tests read like the reference's own tests.  Nothing here computes: every method is one call
into liblsk.so.  numpy arrays are host staging only.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _abi
from ._abi_ext import Stencil

RECT_DTYPE = np.dtype([("lo", np.int64), ("hi", np.int64)])  # Legion::Rect<1, long long>
SOLVER_CG, SOLVER_BICGSTAB, SOLVER_GMRES = 1, 2, 3


def _check(status: int, where: str) -> None:
    if status != 0:
        msg = _abi.lib().lsk_last_error().decode(errors="replace")
        raise RuntimeError(f"{where} failed ({status}): {msg}")


def _np_ptr(a, dtype):
    assert isinstance(a, np.ndarray) and a.flags["C_CONTIGUOUS"] and a.dtype == dtype, (a.dtype, dtype)
    return a.ctypes.data_as(C.c_void_p)


def make_stencil(shape, offsets, values, order: int = 0) -> Stencil:
    st = Stencil()
    st.dim, st.order, st.noff = len(shape), order, len(values)
    for d, s in enumerate(shape):
        st.shape[d] = int(s)
    for j, (o, v) in enumerate(zip(offsets, values)):
        for d in range(len(shape)):
            st.offsets[j][d] = int(o[d])
        st.values[j] = float(v)
    return st


def benchmark_stencil(dim_flag: int, nx: int, ny: int = 1, nz: int = 1) -> Stencil:
    """The matrices of test/BenchmarkStencil.cpp: -dim 1, 2, 3, 4 (= 3-D 27-point)."""
    st = Stencil()
    _check(_abi.lib().lsk_benchmark_stencil(dim_flag, nx, ny, nz, C.byref(st)), "lsk_benchmark_stencil")
    return st


def stencil_size(st: Stencil) -> int:
    return int(_abi.lib().lsk_stencil_size(C.byref(st)))


class Runtime:
    """One per process / GPU: stream, scalar arena, traces (CUDA graphs), NCCL communicator."""

    def __init__(self, device: int = 0, rank: int = 0, nranks: int = 1, stream: int | None = None):
        h = C.c_void_p()
        _check(_abi.lib().lsk_rt_create(device, rank, nranks, stream, C.byref(h)), "lsk_rt_create")
        self.h, self.rank, self.nranks, self.device = h, rank, nranks, device

    @staticmethod
    def unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        _check(_abi.lib().lsk_rt_unique_id(buf), "lsk_rt_unique_id")
        return buf.raw

    def comm_init(self, uid: bytes) -> None:
        assert len(uid) == 128
        _check(_abi.lib().lsk_rt_comm_init(self.h, C.create_string_buffer(uid, 128)), "lsk_rt_comm_init")

    @property
    def uses_peer_memory(self) -> bool:
        return _abi.lib().lsk_rt_uses_peer_memory(self.h) > 0

    @property
    def collectives(self) -> str:
        return {0: "nccl", 1: "peer-memory kernels", 2: "peer-memory, fused into producer kernels"}[
            _abi.lib().lsk_rt_uses_peer_memory(self.h)]

    def comm_stats(self) -> dict:
        """{ar_calls, ar_ns, halo_calls, halo_ns}: time spent inside the peer-memory collectives."""
        out = (C.c_uint64 * 4)()
        _check(_abi.lib().lsk_rt_comm_stats(self.h, out), "lsk_rt_comm_stats")
        return {"ar_calls": out[0], "ar_ns": out[1], "halo_calls": out[2], "halo_ns": out[3]}

    def comm_error(self) -> int:
        out = C.c_int()
        _check(_abi.lib().lsk_rt_comm_error(self.h, C.byref(out)), "lsk_rt_comm_error")
        return out.value

    @property
    def ctx(self) -> int:
        return _abi.lib().lsk_rt_ctx(self.h)

    @property
    def stream(self) -> int:
        return _abi.lib().lsk_rt_stream(self.h)

    def fence(self) -> None:
        _check(_abi.lib().lsk_rt_fence(self.h), "lsk_rt_fence")

    @property
    def kernel_launches(self) -> int:
        return int(_abi.lib().lsk_rt_kernel_launches(self.h))

    def begin_trace(self, trace_id: int) -> None:
        _check(_abi.lib().lsk_rt_begin_trace(self.h, trace_id), "lsk_rt_begin_trace")

    def end_trace(self, trace_id: int) -> None:
        _check(_abi.lib().lsk_rt_end_trace(self.h, trace_id), "lsk_rt_end_trace")

    def close(self) -> None:
        if getattr(self, "h", None):
            _abi.lib().lsk_rt_destroy(self.h)
            self.h = None


class PartitionedVector:
    def __init__(self, rt: Runtime, name: str, volume: int, pieces: int):
        h = C.c_void_p()
        _check(_abi.lib().lsk_vector_create(rt.h, name.encode(), volume, pieces, C.byref(h)), "lsk_vector_create")
        self.h, self.rt, self.volume, self.pieces = h, rt, volume, pieces

    def owned_range(self):
        lo, hi = C.c_int64(), C.c_int64()
        _check(_abi.lib().lsk_vector_owned_range(self.h, C.byref(lo), C.byref(hi)), "lsk_vector_owned_range")
        return lo.value, hi.value

    def constant_fill(self, value: float):
        _check(_abi.lib().lsk_vector_constant_fill(self.h, value), "constant_fill")

    def zero_fill(self):
        self.constant_fill(0.0)

    def assign(self, src: "PartitionedVector"):
        _check(_abi.lib().lsk_vector_assign(self.h, src.h), "operator=")

    def scal(self, alpha: float):
        _check(_abi.lib().lsk_vector_scal(self.h, alpha), "scal")

    def axpy(self, alpha: float, x: "PartitionedVector"):
        _check(_abi.lib().lsk_vector_axpy(self.h, alpha, x.h), "axpy")

    def xpay(self, alpha: float, x: "PartitionedVector"):
        _check(_abi.lib().lsk_vector_xpay(self.h, alpha, x.h), "xpay")

    def dot(self, w: "PartitionedVector") -> float:
        out = C.c_double()
        _check(_abi.lib().lsk_vector_dot(self.h, w.h, C.byref(out)), "dot")
        return out.value

    def from_numpy(self, a: np.ndarray):
        assert a.size == self.volume
        _check(_abi.lib().lsk_vector_copy_from_host(self.h, _np_ptr(a, np.float64)), "copy_from_host")

    def to_numpy(self) -> np.ndarray:
        """Global-size array; only this rank's owned rows are filled, the rest is NaN."""
        a = np.full(self.volume, np.nan)
        _check(_abi.lib().lsk_vector_copy_to_host(self.h, _np_ptr(a, np.float64)), "copy_to_host")
        return a

    def destroy(self):
        if getattr(self, "h", None):
            _abi.lib().lsk_vector_destroy(self.h)
            self.h = None


class _Matrix:
    def __init__(self, rt: Runtime, h):
        self.rt, self.h = rt, h
        info = np.zeros(8, dtype=np.int64)
        _check(_abi.lib().lsk_matrix_info(h, _np_ptr(info, np.int64)), "lsk_matrix_info")
        (self.rows, self.cols, self.nnz, self.slab_r_lo, self.slab_r_hi, self.slab_k_lo, self.slab_k_hi,
         self.is_csr) = (int(v) for v in info)

    def slab_to_numpy(self):
        nk = max(0, self.slab_k_hi - self.slab_k_lo + 1)
        nr = max(0, self.slab_r_hi - self.slab_r_lo + 1)
        entry, col = np.zeros(nk), np.zeros(nk, dtype=np.int64)
        third = np.zeros(nr, dtype=RECT_DTYPE) if self.is_csr else np.zeros(nk, dtype=np.int64)
        _check(_abi.lib().lsk_matrix_slab_to_host(self.h, _np_ptr(entry, np.float64), _np_ptr(col, np.int64),
                                                 third.ctypes.data_as(C.c_void_p)), "slab_to_host")
        return entry, col, third

    def device_fields(self):
        e, c, t = C.c_void_p(), C.c_void_p(), C.c_void_p()
        _check(_abi.lib().lsk_matrix_device_fields(self.h, C.byref(e), C.byref(c), C.byref(t)), "device_fields")
        return e.value, c.value, t.value

    def destroy(self):
        if getattr(self, "h", None):
            _abi.lib().lsk_matrix_destroy(self.h)
            self.h = None


# ---- utilities: Matrix Market files (the reference's planned I/O helper, README.md:90-99; host-only) -----------------------
class MMInfo(C.Structure):  # lsk_mm_info
    _fields_ = [("rows", C.c_int64), ("cols", C.c_int64), ("entries", C.c_int64), ("field", C.c_int), ("symmetry", C.c_int)]


def _mm_check(status: int, where: str) -> None:
    if status != 0:
        raise RuntimeError(f"{where} failed ({status}): {_abi.lib().lsk_mm_last_error().decode(errors='replace')}")


def read_matrix_market(path):
    """-> (rows, cols, entry, row, col): the EXPANDED matrix of a `coordinate` file as 0-based COO arrays, in file order."""
    info = MMInfo()
    _mm_check(_abi.lib().lsk_mm_read_info(str(path).encode(), C.byref(info)), "lsk_mm_read_info")
    cap = int(info.entries) * (1 if info.symmetry == 0 else 2)
    entry, row, col = np.zeros(cap, dtype=np.float64), np.zeros(cap, dtype=np.int64), np.zeros(cap, dtype=np.int64)
    nnz = C.c_int64(0)
    _mm_check(_abi.lib().lsk_mm_read_coo_f64(str(path).encode(), cap, _np_ptr(entry, np.float64), _np_ptr(row, np.int64), _np_ptr(col, np.int64),
                                             C.byref(nnz)), "lsk_mm_read_coo_f64")
    n = int(nnz.value)
    return int(info.rows), int(info.cols), entry[:n].copy(), row[:n].copy(), col[:n].copy()


def write_matrix_market(path, rows, cols, entry, row, col):
    e, r, c = (np.ascontiguousarray(entry, dtype=np.float64), np.ascontiguousarray(row, dtype=np.int64), np.ascontiguousarray(col, dtype=np.int64))
    _mm_check(_abi.lib().lsk_mm_write_coo_f64(str(path).encode(), rows, cols, e.size, _np_ptr(e, np.float64), _np_ptr(r, np.int64), _np_ptr(c, np.int64)),
              "lsk_mm_write_coo_f64")


def coo_to_csr(rows, entry, row, col):
    """COO arrays in any order -> (entry, col, rowptr) in the CSRMatrix field layout (inclusive rects, a row's entries in input order)."""
    e, r, c = (np.ascontiguousarray(entry, dtype=np.float64), np.ascontiguousarray(row, dtype=np.int64), np.ascontiguousarray(col, dtype=np.int64))
    eo, co, rp = np.zeros_like(e), np.zeros_like(c), np.zeros(rows, dtype=RECT_DTYPE)
    _mm_check(_abi.lib().lsk_coo_to_csr_f64(rows, e.size, _np_ptr(e, np.float64), _np_ptr(r, np.int64), _np_ptr(c, np.int64), _np_ptr(eo, np.float64),
                                            _np_ptr(co, np.int64), rp.ctypes.data_as(C.c_void_p)), "lsk_coo_to_csr_f64")
    return eo, co, rp


class CSRMatrix(_Matrix):
    @classmethod
    def from_host(cls, rt, rows, cols, entry, col, rowptr, r_range=None, k_range=None, nnz_global=None):
        """entry/col/rowptr are GLOBAL arrays; the slab [r_range] x [k_range] is uploaded."""
        r_lo, r_hi = r_range if r_range else (0, rows - 1)
        k_lo, k_hi = k_range if k_range else (0, entry.size - 1)
        e = np.ascontiguousarray(entry[k_lo:k_hi + 1], dtype=np.float64)
        c = np.ascontiguousarray(col[k_lo:k_hi + 1], dtype=np.int64)
        rp = np.ascontiguousarray(rowptr[r_lo:r_hi + 1], dtype=RECT_DTYPE)
        h = C.c_void_p()
        _check(_abi.lib().lsk_csr_create(rt.h, rows, cols, nnz_global if nnz_global is not None else entry.size,
                                        r_lo, r_hi, k_lo, k_hi, _np_ptr(e, np.float64), _np_ptr(c, np.int64),
                                        rp.ctypes.data_as(C.c_void_p), C.byref(h)), "lsk_csr_create")
        return cls(rt, h)

    @classmethod
    def from_matrix_market(cls, rt, path):
        """The whole matrix of a Matrix Market file on this rank (single-rank use; slabs: read_matrix_market + from_host)."""
        rows, cols, entry, row, col = read_matrix_market(path)
        e, c, rp = coo_to_csr(rows, entry, row, col)
        return cls.from_host(rt, rows, cols, e, c, rp)

    @classmethod
    def stencil(cls, rt, st: Stencil, pieces: int):
        """create_linearized_csr_stencil_matrix, filled on the GPU."""
        h = C.c_void_p()
        _check(_abi.lib().lsk_csr_create_stencil(rt.h, C.byref(st), pieces, C.byref(h)), "lsk_csr_create_stencil")
        return cls(rt, h)


class COOMatrix(_Matrix):
    @classmethod
    def from_matrix_market(cls, rt, path):
        rows, cols, entry, row, col = read_matrix_market(path)
        return cls.from_host(rt, rows, cols, entry, row, col)

    @classmethod
    def from_host(cls, rt, rows, cols, entry, row, col, k_range=None, nnz_global=None):
        k_lo, k_hi = k_range if k_range else (0, entry.size - 1)
        e = np.ascontiguousarray(entry[k_lo:k_hi + 1], dtype=np.float64)
        r = np.ascontiguousarray(row[k_lo:k_hi + 1], dtype=np.int64)
        c = np.ascontiguousarray(col[k_lo:k_hi + 1], dtype=np.int64)
        h = C.c_void_p()
        _check(_abi.lib().lsk_coo_create(rt.h, rows, cols, nnz_global if nnz_global is not None else entry.size,
                                        k_lo, k_hi, _np_ptr(e, np.float64), _np_ptr(r, np.int64), _np_ptr(c, np.int64),
                                        C.byref(h)), "lsk_coo_create")
        return cls(rt, h)


class SquarePlanner:
    SOL, RHS = 0, 1

    def __init__(self, rt: Runtime):
        h = C.c_void_p()
        _check(_abi.lib().lsk_planner_create(rt.h, C.byref(h)), "lsk_planner_create")
        self.h, self.rt = h, rt
        self._keep = []  # the planner borrows vectors and matrices: keep them alive

    def add_sol_vector(self, v: PartitionedVector):
        self._keep.append(v)
        _check(_abi.lib().lsk_planner_add_sol_vector(self.h, v.h), "add_sol_vector")

    def add_rhs_vector(self, v: PartitionedVector):
        self._keep.append(v)
        _check(_abi.lib().lsk_planner_add_rhs_vector(self.h, v.h), "add_rhs_vector")

    def add_row_partitioned_matrix(self, m: _Matrix, domain_index: int, range_index: int):
        self._keep.append(m)
        _check(_abi.lib().lsk_planner_add_row_partitioned_matrix(self.h, m.h, domain_index, range_index),
               "add_row_partitioned_matrix")

    def allocate_workspace(self, n: int):
        _check(_abi.lib().lsk_planner_allocate_workspace(self.h, n), "allocate_workspace")

    def _bounds(self, which, index, color):
        lo, hi = C.c_int64(), C.c_int64()
        _check(_abi.lib().lsk_planner_partition_bounds(self.h, which, index, color, C.byref(lo), C.byref(hi)),
               "partition_bounds")
        return lo.value, hi.value

    def range_bounds(self, space, color):
        return self._bounds(0, space, color)

    def kernel_bounds(self, block, color):
        return self._bounds(1, block, color)

    def ghost_bounds(self, block, color):
        return self._bounds(2, block, color)

    def local_colors(self, space=0):
        a, b = C.c_int(), C.c_int()
        _check(_abi.lib().lsk_planner_local_colors(self.h, space, C.byref(a), C.byref(b)), "local_colors")
        return a.value, b.value

    @property
    def halo_bytes_per_matvec(self) -> int:
        return int(_abi.lib().lsk_planner_halo_bytes_per_matvec(self.h))

    def zero_fill(self, v):
        _check(_abi.lib().lsk_planner_zero_fill(self.h, v), "zero_fill")

    def copy(self, dst, src):
        _check(_abi.lib().lsk_planner_copy(self.h, dst, src), "copy")

    def scal(self, dst, alpha):
        _check(_abi.lib().lsk_planner_scal(self.h, dst, alpha), "scal")

    def axpy(self, dst, alpha, src):
        _check(_abi.lib().lsk_planner_axpy(self.h, dst, alpha, src), "axpy")

    def xpay(self, dst, alpha, src):
        _check(_abi.lib().lsk_planner_xpay(self.h, dst, alpha, src), "xpay")

    def dot(self, v, w) -> float:
        out = C.c_double()
        _check(_abi.lib().lsk_planner_dot(self.h, v, w, C.byref(out)), "dot")
        return out.value

    def matvec(self, dst, src):
        _check(_abi.lib().lsk_planner_matvec(self.h, dst, src), "matvec")

    def rmatvec(self, dst, src):
        """dst = A^T src (CSRRmatvecTask / COORmatvecTask)."""
        _check(_abi.lib().lsk_planner_rmatvec(self.h, dst, src), "rmatvec")

    def matvec_dot(self, dst, src, w, want_yy=False):
        a, b = C.c_double(), C.c_double()
        _check(_abi.lib().lsk_planner_matvec_dot(self.h, dst, src, w, C.byref(a), C.byref(b) if want_yy else None),
               "matvec_dot")
        return (a.value, b.value) if want_yy else a.value

    def vector_to_numpy(self, vec, space, volume) -> np.ndarray:
        a = np.full(volume, np.nan)
        _check(_abi.lib().lsk_planner_vector_to_host(self.h, vec, space, _np_ptr(a, np.float64)), "vector_to_host")
        return a

    def vector_from_numpy(self, vec, space, a: np.ndarray):
        _check(_abi.lib().lsk_planner_vector_from_host(self.h, vec, space, _np_ptr(a, np.float64)), "vector_from_host")

    def vector_to_async(self, vec, space, global_ptr: int, stream: int | None = None):
        """Owned rows -> `global_ptr + 8 * own_lo` (pinned host or device memory), asynchronously on `stream`
        (a cudaStream_t handle; None = the runtime's stream).  No synchronisation: order foreign streams with events."""
        _check(_abi.lib().lsk_planner_vector_to_async(self.h, vec, space, global_ptr, stream), "vector_to_async")

    def vector_from_async(self, vec, space, global_ptr: int, stream: int | None = None):
        _check(_abi.lib().lsk_planner_vector_from_async(self.h, vec, space, global_ptr, stream), "vector_from_async")

    def destroy(self):
        if getattr(self, "h", None):
            _abi.lib().lsk_planner_destroy(self.h)
            self.h = None


class _Solver:
    KIND = 0

    def __init__(self, planner: SquarePlanner, restart: int = 0, fused: bool = True):
        h = C.c_void_p()
        _check(_abi.lib().lsk_solver_create(planner.h, self.KIND, restart, int(fused), C.byref(h)), "lsk_solver_create")
        self.h, self.planner, self.restart = h, planner, restart

    def step(self):
        _check(_abi.lib().lsk_solver_step(self.h), "step")

    def _history(self, which: int) -> np.ndarray:
        n = C.c_int64()
        _check(_abi.lib().lsk_solver_history(self.h, which, None, 0, C.byref(n)), "history")
        out = np.zeros(n.value)
        _check(_abi.lib().lsk_solver_history(self.h, which, _np_ptr(out, np.float64), n.value, C.byref(n)), "history")
        return out

    def reset(self):
        """Start a new solve from the current RHS (SOL taken as 0, like the constructor)."""
        _check(_abi.lib().lsk_solver_reset(self.h), "reset")

    def history_copy_async(self, which: int, dst_ptr: int, n: int, stream: int | None = None):
        """First n entries of history `which` -> dst_ptr (device or pinned host), asynchronously, no synchronisation."""
        _check(_abi.lib().lsk_solver_history_copy_async(self.h, which, dst_ptr, n, stream), "history_copy_async")

    def destroy(self):
        if getattr(self, "h", None):
            _abi.lib().lsk_solver_destroy(self.h)
            self.h = None


class CGSolver(_Solver):
    KIND = SOLVER_CG

    def __init__(self, planner, fused=True):
        super().__init__(planner, 0, int(bool(fused)))

    @property
    def residual_norm_squared(self) -> np.ndarray:
        return self._history(0)


class BiCGStabSolver(_Solver):
    KIND = SOLVER_BICGSTAB

    def __init__(self, planner, fused=True):
        super().__init__(planner, 0, fused)

    rho = property(lambda self: self._history(0))
    alpha = property(lambda self: self._history(1))
    omega = property(lambda self: self._history(2))


class GMRESSolver(_Solver):
    KIND = SOLVER_GMRES

    def __init__(self, planner, restart, fused=True, real_update=False):
        """real_update=False: the reference's placeholder update (DummyTask); True: Givens least squares + SOL += V y."""
        super().__init__(planner, restart, fused)
        if real_update:
            _check(_abi.lib().lsk_solver_set_option(self.h, 1, 1), "set_option")

    @property
    def residual_norm(self) -> np.ndarray:
        """|| b - A x || after each cycle (real update only)."""
        return self._history(1)

    @property
    def inner_products(self) -> np.ndarray:
        return self._history(0).reshape(self.restart + 1, self.restart)
