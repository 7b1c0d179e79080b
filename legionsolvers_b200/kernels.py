"""Thin torch-tensor front end of the leaf-kernel C ABI (include/lsk.h).

Every function enqueues ONE launch on the current torch CUDA stream through liblsk.so and returns
without synchronising.  Tensors only provide device pointers; no torch op computes anything here.
Raises if the CUDA library or a CUDA device is missing (there is no fallback).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _abi

SPMV_AUTO, SPMV_STREAM, SPMV_VECTOR, SPMV_WARP, SPMV_LANES = 0, 1, 2, 3, 4
OP_NEG, OP_ADD, OP_SUB, OP_MUL, OP_DIV, OP_SQRT, OP_RSQRT, OP_DUMMY, OP_COPY = range(9)


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t) -> int | None:
    if t is None:
        return None
    if isinstance(t, int):
        return t
    assert t.is_cuda, "device tensor required"
    return t.data_ptr()


def _sfx(t: torch.Tensor) -> str:
    if t.dtype == torch.float64:
        return "f64"
    if t.dtype == torch.float32:
        return "f32"
    raise TypeError(f"unsupported entry type {t.dtype}")


def _terms(terms):
    terms = list(terms)
    assert len(terms) <= 4
    ptrs = [_ptr(t) for t in terms] + [None] * (4 - len(terms))
    return len(terms), ptrs


class Context:
    """Per-GPU lsk_ctx (replaces the reference's CUDALibraryContext)."""

    def __init__(self, device: int | None = None):
        if not torch.cuda.is_available():
            raise RuntimeError("legionsolvers_b200 needs a CUDA device (no CPU fallback)")
        self.device = torch.cuda.current_device() if device is None else int(device)
        h = C.c_void_p()
        _abi.check(_abi.lib().lsk_ctx_create(self.device, C.byref(h)), "lsk_ctx_create")
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            _abi.lib().lsk_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def sm_count(self) -> int:
        return _abi.lib().lsk_ctx_sm_count(self.h)

    @property
    def launch_count(self) -> int:
        return int(_abi.lib().lsk_ctx_launch_count(self.h))

    def const(self, which: int) -> int:
        """Device pointer of the constant 1.0 (0), -1.0 (1) or 0.0 (2)."""
        return _abi.lib().lsk_ctx_const_f64(self.h, which)

    # ---- BLAS-1 ---------------------------------------------------------------------------------
    def scal(self, terms, x):
        n, p = _terms(terms)
        f = getattr(_abi.lib(), f"lsk_scal_{_sfx(x)}")
        _abi.check(f(self.h, _stream(), x.numel(), n, *p, _ptr(x)), "lsk_scal")

    def axpy(self, terms, x, y):
        n, p = _terms(terms)
        f = getattr(_abi.lib(), f"lsk_axpy_{_sfx(y)}")
        _abi.check(f(self.h, _stream(), y.numel(), n, *p, _ptr(x), _ptr(y)), "lsk_axpy")

    def xpay(self, terms, x, y):
        n, p = _terms(terms)
        f = getattr(_abi.lib(), f"lsk_xpay_{_sfx(y)}")
        _abi.check(f(self.h, _stream(), y.numel(), n, *p, _ptr(x), _ptr(y)), "lsk_xpay")

    def dot(self, v, w, out):
        f = getattr(_abi.lib(), f"lsk_dot_{_sfx(v)}")
        _abi.check(f(self.h, _stream(), v.numel(), _ptr(v), _ptr(w), _ptr(out)), "lsk_dot")

    def fill(self, x, value):
        if isinstance(value, torch.Tensor):
            _abi.check(_abi.lib().lsk_fill_dev_f64(self.h, _stream(), x.numel(), _ptr(value), _ptr(x)), "lsk_fill_dev")
        else:
            f = getattr(_abi.lib(), f"lsk_fill_{_sfx(x)}")
            _abi.check(f(self.h, _stream(), x.numel(), value, _ptr(x)), "lsk_fill")

    def copy(self, src, dst):
        _abi.check(_abi.lib().lsk_copy_f64(self.h, _stream(), dst.numel(), _ptr(src), _ptr(dst)), "lsk_copy")

    def scalar_op(self, op, a, b, out):
        f = getattr(_abi.lib(), f"lsk_scalar_op_{_sfx(out)}")
        _abi.check(f(self.h, _stream(), op, _ptr(a), _ptr(b), _ptr(out)), "lsk_scalar_op")

    # ---- mat-vec -----------------------------------------------------------------------------------
    def csr_spmv(self, rows, nnz, entry, col, rowptr, k_base, x, x_lo, y, dot_w=None, dot_out=None,
                 dot_yy_out=None, variant=SPMV_AUTO):
        """entry/col/rowptr/y are tensors whose element 0 is the piece's first entry/row; `x` holds
        global columns [x_lo, x_lo + len(x)) and is passed shifted to global column 0."""
        sfx = _sfx(y)
        x_shifted = x.data_ptr() - x_lo * x.element_size()
        f = getattr(_abi.lib(), f"lsk_csr_spmv_{sfx}")
        _abi.check(f(self.h, _stream(), rows, nnz, _ptr(entry), _ptr(col), _ptr(rowptr), k_base, x_shifted,
                     _ptr(y), _ptr(dot_w), _ptr(dot_out), _ptr(dot_yy_out), variant), "lsk_csr_spmv")

    def coo_spmv(self, nnz, entry, row, col, x, x_lo, y, y_lo, row_bounds, col_bounds):
        sfx = _sfx(y)
        x_shifted = x.data_ptr() - x_lo * x.element_size()
        y_shifted = y.data_ptr() - y_lo * y.element_size()
        f = getattr(_abi.lib(), f"lsk_coo_spmv_{sfx}")
        _abi.check(f(self.h, _stream(), nnz, _ptr(entry), _ptr(row), _ptr(col), x_shifted, y_shifted,
                     row_bounds[0], row_bounds[1], col_bounds[0], col_bounds[1]), "lsk_coo_spmv")

    # ---- fused solver passes ---------------------------------------------------------------------------
    def cg_update(self, rr_old, pq, p, q, x, r, rr_new):
        _abi.check(_abi.lib().lsk_cg_update_f64(self.h, _stream(), x.numel(), _ptr(rr_old), _ptr(pq), _ptr(p),
                                                _ptr(q), _ptr(x), _ptr(r), _ptr(rr_new)), "lsk_cg_update")

    def cg_direction(self, rr_cur, rr_new, r, p, history=None, history_count=None):
        """history.push_back(rr_new); p = fma(rr_new/rr_cur, p, r); rr_cur <- rr_new  (src/CGSolver.hpp:53-54)."""
        if not _abi.lib().lsk_cg_direction_supported(p.numel(), _ptr(r), _ptr(p)):
            raise RuntimeError("lsk_cg_direction_f64: r and p are not 32-byte congruent")
        _abi.check(_abi.lib().lsk_cg_direction_f64(self.h, _stream(), p.numel(), _ptr(rr_cur), _ptr(rr_new), _ptr(r), _ptr(p), None, 0,
                                                   _ptr(history), history.numel() if history is not None else 0,
                                                   _ptr(history_count)), "lsk_cg_direction")

    def cg_tail(self, rr_cur, pq, rr_new, p, q, x, r, history=None, history_count=None, moves=None, nmoves=0):
        """cg_update followed by cg_direction in ONE launch (src/CGSolver.hpp:50-54), for L2-resident vectors."""
        if not _abi.lib().lsk_cg_tail_supported(self.h, p.numel(), _ptr(p), _ptr(q), _ptr(x), _ptr(r)):
            raise RuntimeError("lsk_cg_tail_f64: vectors not 32-byte congruent, or too large for the one-launch form")
        _abi.check(_abi.lib().lsk_cg_tail_f64(self.h, _stream(), p.numel(), _ptr(rr_cur), _ptr(pq), _ptr(rr_new), _ptr(p), _ptr(q), _ptr(x),
                                              _ptr(r), moves, nmoves, _ptr(history), history.numel() if history is not None else 0,
                                              _ptr(history_count)), "lsk_cg_tail")

    def axpy_dot(self, terms, x, y, w, out):
        n, p = _terms(terms)
        _abi.check(_abi.lib().lsk_axpy_dot_f64(self.h, _stream(), y.numel(), n, *p, _ptr(x), _ptr(y), _ptr(w),
                                               _ptr(out)), "lsk_axpy_dot")

    def dot2(self, v, w, out_vw, out_ww):
        _abi.check(_abi.lib().lsk_dot2_f64(self.h, _stream(), v.numel(), _ptr(v), _ptr(w), _ptr(out_vw),
                                           _ptr(out_ww)), "lsk_dot2")

    def bicg_p_update(self, rho_new, rho_old, alpha, omega, v, r, p):
        _abi.check(_abi.lib().lsk_bicg_p_update_f64(self.h, _stream(), p.numel(), _ptr(rho_new), _ptr(rho_old),
                                                    _ptr(alpha), _ptr(omega), _ptr(v), _ptr(r), _ptr(p)),
                   "lsk_bicg_p_update")

    def bicg_tail(self, alpha, ru, uu, p, u, rt, x, r, rho_next):
        _abi.check(_abi.lib().lsk_bicg_tail_f64(self.h, _stream(), x.numel(), _ptr(alpha), _ptr(ru), _ptr(uu),
                                                _ptr(p), _ptr(u), _ptr(rt), _ptr(x), _ptr(r), _ptr(rho_next)),
                   "lsk_bicg_tail")


def rect_tensor(rowptr_np, device="cuda") -> torch.Tensor:
    """Structured (lo, hi) numpy rect array -> int64 [rows, 2] device tensor (same bytes)."""
    import numpy as np

    flat = np.ascontiguousarray(rowptr_np).view(np.int64).reshape(-1, 2)
    return torch.from_numpy(flat.copy()).to(device)
