"""ctypes binding of the C ABI declared in include/lsk.h (and include/lsk_solvers.h).

The CUDA library is the product: if liblsk.so is missing this module raises -- there is no
Python/CPU fallback path anywhere in the package.
"""
from __future__ import annotations

import ctypes as C
import re
from pathlib import Path

from .build import INCLUDE, LIB_PATH

i64, dbl, flt, vp, ci, u64 = C.c_int64, C.c_double, C.c_float, C.c_void_p, C.c_int, C.c_uint64


class LskError(RuntimeError):
    def __init__(self, status: int, where: str):
        self.status = status
        msg = _lib.lsk_error_string(status).decode() if _lib is not None else "?"
        super().__init__(f"{where} failed with status {status}: {msg}")


_lib = None


def declared_symbols() -> list[str]:
    """Every function name declared in include/*.h (the contract the .so must export)."""
    names: list[str] = []
    for h in sorted(INCLUDE.glob("*.h")):
        text = re.sub(r"/\*.*?\*/", "", h.read_text(), flags=re.S)
        names += re.findall(r"\b(lsk_[a-z0-9_]+)\s*\(", text)
    seen, out = set(), []
    for n in names:
        if n not in seen:
            seen.add(n)
            out.append(n)
    return out


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        import os

        path = Path(os.environ.get("LSK_LIB_PATH", LIB_PATH))  # developer switch: alternative build of the same library
        if not path.exists():
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m legionsolvers_b200.build` "
                "(or __graft_entry__.build()). legionsolvers_b200 has no CPU fallback."
            )
        L = C.CDLL(str(path), mode=C.RTLD_GLOBAL)
        _declare(L)
        _lib = L
    return _lib


class HaloMove(C.Structure):  # lsk_halo_move
    _fields_ = [("peer", ci), ("reserved", ci), ("src", vp), ("n", i64), ("ll_send", vp), ("recv_dst", vp), ("recv_n", i64), ("ll_recv", vp)]


class Peers(C.Structure):  # lsk_peers
    _fields_ = [("rank", ci), ("nranks", ci), ("window", vp * 16)]


def check(status: int, where: str) -> None:
    if status != 0:
        raise LskError(status, where)


def _declare(L: C.CDLL) -> None:
    L.lsk_version.restype = ci
    L.lsk_error_string.argtypes = [ci]
    L.lsk_error_string.restype = C.c_char_p
    L.lsk_ctx_create.argtypes = [ci, C.POINTER(vp)]
    L.lsk_ctx_destroy.argtypes = [vp]
    L.lsk_ctx_device.argtypes = [vp]
    L.lsk_ctx_sm_count.argtypes = [vp]
    L.lsk_ctx_launch_count.argtypes = [vp]
    L.lsk_ctx_launch_count.restype = u64
    L.lsk_ctx_const_f64.argtypes = [vp, ci]
    L.lsk_ctx_const_f64.restype = vp
    for sfx in ("f64", "f32"):
        getattr(L, f"lsk_scal_{sfx}").argtypes = [vp, vp, i64, ci, vp, vp, vp, vp, vp]
        getattr(L, f"lsk_axpy_{sfx}").argtypes = [vp, vp, i64, ci, vp, vp, vp, vp, vp, vp]
        getattr(L, f"lsk_xpay_{sfx}").argtypes = [vp, vp, i64, ci, vp, vp, vp, vp, vp, vp]
        getattr(L, f"lsk_dot_{sfx}").argtypes = [vp, vp, i64, vp, vp, vp]
        getattr(L, f"lsk_scalar_op_{sfx}").argtypes = [vp, vp, ci, vp, vp, vp]
        getattr(L, f"lsk_csr_spmv_{sfx}").argtypes = [vp, vp, i64, i64, vp, vp, vp, i64, vp, vp, vp, vp, vp, ci]
        getattr(L, f"lsk_coo_spmv_{sfx}").argtypes = [vp, vp, i64, vp, vp, vp, vp, vp, i64, i64, i64, i64]
    L.lsk_fill_f64.argtypes = [vp, vp, i64, dbl, vp]
    L.lsk_fill_f32.argtypes = [vp, vp, i64, flt, vp]
    L.lsk_fill_dev_f64.argtypes = [vp, vp, i64, vp, vp]
    L.lsk_copy_f64.argtypes = [vp, vp, i64, vp, vp]
    L.lsk_scalar_append_f64.argtypes = [vp, vp, vp, vp, i64, vp, vp]
    L.lsk_csr_spmv_pick.argtypes = [i64, i64]
    L.lsk_cg_update_f64.argtypes = [vp, vp, i64, vp, vp, vp, vp, vp, vp, vp]
    L.lsk_axpy_dot_f64.argtypes = [vp, vp, i64, ci, vp, vp, vp, vp, vp, vp, vp, vp]
    L.lsk_dot2_f64.argtypes = [vp, vp, i64, vp, vp, vp, vp]
    L.lsk_bicg_p_update_f64.argtypes = [vp, vp, i64, vp, vp, vp, vp, vp, vp, vp]
    L.lsk_bicg_tail_f64.argtypes = [vp, vp, i64, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    L.lsk_cg_direction_supported.argtypes = [i64, vp, vp]
    L.lsk_cg_direction_f64.argtypes = [vp, vp, i64, vp, vp, vp, vp, C.POINTER(HaloMove), ci, vp, i64, vp]
    L.lsk_cg_tail_supported.argtypes = [vp, i64, vp, vp, vp, vp]
    L.lsk_cg_tail_stats.argtypes = [vp, vp, C.POINTER(u64)]
    L.lsk_cg_tail_f64.argtypes = [vp, vp, i64, vp, vp, vp, vp, vp, vp, vp, C.POINTER(HaloMove), ci, vp, i64, vp]
    L.lsk_halo_landing_bytes.argtypes = [i64]
    L.lsk_halo_landing_bytes.restype = C.c_size_t
    # optional groups are declared by the modules that own them (setup / solvers / comm)
    from . import _abi_ext

    _abi_ext.declare(L)
