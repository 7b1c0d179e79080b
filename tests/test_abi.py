"""CPU-side checks of the drop-in boundary: liblsk.so loads and exports every symbol that
include/*.h declares, and fails loudly (no fallback) when there is no GPU."""
import ctypes as C

import pytest
import torch

from legionsolvers_b200 import _abi, build


def test_library_is_built_for_sm100a_only():
    path = build.build_library()
    assert path.exists()
    flags = " ".join(build.NVCC_FLAGS)
    assert "arch=compute_100a,code=sm_100a" in flags and "-lineinfo" in flags


def test_exports_every_declared_symbol():
    L = _abi.lib()
    names = _abi.declared_symbols()
    assert len(names) >= 30, names
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, f"declared in include/*.h but not exported by liblsk.so: {missing}"


def test_version_and_error_strings():
    L = _abi.lib()
    assert L.lsk_version() == 100
    assert L.lsk_error_string(0) == b"success"
    assert b"invalid" in L.lsk_error_string(-1)
    assert b"no CPU fallback" in L.lsk_error_string(-2)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_fails_loudly_without_a_gpu():
    L = _abi.lib()
    h = C.c_void_p()
    assert L.lsk_ctx_create(0, C.byref(h)) == -2  # LSK_E_NO_DEVICE
    assert not h.value
    from legionsolvers_b200.kernels import Context

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Context()


def test_null_context_is_rejected_not_crashing():
    L = _abi.lib()
    assert L.lsk_dot_f64(None, None, 4, None, None, None) == -1
    assert L.lsk_csr_spmv_f64(None, None, 1, 1, None, None, None, 0, None, None, None, None, None, 0) == -1
    assert L.lsk_csr_spmv_pick(100, 700) == 1   # 7 nnz/row -> stream variant
    assert L.lsk_csr_spmv_pick(100, 100000) == 3  # long rows -> warp per row
