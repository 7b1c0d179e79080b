"""CPU-side checks of the drop-in boundary: liblsk.so loads and exports every symbol that
include/*.h declares, and fails loudly (no fallback) when there is no GPU."""
import ctypes as C
from pathlib import Path

import pytest
import torch

from legionsolvers_b200 import _abi, build

ROOT = Path(__file__).resolve().parents[1]


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


def test_spmv_variant_choice_by_row_length_statistics():
    """AUTO: a thread per row up to 12 non-zeros per row, 2-8 lanes per row up to 96, a warp per row beyond."""
    L = _abi.lib()
    assert L.lsk_csr_spmv_pick(1 << 20, 5 << 20) == 1    # 2-D 5-point   -> STREAM (bit-exact)
    assert L.lsk_csr_spmv_pick(1 << 20, 7 << 20) == 1    # 3-D 7-point   -> STREAM
    assert L.lsk_csr_spmv_pick(1 << 20, 27 << 20) == 4   # 3-D 27-point  -> LANES
    assert L.lsk_csr_spmv_pick(1 << 20, 96 << 20) == 4
    assert L.lsk_csr_spmv_pick(1 << 20, 200 << 20) == 3  # long rows     -> WARP
    assert L.lsk_csr_spmv_pick(0, 0) == 1


def test_host_only_entry_points_of_the_cg_step():
    """Eligibility checks of the fused CG kernels and the landing-buffer arithmetic run on the host: no GPU needed."""
    L = _abi.lib()
    # the fused direction kernel streams r and p with TMA: they must be 32-byte congruent and hold one aligned pack
    assert L.lsk_cg_direction_supported(1000, 0x10000, 0x20000) == 1
    assert L.lsk_cg_direction_supported(1000, 0x10008, 0x20008) == 1
    assert L.lsk_cg_direction_supported(1000, 0x10008, 0x20010) == 0
    assert L.lsk_cg_direction_supported(3, 0x10008, 0x20008) == 0
    assert L.lsk_cg_direction_supported(0, 0x10000, 0x20000) == 0
    # calls that would touch the device are refused without a context
    assert L.lsk_cg_direction_f64(None, None, 8, None, None, None, None, None, 0, None, 0, None) == -1
    assert L.lsk_halo_exchange_f64(None, None, None, None, 0) == -1
    assert L.lsk_halo_reduce_f64(None, None, None, None, 0) == -1
    assert L.lsk_cg_tail_supported(None, 1000, 0x10000, 0x20000, 0x30000, 0x40000) == 0
    assert L.lsk_cg_tail_f64(None, None, 8, None, None, None, None, None, None, None, None, 0, None, 0, None) == -1
    assert L.lsk_cg_tail_stats(None, None, None) == -1
    # a landing buffer holds two exchanges of count + 1 (the token) packets of 16 bytes
    assert L.lsk_halo_landing_bytes(0) == 64                      # each half rounded up to 32 bytes
    assert L.lsk_halo_landing_bytes(65536) == 2 * (65537 * 16 + 16)
    assert L.lsk_halo_landing_bytes(65535) == 2 * 65536 * 16
    assert L.lsk_halo_landing_bytes(-1) == 0
    assert C.sizeof(_abi.HaloMove) == 56


# ---- integration/: the reference-side binding is real code, not prose ----------------------------------------------------
def _compile_and_run(src: str, tmp_path, extra=()):
    import subprocess

    c = tmp_path / "t.c"
    c.write_text(src)
    exe = tmp_path / "t"
    subprocess.check_call(["gcc", "-std=c11", "-I", str(ROOT / "integration"), "-I", str(ROOT / "include"), *extra, str(c), "-o", str(exe)])
    return subprocess.check_output([str(exe)]).decode().split()


def test_task_ids_known_answers(tmp_path):
    """integration/lsk_task_ids.h restates src/TaskIDs.hpp:17-53 + src/TaskBaseClasses.hpp:61-103,196-209,262-285; the literal ids
    are the ones a build of the reference prints at registration."""
    src = r'''
#include <stdio.h>
#include "lsk_task_ids.h"
int main(void) {
    printf("%d %d %d %d ", LSK_TID_F64_1D_S64(LSK_BLOCK_SCAL), LSK_TID_F64_1D_S64(LSK_BLOCK_AXPY), LSK_TID_F64_1D_S64(LSK_BLOCK_XPAY), LSK_TID_F64_1D_S64(LSK_BLOCK_DOT));
    printf("%d %d %d %d ", LSK_TID_MATVEC_F64_1D_S64(LSK_BLOCK_COO_MATVEC), LSK_TID_MATVEC_F64_1D_S64(LSK_BLOCK_COO_RMATVEC),
           LSK_TID_MATVEC_F64_1D_S64(LSK_BLOCK_CSR_MATVEC), LSK_TID_MATVEC_F64_1D_S64(LSK_BLOCK_CSR_RMATVEC));
    for (int b = LSK_BLOCK_PRINT_SCALAR; b <= LSK_BLOCK_DUMMY; ++b) printf("%d ", lsk_task_id_t(b, LSK_ENTRY_F64));
    printf("%d %d %d ", LSK_LOAD_CUDA_LIBS_TASK_ID, LSK_TASK_BLOCK_SIZE,
           lsk_task_id_tdi(LSK_BLOCK_SCAL, LSK_ENTRY_F32, 1, LSK_INDEX_S64));
    return 0;
}'''
    got = [int(v) for v in _compile_and_run(src, tmp_path)]
    assert got[:4] == [557044, 562228, 567412, 572596]                       # Scal, Axpy, Xpay, Dot
    assert got[4:8] == [577888, 583072, 593440, 598624]                      # COOMatvec, COORmatvec, CSRMatvec, CSRRmatvec
    assert got[8:17] == [500002, 505186, 510370, 515554, 520738, 525922, 531106, 536290, 541474]  # Print .. Dummy (fp64)
    assert got[17:] == [500000, 5184, 557043]                                # LoadCUDALibs, block size, fp32 = fp64 - 1


def test_patched_cuda_task_bodies_type_check():
    """integration/cuda_task_bodies.cpp -- the six cuda_task_body bodies of INTEGRATION.md as code -- compiles against the
    <= 150-line stub of the Legion types they touch, with lsk.h's real prototypes."""
    import subprocess

    stub = (ROOT / "integration" / "legion_stub.h").read_text().splitlines()
    assert len(stub) <= 150
    proc = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-Wextra", "-Werror", "-I", str(ROOT / "include"), "-I", str(ROOT / "integration"),
                           str(ROOT / "integration" / "cuda_task_bodies.cpp")], capture_output=True, text=True)
    assert proc.returncode == 0, proc.stderr
    body = (ROOT / "integration" / "cuda_task_bodies.cpp").read_text()
    assert "static_assert(sizeof(Legion::Rect<1, long long>) == sizeof(lsk_rect)" in body
    for fn in ("lsk_scal_f64", "lsk_axpy_f64", "lsk_xpay_f64", "lsk_dot_f64", "lsk_csr_spmv_f64", "lsk_coo_spmv_f64", "lsk_csr_rspmv_f64"):
        assert fn in body
