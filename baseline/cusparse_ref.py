"""ctypes driver of baseline/cusparse_ref.cu: the reference's GPU call sequence through cuSPARSE / cuBLAS, timed by
bench.py next to the hand-written kernels (`vs_cusparse` on the main line, `--impl cusparse` as a line of its own).
COMPARATOR ONLY: nothing in legionsolvers_b200 imports this package."""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent
SRC = HERE / "cusparse_ref.cu"
LIB = HERE / "_build" / "libcusparse_ref.so"

_lib = None


def build(force: bool = False) -> Path:
    if LIB.exists() and not force and LIB.stat().st_mtime >= SRC.stat().st_mtime:
        return LIB
    LIB.parent.mkdir(exist_ok=True)
    nvcc = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    host_cxx = "/usr/bin/g++" if Path("/usr/bin/g++").exists() else "g++"
    env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-ccbin", host_cxx, "-Xcompiler", "-fPIC", "-shared",
           "-o", str(LIB), str(SRC), "-lcusparse", "-lcublas", "-lcudart"]
    proc = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed for the cuSPARSE comparator:\n" + proc.stdout + proc.stderr)
    return LIB


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not LIB.exists():
            raise RuntimeError(f"{LIB} is missing: build it with __graft_entry__.build() (comparator only)")
        L = C.CDLL(str(LIB))
        vp, i64, dbl, ci = C.c_void_p, C.c_int64, C.c_double, C.c_int
        L.ref_create.argtypes = [C.POINTER(vp)]
        L.ref_destroy.argtypes = [vp]
        L.ref_csr_matvec.argtypes = [vp, vp, i64, i64, i64, vp, vp, vp, vp, vp]
        L.ref_dot.argtypes = [vp, vp, i64, vp, vp, C.POINTER(dbl)]
        L.ref_axpy.argtypes = [vp, vp, i64, dbl, vp, vp]
        L.ref_xpay.argtypes = [vp, vp, i64, dbl, vp, vp]
        L.ref_cg_steps.argtypes = [vp, vp, ci, i64, i64, vp, vp, vp, vp, vp, vp, vp, C.POINTER(dbl)]
        _lib = L
    return _lib


def _check(rc, what):
    if rc != 0:
        raise RuntimeError(f"cusparse_ref.{what} failed ({rc})")


def measure(mat, n, nnz, stream, solver="cg", spmv_reps=20, cg_iters=10, ours_y=None):
    """One GPU, the whole matrix in one piece.  Returns the library path's SpMV time (full per-call sequence) and, for CG,
    the time of a reference CG iteration; plus the relative difference of the SpMV result to `ours_y` if given."""
    import torch

    L = lib()
    h = C.c_void_p()
    _check(L.ref_create(C.byref(h)), "ref_create")
    try:
        e_ptr, c_ptr, rp_ptr = mat.device_fields()
        g = torch.Generator(device="cuda").manual_seed(7)
        x = torch.rand(n, dtype=torch.float64, device="cuda", generator=g) - 0.5
        y = torch.zeros(n, dtype=torch.float64, device="cuda")

        def spmv():
            _check(L.ref_csr_matvec(h, stream, n, n, nnz, rp_ptr, c_ptr, e_ptr, x.data_ptr(), y.data_ptr()), "ref_csr_matvec")

        for _ in range(3):
            spmv()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(spmv_reps):
            spmv()
        b.record()
        torch.cuda.synchronize()
        out = {"spmv_ms": a.elapsed_time(b) / spmv_reps,
               "spmv_sequence": "per call: cudaMallocAsync indptr, convertGlobalRowptrToLocalIndPtr, cusparseCreateCsr(int64), 2 x CreateDnVec, "
                                "SpMV_bufferSize, workspace, cusparseSpMV(ALG_DEFAULT, beta = 0), destroys (src/CSRMatrixTasks.cu:88-155)"}
        out["spmv_gbs"] = (16 * nnz + 32 * n) / (out["spmv_ms"] * 1e-3) / 1e9
        if ours_y is not None:
            ours_y(x, y, out)
        if solver == "cg":
            sol = torch.zeros(n, dtype=torch.float64, device="cuda")
            p = torch.ones(n, dtype=torch.float64, device="cuda")
            r = torch.ones(n, dtype=torch.float64, device="cuda")
            q = torch.zeros(n, dtype=torch.float64, device="cuda")
            rr = C.c_double(float(n))
            _check(L.ref_cg_steps(h, stream, 2, n, nnz, rp_ptr, c_ptr, e_ptr, sol.data_ptr(), p.data_ptr(), q.data_ptr(), r.data_ptr(), C.byref(rr)),
                   "ref_cg_steps")
            torch.cuda.synchronize()
            a.record()
            _check(L.ref_cg_steps(h, stream, cg_iters, n, nnz, rp_ptr, c_ptr, e_ptr, sol.data_ptr(), p.data_ptr(), q.data_ptr(), r.data_ptr(), C.byref(rr)),
                   "ref_cg_steps")
            b.record()
            torch.cuda.synchronize()
            out["iteration_ms"] = a.elapsed_time(b) / cg_iters
            out["iteration_sequence"] = ("CGSolver::step through the reference's GPU variants: zero fill, cuSPARSE mat-vec, cublasDdot (host result + stream "
                                         "sync), 2 x cublasDaxpy, cublasDdot, xpay_kernel (src/CGSolver.hpp:46-55)")
            out["rr_after"] = rr.value
        return out
    finally:
        L.ref_destroy(h)


def compare(rt, mat, n, nnz, stream, ours_spmv_ms, ours_iteration_ms, solver):
    """`vs_cusparse` of bench.py's main line: ratios > 1 mean the hand-written path is faster."""
    import torch

    from legionsolvers_b200 import _abi

    def check_against_ours(x, y_ref, out):
        e_ptr, c_ptr, rp_ptr = mat.device_fields()
        y = torch.zeros(n, dtype=torch.float64, device="cuda")
        _abi.check(_abi.lib().lsk_csr_spmv_f64(rt.ctx, stream, n, nnz, e_ptr, c_ptr, rp_ptr, 0, x.data_ptr(), y.data_ptr(), None, None, None, 0),
                   "lsk_csr_spmv_f64")
        torch.cuda.synchronize()
        out["result_max_rel_diff_to_ours"] = float(((y - y_ref).abs().max() / y_ref.abs().max()).item())

    m = measure(mat, n, nnz, stream, solver=solver, ours_y=check_against_ours)
    res = {"spmv_ratio": m["spmv_ms"] / ours_spmv_ms, "cusparse_spmv_ms": m["spmv_ms"], "cusparse_spmv_gbs": m["spmv_gbs"],
           "ours_spmv_ms": ours_spmv_ms, "result_max_rel_diff_to_ours": m.get("result_max_rel_diff_to_ours"),
           "library": "cuSPARSE / cuBLAS 12.9, the reference's call sequence (baseline/cusparse_ref.cu)", "spmv_sequence": m["spmv_sequence"]}
    if "iteration_ms" in m:
        res.update({"iteration_ratio": m["iteration_ms"] / ours_iteration_ms, "cusparse_iteration_ms": m["iteration_ms"],
                    "ours_iteration_ms": ours_iteration_ms, "iteration_sequence": m["iteration_sequence"]})
    else:
        res["iteration_ratio"] = None
    return res


def bench_line(args, WORKLOADS, shared_config, metric_name):
    """`bench.py --impl cusparse`: the reference's GPU path as a line of its own (1 GPU, CSR workloads, CG)."""
    import torch

    from legionsolvers_b200 import solvers as S

    kind, dim_flag, shape, default_solver, desc = WORKLOADS[args.workload]
    solver = args.solver or default_solver
    if kind != "stencil":
        return {"impl": "cusparse", "unavailable": "the comparator covers the CSR workloads"}
    if args.shape:
        shape = tuple(int(v) for v in args.shape.split(","))
    ts = torch.cuda.Stream()
    torch.cuda.set_stream(ts)
    rt = S.Runtime(device=0, stream=ts.cuda_stream)
    mat = S.CSRMatrix.stencil(rt, S.benchmark_stencil(dim_flag, *shape), 1)
    n, nnz = mat.rows, mat.nnz
    iters = max(1, args.steps) * args.iters_per_step if solver == "cg" else 0
    m = measure(mat, n, nnz, ts.cuda_stream, solver=solver, cg_iters=min(max(iters, 10), 200))
    value = 1e3 / m["iteration_ms"] if "iteration_ms" in m else None
    return {"impl": "cusparse", "metric": metric_name(solver), "value": value, "unit": "it/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": (m.get("iteration_ms") or 0.0) * args.iters_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": shared_config(args, n, nnz),
            "roofline": {"bound": "hbm", "kernel": "cusparseSpMV(CSR, int64, ALG_DEFAULT) + per-call set-up", "achieved": m["spmv_gbs"], "unit": "GB/s",
                         "ms_per_launch": m["spmv_ms"]},
            "note": "the reference's GPU leaf-task call sequence on one GPU (baseline/cusparse_ref.cu); Legion's own task overhead is not modelled"}
