"""Developer probe (not part of the product or the bench contract): time lsk_csr_spmv_f64 on one of the bench workloads
with the implementation / knobs taken from the environment (LSK_SPMV_IMPL=ws|tma|pipe, LSK_SPMV_DYN, LSK_WS_CTAS, ...),
one process per setting because the switches are read once.  The matrix comes from the product's own GPU generator; the
result is compared bit for bit (thread-per-row variants) with the round-1 kernel's output saved by a previous run of this
probe when `--check FILE` is given.

    python tools/probe_spmv_ab.py c3 [--ndot 1] [--reps 50] [--shape nx,ny,nz] [--save FILE | --check FILE]
"""
import argparse
import json
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from legionsolvers_b200 import _abi  # noqa: E402
from legionsolvers_b200 import solvers as S  # noqa: E402

WORK = {"c3": (3, (256, 256, 256)), "c2": (2, (8192, 8192, 1)), "c4": (4, (192, 192, 192)), "c1": (2, (256, 256, 1)),
        "slab8": (3, (32, 256, 256)), "slab4": (3, (64, 256, 256))}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("workload", choices=sorted(WORK))
    ap.add_argument("--ndot", type=int, default=1)
    ap.add_argument("--reps", type=int, default=50)
    ap.add_argument("--shape", type=str, default=None)
    ap.add_argument("--save", type=str, default=None)
    ap.add_argument("--check", type=str, default=None)
    ap.add_argument("--variant", type=int, default=0)
    args = ap.parse_args()
    dim_flag, shape = WORK[args.workload]
    if args.shape:
        shape = tuple(int(v) for v in args.shape.split(","))
    ts = torch.cuda.Stream()
    torch.cuda.set_stream(ts)
    rt = S.Runtime(device=0, stream=ts.cuda_stream)
    st = S.benchmark_stencil(dim_flag, *shape)
    mat = S.CSRMatrix.stencil(rt, st, 1)
    n, nnz = mat.rows, mat.nnz
    e, c, rp = mat.device_fields()
    g = torch.Generator(device="cuda").manual_seed(7)
    x = torch.rand(n, dtype=torch.float64, device="cuda", generator=g) - 0.5
    y = torch.zeros(n, dtype=torch.float64, device="cuda")
    d = torch.zeros(2, dtype=torch.float64, device="cuda")
    L = _abi.lib()
    w = x.data_ptr() if args.ndot >= 1 else None
    o0 = d.data_ptr() if args.ndot >= 1 else None
    o1 = d.data_ptr() + 8 if args.ndot >= 2 else None

    def spmv():
        _abi.check(L.lsk_csr_spmv_f64(rt.ctx, ts.cuda_stream, n, nnz, e, c, rp, 0, x.data_ptr(), y.data_ptr(), w, o0, o1, args.variant),
                   "lsk_csr_spmv_f64")

    for _ in range(5):
        spmv()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(args.reps):
        spmv()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / args.reps
    nbytes = 16 * nnz + 32 * n
    out = {"workload": args.workload, "shape": shape, "ndot": args.ndot, "ms": round(ms, 5), "gbs": round(nbytes / ms / 1e6, 1),
           "frac_6535": round(nbytes / ms / 1e6 / 6535.7, 4), "dot": d.tolist(),
           "env": {k: v for k, v in os.environ.items() if k.startswith("LSK_")}}
    if args.save:
        torch.save({"y": y.cpu(), "d": d.cpu()}, args.save)
    if args.check:
        ref = torch.load(args.check)
        out["y_bit_identical"] = bool(torch.equal(ref["y"], y.cpu()))
        out["y_max_rel"] = float(((ref["y"] - y.cpu()).abs().max() / ref["y"].abs().max()).item())
        out["dot_rel"] = [float(abs(ref["d"][i] - d[i].item()) / max(abs(ref["d"][i]), 1e-300)) for i in range(args.ndot)]
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
