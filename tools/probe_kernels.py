"""Developer probe (not part of the product or the bench contract): times the leaf kernels on one
GPU with CUDA events.  Matrix comes from the oracle generator -- fine for a probe."""
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from legionsolvers_b200 import kernels as K  # noqa: E402
from oracle import oracle as orc  # noqa: E402


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e-3


def main():
    dim_flag = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    nx = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    shape = (nx,) * (3 if dim_flag >= 3 else dim_flag)
    ctx = K.Context()
    t0 = time.time()
    off, val = orc.benchmark_stencil(dim_flag)
    m = orc.stencil_csr(shape, off, val)
    print(f"generated {shape} nnz={m.nnz} in {time.time() - t0:.1f}s", flush=True)
    n = m.n_rows
    entry, col = torch.from_numpy(m.entry).cuda(), torch.from_numpy(m.col).cuda()
    rowptr = K.rect_tensor(m.rowptr)
    x = torch.rand(n, dtype=torch.float64, device="cuda")
    y = torch.zeros_like(x)
    d = torch.zeros(1, dtype=torch.float64, device="cuda")
    spmv_bytes = 16 * m.nnz + 32 * n
    for name, variant in (("stream", K.SPMV_STREAM), ("vector", K.SPMV_VECTOR), ("warp", K.SPMV_WARP)):
        t = timeit(lambda: ctx.csr_spmv(n, m.nnz, entry, col, rowptr, 0, x, 0, y, variant=variant))
        print(f"spmv {name:7s} {t * 1e3:8.3f} ms  {spmv_bytes / t / 1e9:8.1f} GB/s")
    t = timeit(lambda: ctx.csr_spmv(n, m.nnz, entry, col, rowptr, 0, x, 0, y, dot_w=x, dot_out=d, variant=K.SPMV_STREAM))
    print(f"spmv stream+dot {t * 1e3:8.3f} ms  {spmv_bytes / t / 1e9:8.1f} GB/s")
    a = torch.tensor([0.5], dtype=torch.float64, device="cuda")
    w, z = torch.rand_like(x), torch.rand_like(x)
    for name, fn, nb in (
        ("axpy", lambda: ctx.axpy([a], x, y), 24),
        ("xpay", lambda: ctx.xpay([a], x, y), 24),
        ("scal", lambda: ctx.scal([a], y), 16),
        ("dot", lambda: ctx.dot(x, y, d), 16),
        ("fill", lambda: ctx.fill(y, 0.0), 8),
        ("cg_update", lambda: ctx.cg_update(a, a, x, w, y, z, d), 48),
        ("torch copy", lambda: y.copy_(x), 16),
    ):
        t = timeit(fn)
        print(f"{name:10s} {t * 1e6:8.1f} us  {nb * n / t / 1e9:8.1f} GB/s")
    # tiny-launch latency
    s = torch.zeros(8, dtype=torch.float64, device="cuda")
    t = timeit(lambda: ctx.dot(s, s, d), iters=200)
    print(f"dot n=8 latency {t * 1e6:.2f} us (eager launch)")


if __name__ == "__main__":
    main()
