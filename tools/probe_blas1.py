"""Developer probe (not part of the product or the bench contract): BLAS-1 leaf kernels on 2^24 fp64 elements,
CUDA-event timing, GB/s on algorithmic bytes.  LSK_PROBE_N overrides the length."""
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from legionsolvers_b200 import kernels as K  # noqa: E402


def timeit(fn, iters=30, warm=5):
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
    n = int(os.environ.get("LSK_PROBE_N", 1 << 24))
    ctx = K.Context()
    x, y, w, z = (torch.rand(n, dtype=torch.float64, device="cuda") for _ in range(4))
    a = torch.tensor([0.5], dtype=torch.float64, device="cuda")
    d = torch.zeros(1, dtype=torch.float64, device="cuda")
    d2 = torch.zeros(1, dtype=torch.float64, device="cuda")
    one = torch.ones(1, dtype=torch.float64, device="cuda")
    for name, fn, nb in (
        ("scal", lambda: ctx.scal([one], y), 16), ("axpy", lambda: ctx.axpy([a], x, y), 24), ("xpay", lambda: ctx.xpay([a], x, y), 24),
        ("dot", lambda: ctx.dot(x, y, d), 16), ("dot2", lambda: ctx.dot2(x, y, d, d2), 16),
        ("axpy_dot", lambda: ctx.axpy_dot([a], x, y, w, d), 32), ("cg_update", lambda: ctx.cg_update(one, one, x, w, y, z, d), 48),
        ("cg_direction", lambda: ctx.cg_direction(one.clone(), one, x, y), 24), ("torch copy", lambda: y.copy_(x), 16),
    ):
        t = timeit(fn)
        print(f"{name:13s} {t * 1e6:8.1f} us  {nb * n / t / 1e9:8.1f} GB/s", flush=True)


if __name__ == "__main__":
    main()
