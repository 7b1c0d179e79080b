"""Developer probe (not part of the product or the bench contract): config C5 of SURVEY.md section 8d -- a synthetic
COO matrix with power-law row lengths (N = 2^22, row length min(1e5, floor(8 U^(-1/1.5))), ~1e8 non-zeros, entries
sorted by (row, col)) -- generated on the GPU with torch, then the COO segmented-reduction kernel timed with CUDA
events and checked against a CSR pass over the same matrix."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from legionsolvers_b200 import kernels as K  # noqa: E402


def main():
    n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 22)
    g = torch.Generator(device="cuda").manual_seed(12345)
    u = torch.rand(n, device="cuda", dtype=torch.float64, generator=g).clamp_min(1e-12)
    length = torch.clamp((8.0 * u.pow(-1.0 / 1.5)).floor().to(torch.int64), max=100_000)
    nnz = int(length.sum().item())
    row = torch.repeat_interleave(torch.arange(n, device="cuda", dtype=torch.int64), length)
    col = torch.randint(0, n, (nnz,), device="cuda", dtype=torch.int64, generator=g)
    key, _ = torch.sort(row * n + col)
    row, col = key // n, key % n
    del key
    entry = torch.rand(nnz, device="cuda", dtype=torch.float64, generator=g) * 2 - 1
    x = torch.rand(n, device="cuda", dtype=torch.float64, generator=g)
    y = torch.zeros(n, device="cuda", dtype=torch.float64)
    ctx = K.Context()
    print(f"N={n} nnz={nnz} mean row {nnz / n:.1f} max row {int(length.max())}", flush=True)

    def run():
        y.zero_()
        ctx.coo_spmv(nnz, entry, row, col, x, 0, y, 0, (0, n - 1), (0, n - 1))

    for _ in range(3):
        run()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    a.record()
    for _ in range(reps):
        run()
    b.record()
    torch.cuda.synchronize()
    t = a.elapsed_time(b) / reps * 1e-3
    bytes_ = 24 * nnz + 24 * n + 8 * n  # 24/nnz + x 8 + y RMW 16 per row (+ the zero fill timed with it)
    print(f"coo_spmv (+ zero fill) {t * 1e3:.3f} ms  {bytes_ / t / 1e9:.1f} GB/s  ({2 * nnz / t / 1e9:.1f} GFLOP/s)")
    # check against torch's own sparse product on the same data (fp64, different summation order)
    ref = torch.zeros_like(y)
    ref.index_add_(0, row, entry * x[col])
    err = float((y - ref).abs().max() / ref.abs().max())
    print(f"max rel difference vs index_add reference: {err:.2e}")
    assert err < 1e-12


if __name__ == "__main__":
    main()
