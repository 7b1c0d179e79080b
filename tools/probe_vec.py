"""Developer probe: the two vector kernels of the fused CG step timed alone at a given length (L2-resident or not)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from legionsolvers_b200 import kernels as K  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 21
ctx = K.Context()
p, q, x, r = (torch.rand(n, dtype=torch.float64, device="cuda") for _ in range(4))
sc = torch.tensor([1.0, 2.0, 0.5, 0.0], dtype=torch.float64, device="cuda")
hist = torch.zeros(1 << 16, dtype=torch.float64, device="cuda")
cnt = torch.zeros(1, dtype=torch.int64, device="cuda")


def timed(f, reps=200):
    for _ in range(10):
        f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        f()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3


t_upd = timed(lambda: ctx.cg_update(sc[0:1], sc[1:2], p, q, x, r, sc[3:4]))
t_dir = timed(lambda: ctx.cg_direction(sc[0:1].clone(), sc[2:3], r, p, hist, cnt))
t_dot = timed(lambda: ctx.dot(p, q, sc[3:4]))
t_axpy = timed(lambda: ctx.axpy([sc[2:3]], p, q))
print(f"n={n}: cg_update {t_upd:.2f} us ({48 * n / t_upd / 1e6:.0f} GB/s)  cg_direction {t_dir:.2f} us ({24 * n / t_dir / 1e6:.0f} GB/s)  "
      f"dot {t_dot:.2f} us  axpy {t_axpy:.2f} us")
