"""Developer probe: a handful of SpMV launches for ncu (matrix from the oracle generator)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from legionsolvers_b200 import kernels as K  # noqa: E402
from oracle import oracle as orc  # noqa: E402

dim_flag = int(sys.argv[1]) if len(sys.argv) > 1 else 3
nx = int(sys.argv[2]) if len(sys.argv) > 2 else 256
variant = int(sys.argv[3]) if len(sys.argv) > 3 else K.SPMV_STREAM
shape = (nx,) * (3 if dim_flag >= 3 else dim_flag)
ctx = K.Context()
off, val = orc.benchmark_stencil(dim_flag)
m = orc.stencil_csr(shape, off, val)
n = m.n_rows
entry, col = torch.from_numpy(m.entry).cuda(), torch.from_numpy(m.col).cuda()
rowptr = K.rect_tensor(m.rowptr)
x = torch.rand(n, dtype=torch.float64, device="cuda")
y = torch.zeros_like(x)
d = torch.zeros(1, dtype=torch.float64, device="cuda")
for _ in range(4):
    ctx.csr_spmv(n, m.nnz, entry, col, rowptr, 0, x, 0, y, dot_w=x, dot_out=d, variant=variant)
torch.cuda.synchronize()
print("done", d.item())
