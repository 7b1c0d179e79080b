import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from oracle import oracle as orc
from legionsolvers_b200 import solvers as S
sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "tests"))
from test_host_gpu import build_system
rt = S.Runtime(0)
off, val = orc.benchmark_stencil(2)
m = orc.stencil_csr((96, 96), off, val)
rng = np.random.default_rng(5)
rhs = [rng.uniform(0.5, 1.5, m.n_rows)]
for fused in (True, False):
    pl, opl, _, _ = build_system(rt, orc, m, 2, rhs=rhs)
    s, o = S.BiCGStabSolver(pl, fused=fused), orc.BiCGStabSolver(opl)
    for _ in range(25):
        s.step(); o.step()
    for name in ("rho", "alpha", "omega"):
        g, w = getattr(s, name), getattr(o, name)
        print(fused, name, np.array2string(np.abs(g - w) / np.abs(w), precision=1))
# GMRES per-column error
off, val = orc.benchmark_stencil(3)
m = orc.stencil_csr((14, 14, 14), off, val)
pl, opl, _, _ = build_system(rt, orc, m, 2)
s, o = S.GMRESSolver(pl, 30, fused=True), orc.GMRESSolver(opl, 30)
s.step(); o.step()
H, Ho = s.inner_products, o.inner_products
print("gmres col err", np.array2string(np.max(np.abs(H - Ho), axis=0), precision=1))
print("subdiag", np.array2string(np.array([Ho[j + 1, j] for j in range(30)]), precision=2))
