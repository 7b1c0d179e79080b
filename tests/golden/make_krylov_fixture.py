"""Generates tests/golden/krylov_cg.npz by IMPORTING the reference's own numpy CG
(/root/reference/scripts/krylov.py, function cg) in the authoring container.  The reference tree
does not travel to the GPU box, so the vectors are committed; rerun this script to regenerate.

    python tests/golden/make_krylov_fixture.py
"""
import importlib.util
import sys
from pathlib import Path

import numpy as np

REF = Path("/root/reference/scripts/krylov.py")
spec = importlib.util.spec_from_file_location("ref_krylov", REF)
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

out = {}
for name, n in (("lap1d_n24", 24), ("lap1d_n100", 100)):
    A = 2.0 * np.eye(n) - np.eye(n, k=1) - np.eye(n, k=-1)  # the Test05/06 matrix, dense
    b = np.ones(n)
    x0 = np.zeros(n)
    iterates = ref.cg(A, b, x0.copy())
    out[f"{name}_iterates"] = np.array(iterates[:12])
# a non-trivial SPD system with a deterministic pseudo-random right-hand side
n = 36
g = np.arange(n * n, dtype=np.float64).reshape(n, n)
M = np.cos(0.37 * g) / n
A = M @ M.T + np.eye(n)
b = np.sin(0.11 * np.arange(n) + 0.3)
out["spd36_A"] = A
out["spd36_b"] = b
out["spd36_iterates"] = np.array(ref.cg(A, b, np.zeros(n))[:12])
np.savez_compressed(Path(__file__).with_name("krylov_cg.npz"), **out)
print({k: v.shape for k, v in out.items()})
