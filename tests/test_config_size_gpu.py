"""Parity AT THE BENCHMARK CONFIGURATIONS' SIZES (BASELINE.json configs 2-5), through the C ABI, against the oracle:

    C3  3-D 7-point 256^3   SpMV bit-exact, 20-iteration CG residual history <= 1e-10
    C2  2-D 5-point 8192^2  SpMV bit-exact
    C4  3-D 27-point 192^3  SpMV <= 1e-12 of |A||x| (4 lanes per row, tree order)
    C5  COO power-law       SpMV <= 1e-12 of |A||x| at N = 2^22 (~1e8 nnz); GMRES(30) Hessenberg at N = 2^18

The matrices are generated on the GPU by the product's generator (whose bit-exactness against the oracle's generator is
tested at small sizes in test_host_gpu.py and re-checked here on a window of rows) and downloaded for the oracle."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

REL = 1e-12


@pytest.fixture(scope="module")
def rt():
    from legionsolvers_b200.solvers import Runtime

    r = Runtime(device=0)
    yield r
    r.close()


def ramp(n):
    """x_i = ((i * 2654435761) mod 2^32) / 2^32 - 0.5  (SURVEY.md section 8d, config C2)."""
    i = np.arange(n, dtype=np.uint64)
    return ((i * np.uint64(2654435761)) % np.uint64(2 ** 32)).astype(np.float64) / 2.0 ** 32 - 0.5


def gpu_stencil(rt, oracle, dim_flag, shape):
    from legionsolvers_b200 import solvers as S

    nx, ny, nz = (list(shape) + [1, 1])[:3]
    gm = S.CSRMatrix.stencil(rt, S.benchmark_stencil(dim_flag, nx, ny, nz), 1)
    entry, col, rowptr = gm.slab_to_numpy()
    return gm, oracle.Matrix(gm.rows, gm.cols, entry, col, rowptr=rowptr)


def spmv_through_abi(rt, gm, x):
    from legionsolvers_b200 import _abi

    e, c, rp = gm.device_fields()
    xd = torch.from_numpy(x).cuda()
    y = torch.full((gm.rows,), 3.0, dtype=torch.float64, device="cuda")
    d = torch.zeros(1, dtype=torch.float64, device="cuda")
    _abi.check(_abi.lib().lsk_csr_spmv_f64(rt.ctx, rt.stream, gm.rows, gm.nnz, e, c, rp, 0, xd.data_ptr(), y.data_ptr(), xd.data_ptr(), d.data_ptr(),
                                           None, 0), "lsk_csr_spmv_f64")
    rt.fence()
    torch.cuda.synchronize()
    return y.cpu().numpy(), float(d.item())


@pytest.mark.parametrize("name,dim_flag,shape,exact", [("C3", 3, (256, 256, 256), True), ("C2", 2, (8192, 8192), True),
                                                       ("C4", 4, (192, 192, 192), False)])
def test_csr_spmv_at_config_size(rt, oracle, name, dim_flag, shape, exact):
    gm, m = gpu_stencil(rt, oracle, dim_flag, shape)
    off, val = oracle.benchmark_stencil(dim_flag)
    assert m.nnz == oracle.stencil_size(shape, off)
    # the generator, on a window of rows in the middle of the matrix, against the oracle's
    n = m.n_rows
    r_lo, r_hi = n // 2 - 1000, n // 2 + 1000
    k_lo, k_hi = int(m.rowptr["lo"][r_lo]), int(m.rowptr["hi"][r_hi])
    ref = oracle.stencil_csr(shape, off, val, k_range=(k_lo, k_hi), r_range=(r_lo, r_hi))
    np.testing.assert_array_equal(ref.col[k_lo:k_hi + 1], m.col[k_lo:k_hi + 1])
    np.testing.assert_array_equal(ref.entry[k_lo:k_hi + 1], m.entry[k_lo:k_hi + 1])
    np.testing.assert_array_equal(ref.rowptr[r_lo:r_hi + 1], m.rowptr[r_lo:r_hi + 1])
    del ref
    x = ramp(n)
    oracle.set_threads(16)
    want = np.zeros(n)
    oracle.csr_matvec(m, x, want)
    got, dot = spmv_through_abi(rt, gm, x)
    if exact:
        np.testing.assert_array_equal(got, want)  # same order, same rounding as the reference CPU body
    else:
        absax = np.zeros(n)
        oracle.csr_matvec(oracle.Matrix(n, n, np.abs(m.entry), m.col, rowptr=m.rowptr), np.abs(x), absax)
        assert np.all(np.abs(got - want) <= REL * np.maximum(absax, 1e-300))
    assert abs(dot - float(want @ x)) <= REL * float(np.abs(want) @ np.abs(x))
    gm.destroy()


def test_cg_history_256cubed(rt, oracle):
    """20 CG iterations on the headline system: GPU-generated matrix + fused kernels vs oracle-generated matrix + CPU bodies."""
    from legionsolvers_b200 import solvers as S

    shape, its = (256, 256, 256), 20
    n = 256 ** 3
    mat = S.CSRMatrix.stencil(rt, S.benchmark_stencil(3, *shape), 1)
    sol, rhs = S.PartitionedVector(rt, "sol", n, 1), S.PartitionedVector(rt, "rhs", n, 1)
    sol.zero_fill()
    rhs.constant_fill(1.0)
    pl = S.SquarePlanner(rt)
    pl.add_sol_vector(sol)
    pl.add_rhs_vector(rhs)
    pl.add_row_partitioned_matrix(mat, 0, 0)
    cg = S.CGSolver(pl)
    for _ in range(its):
        cg.step()
    got = cg.residual_norm_squared
    off, val = oracle.benchmark_stencil(3)
    m = oracle.stencil_csr(shape, off, val)
    oracle.set_threads(16)
    opl = oracle.Planner([n], [16])
    opl.fill(1, 1.0)
    opl.add_matrix(m)
    ocg = oracle.CGSolver(opl)
    for _ in range(its):
        ocg.step()
    want = ocg.residual_norm_squared
    assert got.size == want.size == its + 1
    assert float(np.max(np.abs(got - want) / np.abs(want))) <= 1e-10
    x, xo = pl.vector_to_numpy(0, 0, n), opl.vector(0)
    assert np.max(np.abs(x - xo)) <= 1e-10 * np.max(np.abs(xo))


def test_coo_spmv_power_law_1e8(rt, oracle):
    """Config C5's matrix at full size through lsk_coo_spmv_f64 (beta = 1 on a zero-filled y) vs the oracle's COO body."""
    from legionsolvers_b200 import _abi
    from legionsolvers_b200.workloads import power_law_coo

    n, entry, row, col = power_law_coo(22)
    assert entry.size > 95_000_000
    m = oracle.Matrix(n, n, entry, col, row=row)
    x = ramp(n)
    oracle.set_threads(16)
    want = np.zeros(n)
    oracle.coo_matvec(m, x, want)
    absax = np.zeros(n)
    oracle.coo_matvec(oracle.Matrix(n, n, np.abs(entry), col, row=row), np.abs(x), absax)
    e, r, c, xd = (torch.from_numpy(a).cuda() for a in (entry, row, col, x))
    y = torch.zeros(n, dtype=torch.float64, device="cuda")
    _abi.check(_abi.lib().lsk_coo_spmv_f64(rt.ctx, rt.stream, entry.size, e.data_ptr(), r.data_ptr(), c.data_ptr(), xd.data_ptr(), y.data_ptr(),
                                           0, n - 1, 0, n - 1), "lsk_coo_spmv_f64")
    rt.fence()
    torch.cuda.synchronize()
    got = y.cpu().numpy()
    assert np.all(np.abs(got - want) <= REL * np.maximum(absax, 1e-300))


def test_gmres30_hessenberg_on_power_law_coo(rt, oracle):
    """GMRES(30) on the C5 matrix at N = 2^18 (6 M nnz): Hessenberg entries of the first cycle against the oracle,
    per column window (Arnoldi amplifies the dot-order difference column by column), and the Arnoldi relation itself."""
    from legionsolvers_b200 import solvers as S
    from legionsolvers_b200.workloads import power_law_coo

    n, entry, row, col = power_law_coo(18)
    m = oracle.Matrix(n, n, entry, col, row=row)
    gm = S.COOMatrix.from_host(rt, n, n, entry, row, col)
    sol, rhs = S.PartitionedVector(rt, "sol", n, 1), S.PartitionedVector(rt, "rhs", n, 1)
    sol.zero_fill()
    rhs.constant_fill(1.0)
    pl = S.SquarePlanner(rt)
    pl.add_sol_vector(sol)
    pl.add_rhs_vector(rhs)
    pl.add_row_partitioned_matrix(gm, 0, 0)
    restart = 30
    s = S.GMRESSolver(pl, restart)
    s.step()
    H = s.inner_products
    oracle.set_threads(8)
    opl = oracle.Planner([n], [8])
    opl.fill(1, 1.0)
    opl.add_matrix(m)
    os_ = oracle.GMRESSolver(opl, restart)
    os_.step()
    Ho = os_.inner_products
    scale = np.max(np.abs(Ho))
    # a strictly diagonally dominant, non-symmetric matrix: Arnoldi is well conditioned here, the drift stays small
    assert np.max(np.abs(H[:, :10] - Ho[:, :10])) <= 1e-11 * scale
    assert np.max(np.abs(H - Ho)) <= 1e-8 * scale
    A = m.to_scipy()
    V = np.stack([pl.vector_to_numpy(2 + j, 0, n) for j in range(restart + 1)], axis=1)
    V[:, restart] /= H[restart, restart - 1]  # the reference leaves the last vector un-normalised
    np.testing.assert_allclose(A @ V[:, :restart], V @ H, rtol=0, atol=1e-10 * scale)
