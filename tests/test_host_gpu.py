"""GPU parity of the host layer (PartitionedVector / CSRMatrix / COOMatrix / SquarePlanner / solvers,
driven through include/lsk_solvers.h) against the CPU oracle and the reference's golden vectors.

The tests mirror the reference's own programs: Test02VectorOperations, Test03/04 partitioning,
Test05/06 CG, BenchmarkStencil (CG / BiCGStab / GMRES on the stencil matrices, 1 or 2 spaces)."""
import json
from pathlib import Path

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLD = json.loads((Path(__file__).parent / "golden" / "reference_goldens.json").read_text())


@pytest.fixture(scope="module")
def rt():
    from legionsolvers_b200.solvers import Runtime

    r = Runtime(device=0)
    yield r
    r.close()


def build_system(rt, oracle, m, pieces, spaces=1, rhs=None):
    """sol = 0, rhs = 1 (or given), the matrix on every diagonal block -- on the GPU and in the oracle."""
    from legionsolvers_b200 import solvers as S

    n = m.n_rows
    gm = (S.CSRMatrix.from_host(rt, n, m.n_cols, m.entry, m.col, m.rowptr) if m.is_csr
          else S.COOMatrix.from_host(rt, n, m.n_cols, m.entry, m.row, m.col))
    pl = S.SquarePlanner(rt)
    opl = oracle.Planner([n] * spaces, [pieces] * spaces)
    vecs = []
    for s in range(spaces):
        sol = S.PartitionedVector(rt, f"sol{s}", n, pieces)
        sol.zero_fill()
        pl.add_sol_vector(sol)
        vecs.append(sol)
    for s in range(spaces):
        b = S.PartitionedVector(rt, f"rhs{s}", n, pieces)
        if rhs is None:
            b.constant_fill(1.0)
            opl.fill(1, 1.0)
        else:
            b.from_numpy(rhs[s])
            opl.vector(1, s)[:] = rhs[s]
        pl.add_rhs_vector(b)
        vecs.append(b)
    for s in range(spaces):
        pl.add_row_partitioned_matrix(gm, s, s)
        opl.add_matrix(m, s, s)
    return pl, opl, gm, vecs


_trace_ids = iter(range(1000, 1_000_000))


def new_trace_id():
    """A trace id names ONE recorded launch sequence (as in Legion): never reuse it for another solver."""
    return next(_trace_ids)


def rel(got, want):
    got, want = np.asarray(got), np.asarray(want)
    return float(np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-300)))


def history_err(got, want, floor=1e-12):
    """Relative error of a residual^2 history over its meaningful window: entries are compared relative
    to max(want_i, floor * want_0).  Once |r|^2 has dropped 12 orders of magnitude below |b|^2 the
    recurrence residual is rounding noise of the r -= alpha q updates in BOTH implementations, and a
    relative comparison of that noise says nothing; there the error is measured against the floor."""
    got, want = np.asarray(got), np.asarray(want)
    return float(np.max(np.abs(got - want) / np.maximum(np.abs(want), floor * abs(want[0]))))


# ---- StencilGenerator on the GPU: bit-exact -----------------------------------------------------------
@pytest.mark.parametrize("dim_flag,shape", [(1, (101,)), (2, (19, 23)), (2, (256, 256)), (3, (9, 10, 11)),
                                            (3, (32, 32, 32)), (4, (7, 8, 9)), (4, (24, 24, 24))])
@pytest.mark.parametrize("pieces", [1, 4])
def test_stencil_generator_bit_exact(rt, oracle, dim_flag, shape, pieces):
    from legionsolvers_b200 import solvers as S

    off, val = oracle.benchmark_stencil(dim_flag)
    want = oracle.stencil_csr(shape, off, val)
    nx, ny, nz = (list(shape) + [1, 1])[:3]
    st = S.benchmark_stencil(dim_flag, nx, ny, nz)
    assert S.stencil_size(st) == want.nnz
    gm = S.CSRMatrix.stencil(rt, st, pieces)
    assert (gm.rows, gm.cols, gm.nnz) == (want.n_rows, want.n_cols, want.nnz)
    assert (gm.slab_r_lo, gm.slab_r_hi, gm.slab_k_lo, gm.slab_k_hi) == (0, want.n_rows - 1, 0, want.nnz - 1)
    entry, col, rowptr = gm.slab_to_numpy()
    np.testing.assert_array_equal(col, want.col)
    np.testing.assert_array_equal(entry, want.entry)
    np.testing.assert_array_equal(rowptr, want.rowptr)
    gm.destroy()


def test_stencil_generator_custom_and_column_major(rt, oracle):
    from legionsolvers_b200 import solvers as S

    shape = (6, 5, 7)
    off = np.array([(0, 0, 0), (2, -1, 0), (-1, 0, 3), (0, 1, -2), (1, 1, 1), (0, 0, 0)], dtype=np.int64)
    val = np.array([3.0, -1.5, 0.25, 7.0, -2.0, 1.0])  # duplicate offset: ties broken by entry
    for order in (0, 1):
        want = oracle.stencil_csr(shape, off, val, order=order)
        gm = S.CSRMatrix.stencil(rt, S.make_stencil(shape, off, val, order), 3)
        entry, col, rowptr = gm.slab_to_numpy()
        np.testing.assert_array_equal(col, want.col)
        np.testing.assert_array_equal(entry, want.entry)
        np.testing.assert_array_equal(rowptr, want.rowptr)
        gm.destroy()


# ---- Test03 / Test04: partitions ----------------------------------------------------------------------------
@pytest.mark.parametrize("fmt", ["csr", "coo"])
def test_partition_goldens(rt, oracle, fmt):
    g = GOLD["partition_n20_p4"]
    m = oracle.laplacian_1d_csr(g["n"]) if fmt == "csr" else oracle.laplacian_1d_coo(g["n"])
    pl, opl, gm, _ = build_system(rt, oracle, m, g["pieces"])
    for c in range(g["pieces"]):
        assert list(pl.range_bounds(0, c)) == g["range_partition"][c]
        assert list(pl.kernel_bounds(0, c)) == g["matrix_partition"][c]
        assert list(pl.ghost_bounds(0, c)) == g["domain_partition"][c]


@pytest.mark.parametrize("dim_flag,shape,pieces", [(2, (20, 24), 4), (3, (12, 10, 8), 8), (4, (9, 9, 9), 3), (3, (8, 8, 8), 5)])
def test_partition_index_sets_bit_exact(rt, oracle, dim_flag, shape, pieces):
    """Bounding intervals through the planner AND the exact index sets through the *_flags kernels
    equal the oracle's image_range / image / preimage / preimage_range, piece by piece."""
    import ctypes as C

    from legionsolvers_b200 import _abi
    from legionsolvers_b200 import kernels as K

    off, val = oracle.benchmark_stencil(dim_flag)
    m = oracle.stencil_csr(shape, off, val)
    coo = m.to_coo()
    pl, opl, gm, _ = build_system(rt, oracle, m, pieces)
    L, ctx, st = _abi.lib(), rt.ctx, rt.stream
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()  # noqa: E731
    rowptr_d, col_d, row_d = K.rect_tensor(m.rowptr), dev(m.col), dev(coo.row)
    lo, hi = oracle.equal_partition(m.n_rows, pieces)
    for c in range(pieces):
        assert pl.range_bounds(0, c) == (int(lo[c]), int(hi[c])) == opl.piece_bounds(0, c)
        assert pl.kernel_bounds(0, c) == opl.kernel_bounds(0, c)
        assert pl.ghost_bounds(0, c) == opl.ghost_bounds(0, c)
        r_lo, r_hi = int(lo[c]), int(hi[c])
        kflags = torch.zeros(m.nnz, dtype=torch.uint8, device="cuda")
        _abi.check(L.lsk_image_range_flags(ctx, st, r_hi - r_lo + 1, rowptr_d[r_lo:].data_ptr(), 0, m.nnz, kflags.data_ptr()), "image_range")
        want_k = oracle.image_range(m, r_lo, r_hi)
        np.testing.assert_array_equal(kflags.cpu().numpy(), want_k)
        dflags = torch.zeros(m.n_cols, dtype=torch.uint8, device="cuda")
        _abi.check(L.lsk_image_flags(ctx, st, m.nnz, col_d.data_ptr(), kflags.data_ptr(), 0, m.n_cols, dflags.data_ptr()), "image")
        np.testing.assert_array_equal(dflags.cpu().numpy(), oracle.image(m.col, want_k, m.n_cols))
        pflags = torch.zeros(m.nnz, dtype=torch.uint8, device="cuda")
        _abi.check(L.lsk_preimage_flags(ctx, st, m.nnz, row_d.data_ptr(), r_lo, r_hi, pflags.data_ptr()), "preimage")
        np.testing.assert_array_equal(pflags.cpu().numpy(), oracle.preimage(coo.row, r_lo, r_hi))
        np.testing.assert_array_equal(pflags.cpu().numpy(), want_k)  # COO and CSR kernel pieces coincide
        rflags = torch.zeros(m.n_rows, dtype=torch.uint8, device="cuda")
        _abi.check(L.lsk_preimage_range_flags(ctx, st, m.n_rows, rowptr_d.data_ptr(), 0, m.nnz, kflags.data_ptr(), rflags.data_ptr()), "preimage_range")
        np.testing.assert_array_equal(rflags.cpu().numpy(), oracle.preimage_range(m, want_k))
        # kernel partition from a DOMAIN partition: preimage of col (src/CSRMatrix.cpp:68-86)
        _abi.check(L.lsk_preimage_flags(ctx, st, m.nnz, col_d.data_ptr(), r_lo, r_hi, pflags.data_ptr()), "preimage col")
        np.testing.assert_array_equal(pflags.cpu().numpy(), oracle.preimage(m.col, r_lo, r_hi))
    lo2 = np.zeros(pieces, dtype=np.int64); hi2 = np.zeros(pieces, dtype=np.int64)
    L.lsk_equal_partition(m.n_rows, pieces, lo2.ctypes.data_as(C.c_void_p), hi2.ctypes.data_as(C.c_void_p))
    np.testing.assert_array_equal(lo2, lo); np.testing.assert_array_equal(hi2, hi)
    for p in range(pieces):
        assert L.lsk_shard(p, pieces, 2) == oracle.shard(p, pieces, 2)


# ---- Test02: vector operations ---------------------------------------------------------------------------------
def test_vector_operations_chain(rt):
    from legionsolvers_b200.solvers import PartitionedVector

    g = GOLD["blas1_chain"]
    u = PartitionedVector(rt, "u", g["elements"], g["pieces"])
    v = PartitionedVector(rt, "v", g["elements"], g["pieces"])
    w = PartitionedVector(rt, "w", g["elements"], g["pieces"])
    u.constant_fill(g["u"]); v.constant_fill(g["v"]); w.assign(u)
    w.axpy(1.0, v); v.xpay(-1.0, u); u.axpy(-0.5, v); u.axpy(-0.5, w)
    assert u.dot(u) == g["expected_dot"]
    assert v.dot(v) == pytest.approx(100 * (1.5 - 2.7) ** 2, rel=1e-14)
    u.scal(3.0)
    assert np.all(u.to_numpy() == 0.0)


# ---- Test05 / Test06: CG golden history ---------------------------------------------------------------------------
@pytest.mark.parametrize("fmt", ["csr", "coo"])
@pytest.mark.parametrize("fused", [True, False])
@pytest.mark.parametrize("traced", [False, True])
def test_cg_golden_history(rt, oracle, fmt, fused, traced):
    from legionsolvers_b200.solvers import CGSolver

    g = GOLD["cg_residual_norm_squared_sorted"]
    m = oracle.laplacian_1d_csr(g["n"]) if fmt == "csr" else oracle.laplacian_1d_coo(g["n"])
    pl, _, _, _ = build_system(rt, oracle, m, g["pieces"])
    cg = CGSolver(pl, fused=fused)
    tid = new_trace_id()
    for i in range(g["iterations"]):
        if traced:
            rt.begin_trace(tid)
        cg.step()
        if traced:
            rt.end_trace(tid)
    rr = cg.residual_norm_squared
    assert sorted(rr) == [float(v) for v in g["values"]]
    assert list(rr) == [100.0, 4900.0, 4704.0, 4512.0, 4324.0, 4140.0, 3960.0, 3784.0, 3612.0, 3444.0, 3280.0]


# ---- planner mat-vec: bit-exact --------------------------------------------------------------------------------------
@pytest.mark.parametrize("dim_flag,shape,pieces", [(2, (64, 64), 4), (3, (24, 24, 24), 8), (4, (16, 16, 16), 4), (3, (20, 20, 20), 1)])
def test_planner_matvec_bit_exact(rt, oracle, dim_flag, shape, pieces):
    """Bit-exact where the planner's mat-vec resolves to the thread-per-row kernel (<= 12 non-zeros per row: the 5- and
    7-point stencils); within 1e-12 of |A||x| for the 27-point stencil (2-8 lanes per row, tree order)."""
    off, val = oracle.benchmark_stencil(dim_flag)
    m = oracle.stencil_csr(shape, off, val)
    rng = np.random.default_rng(11)
    x = rng.standard_normal(m.n_rows)
    pl, opl, _, _ = build_system(rt, oracle, m, pieces, rhs=[x])
    pl.allocate_workspace(2); opl.allocate_workspace(2)
    pl.matvec(2, 1); opl.matvec(2, 1)
    y = opl.vector(2)
    exact = m.nnz <= 12 * m.n_rows
    absax = np.zeros(m.n_rows)
    oracle.csr_matvec(oracle.Matrix(m.n_rows, m.n_cols, np.abs(m.entry), m.col, rowptr=m.rowptr), np.abs(x), absax)

    def check(got):
        if exact:
            np.testing.assert_array_equal(got, y)
        else:
            assert np.all(np.abs(got - y) <= 1e-12 * np.maximum(absax, 1e-300))

    check(pl.vector_to_numpy(2, 0, m.n_rows))
    yw, yy = pl.matvec_dot(3, 1, 1, want_yy=True)
    check(pl.vector_to_numpy(3, 0, m.n_rows))
    assert abs(yw - float(y @ x)) <= 1e-12 * float(np.abs(y) @ np.abs(x))
    assert abs(yy - float(y @ y)) <= 1e-12 * float(y @ y)
    # COO block through the same planner: accumulate semantics on a zero-filled destination
    plc, oplc, _, _ = build_system(rt, oracle, m.to_coo(), pieces, rhs=[x])
    plc.allocate_workspace(1)
    plc.matvec(2, 1)
    got = plc.vector_to_numpy(2, 0, m.n_rows)
    np.testing.assert_allclose(got, y, rtol=0, atol=1e-12 * np.max(np.abs(y)))


# ---- solver histories vs the oracle ------------------------------------------------------------------------------------
CG_CASES = [
    ("C1: 2-D 5-pt 256x256, 4 pieces", 2, (256, 256), 4, 1, 80),
    ("C1: 2-D 5-pt 256x256, 1 piece", 2, (256, 256), 1, 1, 80),
    ("3-D 7-pt 40^3, 8 pieces", 3, (40, 40, 40), 8, 1, 60),
    ("3-D 27-pt 24^3, 4 pieces", 4, (24, 24, 24), 4, 1, 40),
    ("BenchmarkStencil: 2 spaces, 3-D 7-pt 24^3", 3, (24, 24, 24), 4, 2, 40),
]


@pytest.mark.parametrize("name,dim_flag,shape,pieces,spaces,its", CG_CASES, ids=[c[0] for c in CG_CASES])
@pytest.mark.parametrize("fused", [True, False])
def test_cg_history_vs_oracle(rt, oracle, name, dim_flag, shape, pieces, spaces, its, fused):
    """CG residual histories within 1e-10 relative (north_star tolerance), solution within 1e-10."""
    from legionsolvers_b200.solvers import CGSolver

    off, val = oracle.benchmark_stencil(dim_flag)
    m = oracle.stencil_csr(shape, off, val)
    pl, opl, _, _ = build_system(rt, oracle, m, pieces, spaces=spaces)
    cg, ocg = CGSolver(pl, fused=bool(fused)), oracle.CGSolver(opl)
    tid = new_trace_id()
    for i in range(its):
        rt.begin_trace(tid)
        cg.step()
        rt.end_trace(tid)
        ocg.step()
    got, want = cg.residual_norm_squared, ocg.residual_norm_squared
    assert got.size == want.size == its + 1
    assert history_err(got, want) <= 1e-10
    for s in range(spaces):
        x, xo = pl.vector_to_numpy(0, s, m.n_rows), opl.vector(0, s)
        assert np.max(np.abs(x - xo)) <= 1e-10 * np.max(np.abs(xo))


def test_cg_converges_in_same_iteration_count(rt, oracle):
    """Iterations to reach |r|^2 <= 1e-16 |b|^2 identical +-1 between GPU and oracle."""
    from legionsolvers_b200.solvers import CGSolver

    off, val = oracle.benchmark_stencil(3)
    m = oracle.stencil_csr((20, 20, 20), off, val)
    pl, opl, _, _ = build_system(rt, oracle, m, 4)
    cg, ocg = CGSolver(pl), oracle.CGSolver(opl)
    for _ in range(120):
        cg.step(); ocg.step()
    got, want = cg.residual_norm_squared, ocg.residual_norm_squared
    tol = 1e-16 * want[0]
    it_gpu, it_cpu = int(np.argmax(got <= tol)), int(np.argmax(want <= tol))
    assert it_cpu > 0 and abs(it_gpu - it_cpu) <= 1


@pytest.mark.parametrize("dim_flag,shape,pieces,spaces", [(4, (20, 20, 20), 4, 1), (2, (96, 96), 2, 1), (3, (16, 16, 16), 4, 2)])
@pytest.mark.parametrize("fused", [True, False])
def test_bicgstab_history_vs_oracle(rt, oracle, dim_flag, shape, pieces, spaces, fused):
    from legionsolvers_b200.solvers import BiCGStabSolver

    off, val = oracle.benchmark_stencil(dim_flag)
    m = oracle.stencil_csr(shape, off, val)
    rng = np.random.default_rng(5)
    rhs = [rng.uniform(0.5, 1.5, m.n_rows) for _ in range(spaces)]
    pl, opl, _, _ = build_system(rt, oracle, m, pieces, spaces=spaces, rhs=rhs)
    s, os_ = BiCGStabSolver(pl, fused=fused), oracle.BiCGStabSolver(opl)
    its = 25
    tid = new_trace_id()
    for _ in range(its):
        rt.begin_trace(tid)
        s.step()
        rt.end_trace(tid)
        os_.step()
    # BiCGStab is a Lanczos-type recurrence: two valid floating-point evaluations that differ only in the ORDER of the
    # dot-product sums (sequential in the reference CPU task, tree on the GPU) drift apart geometrically, exactly as
    # the reference's own cuBLAS variant would.  The window is not tuned: the ORACLE ITSELF is run a second time with
    # another valid dot order (pairwise tree), and the GPU may deviate from the reference order by no more than a
    # constant times what that re-ordering alone does to the same recurrence (plus the 1e-12 of a single dot).
    oracle.set_dot_order(1)
    try:
        _, opl2, _, _ = build_system(rt, oracle, m, pieces, spaces=spaces, rhs=rhs)
        os2 = oracle.BiCGStabSolver(opl2)
        for _ in range(its):
            os2.step()
        alt = {name: getattr(os2, name) for name in ("rho", "alpha", "omega")}
    finally:
        oracle.set_dot_order(0)
    for name in ("rho", "alpha", "omega"):
        got, want = getattr(s, name), getattr(os_, name)
        assert got.size == want.size == its + 1
        assert got[0] == want[0]
        err_gpu = np.abs(got[1:] - want[1:]) / np.abs(want[1:])
        err_alt = np.abs(alt[name][1:] - want[1:]) / np.abs(want[1:])
        bound = 50.0 * np.maximum.accumulate(err_alt) + 1e-12
        assert np.all(err_gpu <= bound), (name, err_gpu / bound)
        assert np.max(err_gpu[:8]) <= 1e-11, name  # the first steps, before any amplification
    for sp in range(spaces):
        x, xo, xa = pl.vector_to_numpy(0, sp, m.n_rows), opl.vector(0, sp), opl2.vector(0, sp)
        assert np.max(np.abs(x - xo)) <= 50.0 * np.max(np.abs(xa - xo)) + 1e-10 * np.max(np.abs(xo))
    # and the GPU solution really solves the system: true residual has dropped
    A = m.to_scipy()
    for sp in range(spaces):
        x = pl.vector_to_numpy(0, sp, m.n_rows)
        xo = opl.vector(0, sp)
        res, res_o = np.linalg.norm(rhs[sp] - A @ x), np.linalg.norm(rhs[sp] - A @ xo)
        assert res <= 1.001 * res_o + 1e-12 * np.linalg.norm(rhs[sp]), (res, res_o)


@pytest.mark.parametrize("dim_flag,shape,pieces,restart", [(2, (48, 48), 4, 10), (3, (14, 14, 14), 2, 30), (4, (10, 10, 10), 1, 6)])
@pytest.mark.parametrize("fused", [True, False])
def test_gmres_hessenberg_vs_oracle(rt, oracle, dim_flag, shape, pieces, restart, fused):
    """Two restart cycles; parity on the Hessenberg ("inner_products") entries and on the
    reference's placeholder solution update."""
    from legionsolvers_b200.solvers import GMRESSolver

    off, val = oracle.benchmark_stencil(dim_flag)
    m = oracle.stencil_csr(shape, off, val)
    pl, opl, _, _ = build_system(rt, oracle, m, pieces)
    s, os_ = GMRESSolver(pl, restart, fused=fused), oracle.GMRESSolver(opl, restart)
    A = m.to_scipy()
    Ho_alt = []
    oracle.set_dot_order(1)  # the same two cycles on the oracle with another valid dot order: the sensitivity reference
    try:
        _, opl2, _, _ = build_system(rt, oracle, m, pieces)
        os2 = oracle.GMRESSolver(opl2, restart)
        for _ in range(2):
            os2.step()
            Ho_alt.append(os2.inner_products.copy())
        xo_alt = opl2.vector(0).copy()
    finally:
        oracle.set_dot_order(0)
    for cycle in range(2):
        s.step(); os_.step()
        H, Ho = s.inner_products, os_.inner_products
        scale = np.max(np.abs(Ho))
        # Arnoldi on a symmetric matrix is Lanczos: rounding differences between two summation orders grow geometrically
        # with the column index.  As for BiCGStab the window is DERIVED: the oracle re-run with a pairwise dot order
        # (Ho_alt, below) shows what a change of order alone does, column by column; the GPU stays within a constant of it.
        err_gpu = np.max(np.abs(H - Ho), axis=0) / scale
        err_alt = np.max(np.abs(Ho_alt[cycle] - Ho), axis=0) / scale
        bound = 50.0 * np.maximum.accumulate(err_alt) + 1e-12
        assert np.all(err_gpu <= bound), (cycle, err_gpu / bound)
        if cycle == 0:
            assert np.max(err_gpu[:min(10, restart)]) <= 1e-11
        # ... and the GPU result must satisfy the Arnoldi relation A V_m = V_{m+1} H on its own
        V = np.stack([pl.vector_to_numpy(2 + j, 0, m.n_rows) for j in range(restart + 1)], axis=1)
        V[:, restart] /= H[restart, restart - 1]  # the reference leaves the last vector un-normalised
        np.testing.assert_allclose(A @ V[:, :restart], V @ H, rtol=0, atol=1e-10 * scale)
        x, xo = pl.vector_to_numpy(0, 0, m.n_rows), opl.vector(0)
        if cycle == 1:  # the placeholder update inherits the drift: bounded by the re-ordered oracle's own deviation
            assert np.max(np.abs(x - xo)) <= 50.0 * np.max(np.abs(xo_alt - xo)) + 1e-10 * np.max(np.abs(xo))


def test_matrix_market_to_gpu_solve(rt, oracle, tmp_path):
    """A matrix written to / read from a Matrix Market file (lsk_mm_*), uploaded as CSR and as COO: the mat-vec equals the one
    of the matrix it came from bit for bit (CSR, thread-per-row) / to 1e-12 (COO)."""
    import scipy.io
    import scipy.sparse as sp

    from legionsolvers_b200 import solvers as S

    off, val = oracle.benchmark_stencil(2)
    m = oracle.stencil_csr((40, 36), off, val)
    n = m.n_rows
    rp = np.ascontiguousarray(m.rowptr).view(np.int64).reshape(-1, 2)
    rows_of = np.repeat(np.arange(n), rp[:, 1] - rp[:, 0] + 1)
    path = tmp_path / "laplace2d.mtx"
    S.write_matrix_market(path, n, n, m.entry, rows_of, m.col)
    a = sp.csr_matrix(scipy.io.mmread(str(path)))        # an independent reader agrees on the file's content
    x = np.random.default_rng(3).standard_normal(n)
    for cls, exact in ((S.CSRMatrix, True), (S.COOMatrix, False)):
        gm = cls.from_matrix_market(rt, path)
        pl = S.SquarePlanner(rt)
        sol, rhs = S.PartitionedVector(rt, "sol", n, 4), S.PartitionedVector(rt, "rhs", n, 4)
        sol.zero_fill()
        rhs.from_numpy(x)
        pl.add_sol_vector(sol)
        pl.add_rhs_vector(rhs)
        pl.add_row_partitioned_matrix(gm, 0, 0)
        pl.allocate_workspace(1)
        pl.matvec(2, 1)
        y = pl.vector_to_numpy(2, 0, n)
        want = np.zeros(n)
        oracle.csr_matvec(m, x, want)
        if exact:
            np.testing.assert_array_equal(y, want)
        else:
            assert np.max(np.abs(y - want)) <= 1e-12 * np.max(np.abs(want))
        assert np.max(np.abs(y - a @ x)) <= 1e-12 * np.max(np.abs(want))


def test_kernel_launch_accounting(rt, oracle):
    from legionsolvers_b200.solvers import CGSolver

    off, val = oracle.benchmark_stencil(3)
    m = oracle.stencil_csr((16, 16, 16), off, val)
    pl, _, _, _ = build_system(rt, oracle, m, 1)
    # the fused step on one piece: spmv+dot, cg_update, cg_direction (xpay + history append)
    cg1 = CGSolver(pl, fused=True)
    rt.fence()
    before = rt.kernel_launches
    tid = new_trace_id()
    for _ in range(5):
        rt.begin_trace(tid)
        cg1.step()
        rt.end_trace(tid)
    rt.fence()
    # (a piece this small takes the one-launch tail: spmv+dot, then cg_update + cg_direction as lsk_cg_tail_f64)
    assert rt.kernel_launches - before == 5 * 2
    # ... and on 4 local pieces
    pl4, _, _, _ = build_system(rt, oracle, m, 4)
    cg4 = CGSolver(pl4, fused=True)
    rt.fence()
    before = rt.kernel_launches
    tid = new_trace_id()
    for _ in range(5):
        rt.begin_trace(tid)
        cg4.step()
        rt.end_trace(tid)
    rt.fence()
    # per step: 4 pieces x (spmv+dot, cg_update, xpay) + 2 colour-order folds of 4 partials (copy + 3 adds) + append
    assert rt.kernel_launches - before == 5 * (4 * 3 + 2 * 4 + 1)


# ---- transposed mat-vecs and the finished GMRES update ----------------------------------------
@pytest.mark.parametrize("fmt", ["csr", "coo"])
@pytest.mark.parametrize("pieces", [1, 4])
def test_rmatvec_vs_oracle_and_scipy(rt, oracle, fmt, pieces):
    """dst = A^T src through the planner (CSRRmatvecTask / COORmatvecTask, reserved but `assert(false)` in the reference):
    against the oracle's definition and, independently, scipy's transpose.  Non-symmetric matrix, random values."""
    rng = np.random.default_rng(3)
    n = 5000
    lens = rng.integers(1, 12, n)
    lens[[7, 3000]] = [400, 1500]
    starts = np.concatenate([[0], np.cumsum(lens)[:-1]])
    col = np.concatenate([np.sort(rng.choice(n, size=l, replace=False)) for l in lens]).astype(np.int64)
    entry = rng.standard_normal(col.size)
    rowptr = np.empty(n, dtype=oracle.RECT_DTYPE)
    rowptr["lo"], rowptr["hi"] = starts, starts + lens - 1
    m = oracle.Matrix(n, n, entry, col, rowptr=rowptr)
    if fmt == "coo":
        m = m.to_coo()
    x = rng.standard_normal(n)
    pl, _, _, _ = build_system(rt, oracle, m, pieces, rhs=[x])
    pl.allocate_workspace(1)
    pl.rmatvec(2, 1)
    got = pl.vector_to_numpy(2, 0, n)
    want = np.zeros(n)
    oracle.rmatvec(m, x, want)
    At = m.to_scipy().T.tocsr()
    scale = np.abs(At) @ np.abs(x)
    assert np.all(np.abs(got - want) <= 1e-12 * np.maximum(scale, 1e-300))
    assert np.all(np.abs(got - At @ x) <= 1e-12 * np.maximum(scale, 1e-300))
    # adjoint identity through both products of the planner: <A^T x, z> = <x, A z>
    z = rng.standard_normal(n)
    pl.vector_from_numpy(0, 0, z)
    pl.matvec(2, 0)
    Az = pl.vector_to_numpy(2, 0, n)
    assert abs(float(got @ z) - float(x @ Az)) <= 1e-11 * float(np.abs(got) @ np.abs(z))


@pytest.mark.parametrize("case,restart", [("7pt", 10), ("7pt", 30), ("coo", 30)])
@pytest.mark.parametrize("traced", [False, True])
def test_gmres_real_update_reduces_the_true_residual(rt, oracle, case, restart, traced):
    """GMRES with the FINISHED update (Givens least squares + SOL += V y): the true residual || b - A x || equals the
    least-squares minimum the kernel reports, decreases from cycle to cycle, and the update is what numpy's lstsq gives.
    The default (placeholder) mode is covered by test_gmres_hessenberg_vs_oracle."""
    from legionsolvers_b200.solvers import GMRESSolver
    from legionsolvers_b200.workloads import power_law_coo

    if case == "7pt":
        off, val = oracle.benchmark_stencil(3)
        m = oracle.stencil_csr((18, 18, 18), off, val)
    else:
        n_, e_, r_, c_ = power_law_coo(13)
        m = oracle.Matrix(n_, n_, e_, c_, row=r_)
    n = m.n_rows
    rng = np.random.default_rng(9)
    b = rng.uniform(0.5, 1.5, n)
    pl, _, _, _ = build_system(rt, oracle, m, 2, rhs=[b])
    s = GMRESSolver(pl, restart, real_update=True)
    A = m.to_scipy()
    res = [np.linalg.norm(b)]
    tid = new_trace_id()
    for cycle in range(3):
        x_before = pl.vector_to_numpy(0, 0, n)
        if traced:
            rt.begin_trace(tid)
        s.step()
        if traced:
            rt.end_trace(tid)
        x = pl.vector_to_numpy(0, 0, n)
        H = s.inner_products
        V = np.stack([pl.vector_to_numpy(2 + j, 0, n) for j in range(restart)], axis=1)
        beta = np.linalg.norm(b - A @ x_before)
        e1 = np.zeros(restart + 1)
        e1[0] = beta
        y, *_ = np.linalg.lstsq(H, e1, rcond=None)
        # (near convergence the Hessenberg is close to breakdown and the update is tiny: compare on the scale of x)
        np.testing.assert_allclose(x - x_before, V @ y, rtol=0, atol=1e-9 * max(np.max(np.abs(V @ y)), 1e-5 * np.max(np.abs(x))))
        res.append(np.linalg.norm(b - A @ x))
        assert res[-1] < 0.9 * res[-2] or res[-1] <= 1e-9 * res[0]
        assert abs(s.residual_norm[-1] - res[-1]) <= 1e-7 * res[0]
    assert res[-1] <= 0.1 * res[0]  # three restart cycles


# ---- several operators on one system, off-diagonal blocks ---------------------------------------
@pytest.mark.parametrize("pieces", [1, 3])
def test_multi_operator_planner_with_off_diagonal_blocks(rt, oracle, pieces):
    """A 2 x 2 block system [[A, B], [C, A + D]] on two index spaces, registered block by block the way the reference's
    SquarePlanner takes (matrix, domain_index, range_index) (src/SquarePlanner.hpp:209-235): two CSR blocks and one COO block
    land on the same range space, so the first overwrites and the others accumulate.  planner.matvec is bit-exact against the
    oracle's planner (which zero-fills and accumulates, like the reference's CPU bodies), and BiCGStab solves the coupled system."""
    from legionsolvers_b200 import solvers as S

    off, val = oracle.benchmark_stencil(3)
    A = oracle.stencil_csr((12, 10, 8), off, val)
    n = A.n_rows
    rng = np.random.default_rng(21)

    def random_csr(density_per_row, scale):
        lens = rng.integers(0, density_per_row + 1, n)
        starts = np.concatenate([[0], np.cumsum(lens)[:-1]])
        col = np.concatenate([np.sort(rng.choice(n, size=l, replace=False)) for l in lens] + [np.zeros(0, dtype=np.int64)]).astype(np.int64)
        rp = np.empty(n, dtype=oracle.RECT_DTYPE)
        rp["lo"], rp["hi"] = starts, starts + lens - 1
        return oracle.Matrix(n, n, scale * rng.standard_normal(col.size), col, rowptr=rp)

    B, Cm, D = random_csr(3, 0.05), random_csr(2, 0.05), random_csr(2, 0.05)
    blocks = [(A, 0, 0), (B, 1, 0), (Cm.to_coo(), 0, 1), (A, 1, 1), (D, 1, 1)]  # (matrix, domain space, range space)
    x = [rng.standard_normal(n), rng.standard_normal(n)]
    pl = S.SquarePlanner(rt)
    opl = oracle.Planner([n, n], [pieces, pieces])
    keep = []
    for sidx in range(2):
        v = S.PartitionedVector(rt, f"sol{sidx}", n, pieces)
        v.zero_fill()
        pl.add_sol_vector(v)
        keep.append(v)
    for sidx in range(2):
        v = S.PartitionedVector(rt, f"rhs{sidx}", n, pieces)
        v.from_numpy(x[sidx])
        pl.add_rhs_vector(v)
        opl.vector(1, sidx)[:] = x[sidx]
        keep.append(v)
    for m, d, r in blocks:
        gm = (S.CSRMatrix.from_host(rt, n, n, m.entry, m.col, m.rowptr) if m.is_csr else S.COOMatrix.from_host(rt, n, n, m.entry, m.row, m.col))
        keep.append(gm)
        pl.add_row_partitioned_matrix(gm, d, r)
        opl.add_matrix(m, d, r)
    pl.allocate_workspace(1)
    opl.allocate_workspace(1)
    pl.matvec(2, 1)
    opl.matvec(2, 1)
    for sidx in range(2):
        got, want = pl.vector_to_numpy(2, sidx, n), opl.vector(2, sidx)
        if sidx == 0:
            np.testing.assert_array_equal(got, want)  # CSR + CSR on one range space: products added one by one, as the CPU bodies do
        else:
            np.testing.assert_allclose(got, want, rtol=0, atol=1e-12 * np.max(np.abs(want)))  # a COO block (atomics) takes part
    # and the coupled system is solvable through the same planner
    pl2, opl2 = S.SquarePlanner(rt), None
    for sidx in range(2):
        v = S.PartitionedVector(rt, f"s{sidx}", n, pieces)
        v.zero_fill()
        pl2.add_sol_vector(v)
        keep.append(v)
    for sidx in range(2):
        v = S.PartitionedVector(rt, f"b{sidx}", n, pieces)
        v.from_numpy(x[sidx])
        pl2.add_rhs_vector(v)
        keep.append(v)
    for gm, (m, d, r) in zip([k for k in keep if isinstance(k, (S.CSRMatrix, S.COOMatrix))], blocks):
        pl2.add_row_partitioned_matrix(gm, d, r)
    sv = S.BiCGStabSolver(pl2)
    for _ in range(45):  # (scipy's BiCGStab needs 37 iterations for 1e-8 on this system)
        sv.step()
    import scipy.sparse as sp

    full = sp.bmat([[A.to_scipy(), B.to_scipy()], [Cm.to_scipy(), A.to_scipy() + D.to_scipy()]]).tocsr()
    sol = np.concatenate([pl2.vector_to_numpy(0, 0, n), pl2.vector_to_numpy(0, 1, n)])
    rhs = np.concatenate(x)
    assert np.linalg.norm(rhs - full @ sol) <= 1e-6 * np.linalg.norm(rhs)
