"""GPU parity of every leaf kernel against the CPU oracle, through the C ABI (include/lsk.h).

Bars: element-wise vector results and the STREAM SpMV are BIT-EXACT (same fma / rounded-product
arithmetic and, for SpMV, the same k-ascending order as the reference CPU bodies); reductions and
the tree-ordered SpMV variants are within 1e-12 relative (north_star tolerance), fp32 within 1e-5.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

REL = 1e-12


@pytest.fixture(scope="module")
def ctx():
    from legionsolvers_b200.kernels import Context

    c = Context()
    yield c
    c.close()


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def scalars(*vals, dtype=torch.float64):
    return [torch.tensor([v], dtype=dtype, device="cuda") for v in vals]


def rel_err(got, want):
    want = np.asarray(want, dtype=np.float64)
    got = np.asarray(got, dtype=np.float64)
    scale = np.max(np.abs(want)) if want.size else 1.0
    return float(np.max(np.abs(got - want)) / (scale if scale > 0 else 1.0)) if want.size else 0.0


# sizes cover: empty, tiny, sub-pack, ragged tails, multi-CTA; offsets cover every 8-byte alignment mod 32
SIZES = [0, 1, 3, 4, 5, 31, 257, 4099, 100_003, 1_048_583]
OFFSETS = [0, 1, 2, 3]


@pytest.mark.parametrize("n", SIZES)
@pytest.mark.parametrize("off", OFFSETS)
def test_blas1_bit_exact(ctx, oracle, n, off):
    rng = np.random.default_rng(n * 7 + off)
    x0, y0 = rng.standard_normal(n), rng.standard_normal(n)
    f = [0.37, -1.9, 0.61, 2.3]
    for nterms in range(0, 5):
        alpha = oracle.get_alpha(f[:nterms])
        terms = scalars(*f[:nterms])
        for name in ("scal", "axpy", "xpay"):
            xw, yw = x0.copy(), y0.copy()
            buf_x = torch.zeros(n + 8, dtype=torch.float64, device="cuda")
            buf_y = torch.zeros(n + 8, dtype=torch.float64, device="cuda")
            xd, yd = buf_x[off:off + n], buf_y[off:off + n]
            xd.copy_(dev(x0)); yd.copy_(dev(y0))
            if name == "scal":
                oracle.scal(alpha, xw); ctx.scal(terms, xd)
                np.testing.assert_array_equal(xd.cpu().numpy(), xw)
            elif name == "axpy":
                oracle.axpy(alpha, xw, yw); ctx.axpy(terms, xd, yd)
                np.testing.assert_array_equal(yd.cpu().numpy(), yw)
            else:
                oracle.xpay(alpha, xw, yw); ctx.xpay(terms, xd, yd)
                np.testing.assert_array_equal(yd.cpu().numpy(), yw)
            # guard elements around the slice are untouched
            assert float(buf_y[:off].abs().sum()) == 0.0 and float(buf_y[off + n:].abs().sum()) == 0.0
            assert float(buf_x[:off].abs().sum()) == 0.0 and float(buf_x[off + n:].abs().sum()) == 0.0


def test_blas1_mixed_alignment_falls_back_to_scalar_path(ctx, oracle):
    n = 10_007
    rng = np.random.default_rng(3)
    x0, y0 = rng.standard_normal(n), rng.standard_normal(n)
    bx = torch.zeros(n + 8, dtype=torch.float64, device="cuda")
    by = torch.zeros(n + 8, dtype=torch.float64, device="cuda")
    xd, yd = bx[1:1 + n], by[2:2 + n]  # different residues mod 32 bytes
    xd.copy_(dev(x0)); yd.copy_(dev(y0))
    a = scalars(0.77)
    ctx.axpy(a, xd, yd)
    yw = y0.copy(); oracle.axpy(0.77, x0, yw)
    np.testing.assert_array_equal(yd.cpu().numpy(), yw)
    out = torch.zeros(1, dtype=torch.float64, device="cuda")
    ctx.dot(xd, yd, out)
    assert abs(out.item() - oracle.dot(x0, yw)) <= REL * abs(oracle.dot(np.abs(x0), np.abs(yw)))


@pytest.mark.parametrize("n", SIZES)
def test_dot_and_fill(ctx, oracle, n):
    rng = np.random.default_rng(n + 11)
    v, w = rng.standard_normal(n), rng.standard_normal(n)
    out = torch.full((1,), 123.0, dtype=torch.float64, device="cuda")
    ctx.dot(dev(v), dev(w), out)
    want = oracle.dot(v, w)
    bound = REL * max(1e-300, float(np.dot(np.abs(v), np.abs(w))))
    assert abs(out.item() - want) <= bound
    # r.r of a positive vector: relative to the value itself
    ctx.dot(dev(v), dev(v), out)
    assert abs(out.item() - oracle.dot(v, v)) <= REL * max(oracle.dot(v, v), 1e-300)
    # deterministic: same launch twice gives the same bits
    a = torch.zeros(1, dtype=torch.float64, device="cuda"); b = torch.zeros_like(a)
    ctx.dot(dev(v), dev(w), a); ctx.dot(dev(v), dev(w), b)
    assert a.item() == b.item()
    x = torch.zeros(n + 3, dtype=torch.float64, device="cuda")
    ctx.fill(x[1:1 + n], 2.5)
    assert torch.all(x[1:1 + n] == 2.5) and x[0] == 0 and torch.all(x[1 + n:] == 0)
    s = scalars(-4.25)[0]
    ctx.fill(x[1:1 + n], s)
    assert torch.all(x[1:1 + n] == -4.25)


def test_blas1_f32(ctx, oracle):
    n = 50_021
    rng = np.random.default_rng(5)
    x0 = rng.standard_normal(n).astype(np.float32)
    y0 = rng.standard_normal(n).astype(np.float32)
    t = scalars(1.5, 0.25, dtype=torch.float32)
    alpha = np.float32(1.5) / np.float32(0.25)
    xd, yd = dev(x0), dev(y0)
    ctx.axpy(t, xd, yd)
    yw = y0.copy(); oracle.axpy_f32(alpha, x0, yw)
    np.testing.assert_array_equal(yd.cpu().numpy(), yw)
    ctx.xpay(t, xd, yd)
    oracle.xpay_f32(alpha, x0, yw)
    np.testing.assert_array_equal(yd.cpu().numpy(), yw)
    ctx.scal(t, yd)
    oracle.scal_f32(alpha, yw)
    np.testing.assert_array_equal(yd.cpu().numpy(), yw)
    out = torch.zeros(1, dtype=torch.float32, device="cuda")
    ctx.dot(xd, yd, out)
    assert abs(out.item() - float(np.dot(x0.astype(np.float64), yw.astype(np.float64)))) <= 1e-5 * float(
        np.dot(np.abs(x0), np.abs(yw)))


def test_test02_chain_prints_zero(ctx):
    """test/Test02VectorOperations.cpp:128-137 through the GPU kernels, 10 pieces of 10."""
    u = torch.zeros(100, dtype=torch.float64, device="cuda")
    v = torch.zeros_like(u); w = torch.zeros_like(u)
    one, neg_one, neg_half = scalars(1.0, -1.0, -0.5)
    pieces = [slice(10 * c, 10 * c + 10) for c in range(10)]
    for p in pieces:
        ctx.fill(u[p], 1.5); ctx.fill(v[p], 2.7); ctx.copy(u[p], w[p])
    for p in pieces:
        ctx.axpy([one], v[p], w[p])
    for p in pieces:
        ctx.xpay([neg_one], u[p], v[p])
    for p in pieces:
        ctx.axpy([neg_half], v[p], u[p])
    for p in pieces:
        ctx.axpy([neg_half], w[p], u[p])
    parts = torch.zeros(10, dtype=torch.float64, device="cuda")
    for c, p in enumerate(pieces):
        ctx.dot(u[p], u[p], parts[c:c + 1])
    assert parts.sum().item() == 0.0


def test_scalar_ops_test01_chain(ctx, oracle):
    """test/Test01ScalarOperations.cpp:17-32: ((2+10)/(2+2)) - 2 == 1.0, f64 and f32."""
    from legionsolvers_b200 import kernels as K

    for dt in (torch.float64, torch.float32):
        two, ten = scalars(2.0, 10.0, dtype=dt)
        t12, t4, t3, t1 = (torch.zeros(1, dtype=dt, device="cuda") for _ in range(4))
        ctx.scalar_op(K.OP_ADD, two, ten, t12)
        ctx.scalar_op(K.OP_ADD, two, two, t4)
        ctx.scalar_op(K.OP_DIV, t12, t4, t3)
        ctx.scalar_op(K.OP_SUB, t3, two, t1)
        assert t1.item() == 1.0
    a, b = scalars(0.3, 7.7)
    o = torch.zeros(1, dtype=torch.float64, device="cuda")
    for op, name in ((K.OP_ADD, "add"), (K.OP_SUB, "sub"), (K.OP_MUL, "mul"), (K.OP_DIV, "div")):
        ctx.scalar_op(op, a, b, o)
        assert o.item() == oracle.scalar(name, 0.3, 7.7)
    for op, name in ((K.OP_NEG, "neg"), (K.OP_SQRT, "sqrt"), (K.OP_RSQRT, "rsqrt")):
        ctx.scalar_op(op, b, None, o)
        assert o.item() == oracle.scalar(name, 7.7)
    ctx.scalar_op(K.OP_DUMMY, None, None, o)
    assert o.item() == 1.0


# ---------------------------------------------------------------------------------------------------
# SpMV
# ---------------------------------------------------------------------------------------------------
STENCILS = [
    (1, (100,)), (1, (4099,)),
    (2, (256, 256)), (2, (37, 53)),
    (3, (32, 32, 32)), (3, (64, 64, 64)), (3, (7, 9, 11)),
    (4, (32, 32, 32)), (4, (48, 48, 48)), (4, (5, 6, 7)),
]


def ramp(n):
    """x_i = ((i * 2654435761) mod 2^32) / 2^32 - 0.5  (SURVEY.md section 8d, config C2)."""
    i = np.arange(n, dtype=np.uint64)
    return ((i * np.uint64(2654435761)) % np.uint64(2 ** 32)).astype(np.float64) / 2.0 ** 32 - 0.5


@pytest.mark.parametrize("dim_flag,shape", STENCILS)
def test_csr_spmv_whole_matrix(ctx, oracle, dim_flag, shape):
    from legionsolvers_b200 import kernels as K

    off, val = oracle.benchmark_stencil(dim_flag)
    m = oracle.stencil_csr(shape, off, val)
    x = ramp(m.n_cols)
    want = np.zeros(m.n_rows)
    oracle.csr_matvec(m, x, want)
    entry, col, rowptr, xd = dev(m.entry), dev(m.col), K.rect_tensor(m.rowptr), dev(x)
    absax = np.zeros(m.n_rows)
    oracle.csr_matvec(oracle.Matrix(m.n_rows, m.n_cols, np.abs(m.entry), m.col, rowptr=m.rowptr), np.abs(x), absax)
    # AUTO resolves to the thread-per-row STREAM kernel up to 12 non-zeros per row on average (5- / 7-point stencils,
    # 1-D Laplacian) and to LANES (2-8 lanes per row, tree order) for the 27-point stencil
    exact = (K.SPMV_STREAM,) + ((K.SPMV_AUTO,) if m.nnz <= 12 * m.n_rows else ())
    for variant in (K.SPMV_AUTO, K.SPMV_STREAM, K.SPMV_LANES, K.SPMV_VECTOR, K.SPMV_WARP):
        y = torch.full((m.n_rows,), 7.0, dtype=torch.float64, device="cuda")  # beta = 0: overwritten
        ctx.csr_spmv(m.n_rows, m.nnz, entry, col, rowptr, 0, xd, 0, y, variant=variant)
        got = y.cpu().numpy()
        if variant in exact:
            np.testing.assert_array_equal(got, want)  # same order, same rounding as the CPU body
        else:
            assert np.all(np.abs(got - want) <= REL * np.maximum(absax, 1e-300))


@pytest.mark.parametrize("pieces", [2, 3, 4, 8])
@pytest.mark.parametrize("dim_flag,shape", [(2, (64, 48)), (3, (16, 20, 24)), (4, (12, 16, 20)), (1, (100,))])
def test_csr_spmv_pieces_with_fused_dots(ctx, oracle, pieces, dim_flag, shape):
    """Row-partitioned launch exactly as the planner issues it: per piece, the kernel sub-range
    (odd k offsets -> every alignment case), the ghost window of x passed as a shifted pointer,
    fused y.w and y.y partials."""
    from legionsolvers_b200 import kernels as K

    off, val = oracle.benchmark_stencil(dim_flag)
    m = oracle.stencil_csr(shape, off, val)
    n = m.n_rows
    x = ramp(n) + 0.25
    w = ramp(n)[::-1].copy()
    pl = oracle.Planner([n], [pieces])
    b = pl.add_matrix(m)
    want = np.zeros(n)
    oracle.csr_matvec(m, x, want)
    entry, col, rowptr, wd = dev(m.entry), dev(m.col), K.rect_tensor(m.rowptr), dev(w)
    y = torch.zeros(n, dtype=torch.float64, device="cuda")
    for variant in (K.SPMV_STREAM, K.SPMV_LANES, K.SPMV_VECTOR):
        y.zero_()
        for c in range(pieces):
            r_lo, r_hi = pl.piece_bounds(0, c)
            k_lo, k_hi = pl.kernel_bounds(b, c)
            g_lo, g_hi = pl.ghost_bounds(b, c)
            rows, nnz = r_hi - r_lo + 1, k_hi - k_lo + 1
            ghost = dev(x[g_lo:g_hi + 1])  # only the ghost window exists on the device
            d1 = torch.zeros(1, dtype=torch.float64, device="cuda"); d2 = torch.zeros_like(d1)
            ctx.csr_spmv(rows, nnz, entry[k_lo:k_hi + 1], col[k_lo:k_hi + 1], rowptr[r_lo:r_hi + 1], k_lo,
                         ghost, g_lo, y[r_lo:r_hi + 1], dot_w=wd[r_lo:r_hi + 1], dot_out=d1, dot_yy_out=d2,
                         variant=variant)
            yp = want[r_lo:r_hi + 1]
            assert abs(d1.item() - float(np.dot(yp, w[r_lo:r_hi + 1]))) <= 1e-11 * float(np.dot(np.abs(yp), np.abs(w[r_lo:r_hi + 1])) + 1e-300)
            assert abs(d2.item() - float(np.dot(yp, yp))) <= 1e-11 * float(np.dot(yp, yp) + 1e-300)
            # y.y alone (y poisoned first: the kernel must use the row's NEW result, not what y held before)
            d3 = torch.zeros_like(d1)
            y[r_lo:r_hi + 1] = 123.0
            ctx.csr_spmv(rows, nnz, entry[k_lo:k_hi + 1], col[k_lo:k_hi + 1], rowptr[r_lo:r_hi + 1], k_lo,
                         ghost, g_lo, y[r_lo:r_hi + 1], dot_yy_out=d3, variant=variant)
            assert d3.item() == d2.item()
        got = y.cpu().numpy()
        if variant == K.SPMV_STREAM:
            np.testing.assert_array_equal(got, want)
        else:
            np.testing.assert_allclose(got, want, rtol=0, atol=REL * np.max(np.abs(want)) * 8)


def test_csr_spmv_ragged_rows_and_empty_rows(ctx, oracle):
    """Rows of wildly different length (0 .. 5000) exercise multi-tile rows, empty rows and the
    tile edges; entry/col sub-arrays start at odd offsets."""
    from legionsolvers_b200 import kernels as K

    rng = np.random.default_rng(42)
    n_rows, n_cols = 700, 9000
    lens = rng.integers(0, 12, n_rows)
    lens[[5, 300, 301, 650]] = [5000, 2049, 2047, 4096]
    lens[[0, 10, 11, 12, 699]] = 0
    nnz = int(lens.sum())
    col = np.concatenate([np.sort(rng.choice(n_cols, size=l, replace=False)) for l in lens]).astype(np.int64)
    entry = rng.standard_normal(nnz)
    rowptr = np.empty(n_rows, dtype=oracle.RECT_DTYPE)
    starts = np.concatenate([[0], np.cumsum(lens)[:-1]])
    rowptr["lo"], rowptr["hi"] = starts, starts + lens - 1
    m = oracle.Matrix(n_rows, n_cols, entry, col, rowptr=rowptr)
    x = rng.standard_normal(n_cols)
    want = np.zeros(n_rows)
    for r in range(n_rows):  # plain sequential row sums (orc_csr_matvec needs every k covered: it is)
        for k in range(rowptr["lo"][r], rowptr["hi"][r] + 1):
            want[r] += entry[k] * x[col[k]]
    absrow = np.array([np.sum(np.abs(entry[rowptr["lo"][r]:rowptr["hi"][r] + 1] * x[col[rowptr["lo"][r]:rowptr["hi"][r] + 1]])) for r in range(n_rows)])
    for pad in (0, 1, 2, 3):
        e_buf = torch.zeros(nnz + 8, dtype=torch.float64, device="cuda")
        c_buf = torch.zeros(nnz + 8, dtype=torch.int64, device="cuda")
        e_buf[pad:pad + nnz] = dev(entry); c_buf[pad:pad + nnz] = dev(col)
        for variant in (K.SPMV_STREAM, K.SPMV_LANES, K.SPMV_VECTOR, K.SPMV_WARP):
            y = torch.full((n_rows,), -3.0, dtype=torch.float64, device="cuda")
            ctx.csr_spmv(n_rows, nnz, e_buf[pad:pad + nnz], c_buf[pad:pad + nnz], K.rect_tensor(rowptr), 0,
                         dev(x), 0, y, variant=variant)
            got = y.cpu().numpy()
            if variant == K.SPMV_STREAM:
                np.testing.assert_array_equal(got, want)
            else:
                assert np.all(np.abs(got - want) <= REL * np.maximum(absrow, 1e-300))
    # mismatched alignment between entry and col -> scalar-load path, still bit exact
    e_buf = torch.zeros(nnz + 8, dtype=torch.float64, device="cuda")
    c_buf = torch.zeros(nnz + 8, dtype=torch.int64, device="cuda")
    e_buf[1:1 + nnz] = dev(entry); c_buf[2:2 + nnz] = dev(col)
    y = torch.zeros(n_rows, dtype=torch.float64, device="cuda")
    ctx.csr_spmv(n_rows, nnz, e_buf[1:1 + nnz], c_buf[2:2 + nnz], K.rect_tensor(rowptr), 0, dev(x), 0, y,
                 variant=K.SPMV_STREAM)
    np.testing.assert_array_equal(y.cpu().numpy(), want)


@pytest.mark.parametrize("order", ["reversed", "shuffled", "block_shuffled"])
def test_csr_spmv_rows_stored_out_of_order(ctx, oracle, order):
    """The reference's rowptr is a field of arbitrary Rect<1> (src/CSRMatrix.hpp:24-26): nothing forces row r + 1 to be
    stored after row r.  The warp-specialised kernel guesses a row block's run of non-zeros from its end rows and
    must still produce the reference result -- bit for bit -- when rows are stored in any order."""
    from legionsolvers_b200 import kernels as K

    rng = np.random.default_rng(7)
    n_rows, n_cols = 3000, 3000
    lens = rng.integers(0, 15, n_rows)
    lens[[17, 1500]] = [900, 2600]  # a few long rows as well (multi-tile)
    perm = {"reversed": np.arange(n_rows)[::-1], "shuffled": rng.permutation(n_rows),
            "block_shuffled": np.concatenate([b for b in rng.permutation(np.array_split(np.arange(n_rows), 40))])}[order]
    starts = np.zeros(n_rows, dtype=np.int64)
    pos = 0
    for r in perm:  # storage order
        starts[r] = pos
        pos += lens[r]
    nnz = int(pos)
    col = np.zeros(nnz, dtype=np.int64)
    entry = rng.standard_normal(nnz)
    for r in range(n_rows):
        col[starts[r]:starts[r] + lens[r]] = np.sort(rng.choice(n_cols, size=lens[r], replace=False))
    rowptr = np.empty(n_rows, dtype=oracle.RECT_DTYPE)
    rowptr["lo"], rowptr["hi"] = starts, starts + lens - 1
    x = rng.standard_normal(n_cols)
    want = np.zeros(n_rows)
    for r in range(n_rows):
        for k in range(rowptr["lo"][r], rowptr["hi"][r] + 1):
            want[r] += entry[k] * x[col[k]]
    absrow = np.array([np.sum(np.abs(entry[starts[r]:starts[r] + lens[r]] * x[col[starts[r]:starts[r] + lens[r]]])) for r in range(n_rows)])
    w = rng.standard_normal(n_rows)
    for variant in (K.SPMV_STREAM, K.SPMV_LANES, K.SPMV_AUTO):
        y = torch.full((n_rows,), 5.0, dtype=torch.float64, device="cuda")
        d = torch.zeros(1, dtype=torch.float64, device="cuda")
        ctx.csr_spmv(n_rows, nnz, dev(entry), dev(col), K.rect_tensor(rowptr), 0, dev(x), 0, y, dot_w=dev(w), dot_out=d, variant=variant)
        got = y.cpu().numpy()
        if variant != K.SPMV_LANES:
            np.testing.assert_array_equal(got, want)
        else:
            assert np.all(np.abs(got - want) <= REL * np.maximum(absrow, 1e-300))
        assert abs(d.item() - float(want @ w)) <= 1e-11 * float(np.abs(want) @ np.abs(w))


def test_csr_spmv_zero_rows_and_f32(ctx, oracle):
    from legionsolvers_b200 import kernels as K

    e = torch.zeros(0, dtype=torch.float64, device="cuda")
    c = torch.zeros(0, dtype=torch.int64, device="cuda")
    rp = torch.zeros((0, 2), dtype=torch.int64, device="cuda")
    y = torch.zeros(0, dtype=torch.float64, device="cuda")
    x = torch.zeros(4, dtype=torch.float64, device="cuda")
    ctx.csr_spmv(0, 0, e, c, rp, 0, x, 0, y)  # no-op, must not fail
    d = torch.full((1,), 9.0, dtype=torch.float64, device="cuda")
    ctx.csr_spmv(0, 0, e, c, rp, 0, x, 0, y, dot_yy_out=d)
    assert d.item() == 0.0
    off, val = oracle.benchmark_stencil(3)
    m = oracle.stencil_csr((20, 20, 20), off, val)
    x = ramp(m.n_cols).astype(np.float32)
    want = m.to_scipy().astype(np.float64) @ x.astype(np.float64)
    for variant in (K.SPMV_STREAM, K.SPMV_VECTOR, K.SPMV_WARP):
        y = torch.zeros(m.n_rows, dtype=torch.float32, device="cuda")
        ctx.csr_spmv(m.n_rows, m.nnz, dev(m.entry.astype(np.float32)), dev(m.col), K.rect_tensor(m.rowptr), 0,
                     dev(x), 0, y, variant=variant)
        np.testing.assert_allclose(y.cpu().numpy(), want, rtol=0, atol=1e-5 * np.max(np.abs(want)))


@pytest.mark.parametrize("sort", [True, False])
def test_coo_spmv(ctx, oracle, sort):
    """COO segmented reduction: accumulates (beta = 1), honours the row/col guards, any ordering."""
    off, val = oracle.benchmark_stencil(4)
    m = oracle.stencil_coo((14, 15, 16), off, val)
    n = m.n_rows
    rng = np.random.default_rng(9)
    perm = np.arange(m.nnz) if sort else rng.permutation(m.nnz)
    entry, row, col = m.entry[perm], m.row[perm], m.col[perm]
    x = ramp(n)
    y0 = rng.standard_normal(n)
    for (r_lo, r_hi, c_lo, c_hi) in [(0, n - 1, 0, n - 1), (n // 3, 2 * n // 3, 0, n - 1), (0, n - 1, 100, n - 200)]:
        want = y0.copy()
        oracle.coo_matvec(oracle.Matrix(n, n, entry, col, row=row), x, want, r=(r_lo, r_hi), cols=(c_lo, c_hi))
        y = dev(y0)
        ctx.coo_spmv(m.nnz, dev(entry), dev(row), dev(col), dev(x), 0, y, 0, (r_lo, r_hi), (c_lo, c_hi))
        np.testing.assert_allclose(y.cpu().numpy(), want, rtol=0, atol=REL * 40)
    # piece launch with shifted pointers: rows [r_lo, r_hi] of y, ghost window of x
    r_lo, r_hi = n // 4, n // 2
    sel = (m.row >= r_lo) & (m.row <= r_hi)
    e_p, r_p, c_p = m.entry[sel], m.row[sel], m.col[sel]
    g_lo, g_hi = int(c_p.min()), int(c_p.max())
    want = np.zeros(n)
    oracle.coo_matvec(oracle.Matrix(n, n, e_p, c_p, row=r_p), x, want, r=(r_lo, r_hi), cols=(g_lo, g_hi))
    yp = torch.zeros(r_hi - r_lo + 1, dtype=torch.float64, device="cuda")
    ctx.coo_spmv(e_p.size, dev(e_p), dev(r_p), dev(c_p), dev(x[g_lo:g_hi + 1]), g_lo, yp, r_lo, (r_lo, r_hi), (g_lo, g_hi))
    np.testing.assert_allclose(yp.cpu().numpy(), want[r_lo:r_hi + 1], rtol=0, atol=REL * 40)


def test_coo_power_law_rows(ctx, oracle):
    rng = np.random.default_rng(12345)
    n = 20_000
    lens = np.minimum(3000, np.floor(8 * rng.random(n) ** (-1 / 1.5))).astype(np.int64)
    row = np.repeat(np.arange(n, dtype=np.int64), lens)
    col = rng.integers(0, n, row.size).astype(np.int64)
    entry = rng.uniform(-1, 1, row.size)
    x = rng.standard_normal(n)
    want = np.zeros(n)
    oracle.coo_matvec(oracle.Matrix(n, n, entry, col, row=row), x, want)
    absw = np.zeros(n)
    oracle.coo_matvec(oracle.Matrix(n, n, np.abs(entry), col, row=row), np.abs(x), absw)
    y = torch.zeros(n, dtype=torch.float64, device="cuda")
    ctx.coo_spmv(row.size, dev(entry), dev(row), dev(col), dev(x), 0, y, 0, (0, n - 1), (0, n - 1))
    assert np.all(np.abs(y.cpu().numpy() - want) <= REL * np.maximum(absw, 1e-300))


# ---------------------------------------------------------------------------------------------------
# fused passes == the sequence of leaf tasks they replace
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [1, 1000, 262_147])
def test_fused_passes_match_leaf_sequences(ctx, oracle, n):
    rng = np.random.default_rng(n)
    p, q, x, r, u, rt, v = (rng.standard_normal(n) for _ in range(7))
    rr_old, pq = 3.7, 1.9
    # cg_update
    xw, rw = x.copy(), r.copy()
    oracle.axpy(oracle.get_alpha([rr_old, pq]), p, xw)
    oracle.axpy(oracle.get_alpha([-1.0, rr_old, pq]), q, rw)
    xd, rd = dev(x), dev(r)
    out = torch.zeros(1, dtype=torch.float64, device="cuda")
    ctx.cg_update(*scalars(rr_old, pq), dev(p), dev(q), xd, rd, out)
    np.testing.assert_array_equal(xd.cpu().numpy(), xw)
    np.testing.assert_array_equal(rd.cpu().numpy(), rw)
    assert abs(out.item() - oracle.dot(rw, rw)) <= REL * oracle.dot(rw, rw)
    # axpy_dot, with and without aliasing w = y
    h = 0.83
    yw = x.copy(); oracle.axpy(oracle.get_alpha([-1.0, h, 1.0]), p, yw)
    yd = dev(x)
    ctx.axpy_dot(scalars(-1.0, h, 1.0), dev(p), yd, dev(q), out)
    np.testing.assert_array_equal(yd.cpu().numpy(), yw)
    assert abs(out.item() - oracle.dot(yw, q)) <= REL * float(np.dot(np.abs(yw), np.abs(q)))
    yd = dev(x)
    ctx.axpy_dot(scalars(-1.0, h, 1.0), dev(p), yd, yd, out)
    assert abs(out.item() - oracle.dot(yw, yw)) <= REL * oracle.dot(yw, yw)
    # dot2
    o1, o2 = torch.zeros(1, dtype=torch.float64, device="cuda"), torch.zeros(1, dtype=torch.float64, device="cuda")
    ctx.dot2(dev(r), dev(u), o1, o2)
    assert abs(o1.item() - oracle.dot(r, u)) <= REL * float(np.dot(np.abs(r), np.abs(u)))
    assert abs(o2.item() - oracle.dot(u, u)) <= REL * oracle.dot(u, u)
    # bicg_p_update: P += (-omega) V ; P = beta P + R
    rho_new, rho_old, alpha, omega = 1.3, 0.9, 0.41, 0.77
    beta = oracle.scalar("mul", oracle.scalar("div", rho_new, rho_old), oracle.scalar("div", alpha, omega))
    pw = p.copy()
    oracle.axpy(-omega, v, pw); oracle.xpay(beta, r, pw)
    pd = dev(p)
    ctx.bicg_p_update(*scalars(rho_new, rho_old, alpha, omega), dev(v), dev(r), pd)
    np.testing.assert_array_equal(pd.cpu().numpy(), pw)
    # bicg_tail
    ru, uu = 2.2, 5.1
    om = oracle.scalar("div", ru, uu)
    xw, rw = x.copy(), r.copy()
    oracle.axpy(alpha, p, xw); oracle.axpy(om, rw, xw); oracle.axpy(-om, u, rw)
    xd, rd = dev(x), dev(r)
    ctx.bicg_tail(*scalars(alpha, ru, uu), dev(p), dev(u), dev(rt), xd, rd, out)
    np.testing.assert_array_equal(xd.cpu().numpy(), xw)
    np.testing.assert_array_equal(rd.cpu().numpy(), rw)
    assert abs(out.item() - oracle.dot(rw, rt)) <= REL * float(np.dot(np.abs(rw), np.abs(rt)))


@pytest.mark.parametrize("n", [40, 1000, 16_384 + 12, 262_147, 6_291_456 + 36])
@pytest.mark.parametrize("off", [0, 1, 3])
def test_cg_update_and_direction_tma_streamed(ctx, oracle, n, off):
    """The two vector passes of the fused CG step (TMA-streamed, dynamically scheduled; the x/r update above 6 M elements):
    element-wise results BIT-EXACT with the leaf sequences axpy/axpy and xpay, r.r within 1e-12, history appended,
    rr_cur advanced; every 8-byte alignment residue (ragged head/tail around the 32-byte body)."""
    rng = np.random.default_rng(n + off)
    p0, q0, x0, r0 = (rng.standard_normal(n) for _ in range(4))
    buf = lambda a: (lambda t: (t.copy_(dev(a)), t)[1])(torch.zeros(n + 8, dtype=torch.float64, device="cuda")[off:off + n])  # noqa: E731
    p, q, x, r = buf(p0), buf(q0), buf(x0), buf(r0)
    rr_old, pq = 3.7, 1.9
    xw, rw = x0.copy(), r0.copy()
    oracle.axpy(oracle.get_alpha([rr_old, pq]), p0, xw)
    oracle.axpy(oracle.get_alpha([-1.0, rr_old, pq]), q0, rw)
    rr_cur, pqd = scalars(rr_old, pq)
    rr_new = torch.zeros(1, dtype=torch.float64, device="cuda")
    ctx.cg_update(rr_cur, pqd, p, q, x, r, rr_new)
    np.testing.assert_array_equal(x.cpu().numpy(), xw)
    np.testing.assert_array_equal(r.cpu().numpy(), rw)
    want_rr = oracle.dot(rw, rw)
    assert abs(rr_new.item() - want_rr) <= REL * want_rr
    # direction: p = fma(rr_new / rr_cur, p, r), history.push_back(rr_new), rr_cur <- rr_new
    hist = torch.zeros(4, dtype=torch.float64, device="cuda")
    count = torch.tensor([5], dtype=torch.int64, device="cuda")  # circular: slot 5 % 4 = 1
    got_rr = rr_new.item()
    pw = p0.copy()
    oracle.xpay(oracle.get_alpha([got_rr, rr_old]), rw, pw)
    ctx.cg_direction(rr_cur, rr_new, r, p, hist, count)
    np.testing.assert_array_equal(p.cpu().numpy(), pw)
    assert hist.cpu().numpy().tolist() == [0.0, got_rr, 0.0, 0.0] and int(count.item()) == 6
    assert rr_cur.item() == got_rr
    # a second pair of launches reuses the (reset) work counters
    ctx.cg_update(rr_cur, pqd, p, q, x, r, rr_new)
    oracle.axpy(oracle.get_alpha([got_rr, pq]), pw, xw)
    oracle.axpy(oracle.get_alpha([-1.0, got_rr, pq]), q0, rw)
    np.testing.assert_array_equal(x.cpu().numpy(), xw)
    np.testing.assert_array_equal(r.cpu().numpy(), rw)


def test_cg_golden_history_through_leaf_kernels(ctx, oracle):
    """Test06CSRSolveCG (n=100, 4 pieces, 10 steps) driven through the C ABI kernel by kernel,
    scalars never leaving the device: reproduces the reference's golden residual history."""
    from legionsolvers_b200 import kernels as K

    n, P, its = 100, 4, 10
    m = oracle.laplacian_1d_csr(n)
    pl = oracle.Planner([n], [P])
    b = pl.add_matrix(m)
    entry, col, rowptr = dev(m.entry), dev(m.col), K.rect_tensor(m.rowptr)
    z = lambda: torch.zeros(n, dtype=torch.float64, device="cuda")  # noqa: E731
    sol, rhs, p, q, r = z(), z(), z(), z(), z()
    ctx.fill(rhs, 1.0)
    ctx.copy(rhs, p); ctx.copy(rhs, r)
    hist = torch.zeros(its + 1, dtype=torch.float64, device="cuda")
    part = torch.zeros(P, dtype=torch.float64, device="cuda")
    pq = torch.zeros(1, dtype=torch.float64, device="cuda")

    def reduce_parts(dst):  # Legion's future-map sum, colour order
        ctx.scalar_op(K.OP_COPY, part[0:1], None, dst)
        for c in range(1, P):
            ctx.scalar_op(K.OP_ADD, dst, part[c:c + 1], dst)

    bounds = [pl.piece_bounds(0, c) for c in range(P)]
    for c, (lo, hi) in enumerate(bounds):
        ctx.dot(r[lo:hi + 1], r[lo:hi + 1], part[c:c + 1])
    reduce_parts(hist[0:1])
    neg_one = scalars(-1.0)[0]
    for it in range(its):
        for c, (lo, hi) in enumerate(bounds):
            k_lo, k_hi = pl.kernel_bounds(b, c)
            ctx.csr_spmv(hi - lo + 1, k_hi - k_lo + 1, entry[k_lo:k_hi + 1], col[k_lo:k_hi + 1], rowptr[lo:hi + 1],
                         k_lo, p, 0, q[lo:hi + 1], dot_w=p[lo:hi + 1], dot_out=part[c:c + 1])
        reduce_parts(pq)
        rr_old, rr_new = hist[it:it + 1], hist[it + 1:it + 2]
        for c, (lo, hi) in enumerate(bounds):
            ctx.axpy([rr_old, pq], p[lo:hi + 1], sol[lo:hi + 1])
            ctx.axpy([neg_one, rr_old, pq], q[lo:hi + 1], r[lo:hi + 1])
            ctx.dot(r[lo:hi + 1], r[lo:hi + 1], part[c:c + 1])
        reduce_parts(rr_new)
        for c, (lo, hi) in enumerate(bounds):
            ctx.xpay([rr_new, rr_old], r[lo:hi + 1], p[lo:hi + 1])
    got = hist.cpu().numpy()
    assert list(got) == [100.0, 4900.0, 4704.0, 4512.0, 4324.0, 4140.0, 3960.0, 3784.0, 3612.0, 3444.0, 3280.0]
    assert ctx.launch_count > 0


# ---- the fused CG step as its three leaf launches (mat-vec + p.q, x/r update + r.r, direction) ----------------------------
def _cg_state(oracle, m, rhs_val=1.0, cap=64):
    from legionsolvers_b200 import kernels as K

    n = m.n_rows
    z = lambda: torch.zeros(n, dtype=torch.float64, device="cuda")  # noqa: E731
    st = dict(entry=dev(m.entry), col=dev(m.col), rowptr=K.rect_tensor(m.rowptr), x=z(), r=z(), p=z(), q=z(),
              rr_cur=torch.zeros(1, dtype=torch.float64, device="cuda"), rr_new=torch.zeros(1, dtype=torch.float64, device="cuda"),
              pq=torch.zeros(1, dtype=torch.float64, device="cuda"), hist=torch.zeros(cap, dtype=torch.float64, device="cuda"),
              count=torch.zeros(1, dtype=torch.int64, device="cuda"))
    st["r"].fill_(rhs_val); st["p"].fill_(rhs_val)
    st["rr_cur"].fill_(float(n) * rhs_val * rhs_val)
    return st


def _cg_steps(ctx, st, niter):
    from legionsolvers_b200 import _abi
    from legionsolvers_b200 import kernels as K

    n, nnz = st["q"].numel(), st["entry"].numel()
    streamed = bool(_abi.lib().lsk_cg_direction_supported(n, st["r"].data_ptr(), st["p"].data_ptr()))
    for _ in range(niter):
        ctx.csr_spmv(n, nnz, st["entry"], st["col"], st["rowptr"], 0, st["p"], 0, st["q"], dot_w=st["p"], dot_out=st["pq"])
        ctx.cg_update(st["rr_cur"], st["pq"], st["p"], st["q"], st["x"], st["r"], st["rr_new"])
        if streamed:
            ctx.cg_direction(st["rr_cur"], st["rr_new"], st["r"], st["p"], st["hist"], st["count"])
        else:  # r and p not 32-byte congruent: the reference's xpay + push_back
            ctx.xpay([st["rr_new"], st["rr_cur"]], st["r"], st["p"])
            c = int(st["count"].item())
            st["hist"][c % st["hist"].numel()] = st["rr_new"][0]
            st["count"] += 1
            ctx.scalar_op(K.OP_COPY, st["rr_new"], None, st["rr_cur"])


def test_cg_steps_golden_history(ctx, oracle):
    """Test06CSRSolveCG's system (1-D Laplacian n = 100) through the three fused leaf kernels: the reference's
    golden residual history (exact integers), whatever the batching of the iterations."""
    m = oracle.laplacian_1d_csr(100)
    want = [4900.0, 4704.0, 4512.0, 4324.0, 4140.0, 3960.0, 3784.0, 3612.0, 3444.0, 3280.0]
    for split in ([10], [1] * 10, [3, 7]):
        st = _cg_state(oracle, m)
        for k in split:
            _cg_steps(ctx, st, k)
        torch.cuda.synchronize()
        assert int(st["count"].item()) == 10
        assert list(st["hist"][:10].cpu().numpy()) == want
        assert st["rr_cur"].item() == want[-1] and st["rr_new"].item() == want[-1]


@pytest.mark.parametrize("dim_flag,shape,its", [(2, (256, 256), 60), (3, (40, 40, 40), 50), (4, (24, 24, 24), 30), (3, (96, 96, 96), 25)])
def test_cg_steps_vs_oracle_and_leaf_sequence(ctx, oracle, dim_flag, shape, its):
    """Residual history within 1e-10 of the oracle's CG (north_star tolerance); vectors agree with the
    oracle to 1e-10 and p.q is the last p.Ap.  Grids above one wave of CTAs exercise the multi-block path."""
    off, val = oracle.benchmark_stencil(dim_flag)
    m = oracle.stencil_csr(shape, off, val)
    n = m.n_rows
    st = _cg_state(oracle, m, cap=its)
    _cg_steps(ctx, st, its)
    torch.cuda.synchronize()
    opl = oracle.Planner([n], [1]); opl.fill(1, 1.0); opl.add_matrix(m)
    ocg = oracle.CGSolver(opl)
    for _ in range(its):
        ocg.step()
    want = ocg.residual_norm_squared[1:]
    got = st["hist"].cpu().numpy()
    live = want >= 1e-12 * want[0]
    assert np.max(np.abs(got[live] - want[live]) / want[live]) <= 1e-10
    xo = opl.vector(0)
    assert np.max(np.abs(st["x"].cpu().numpy() - xo)) <= 1e-10 * np.max(np.abs(xo))
    ro = opl.vector(4)
    assert np.max(np.abs(st["r"].cpu().numpy() - ro)) <= 1e-9 * max(np.max(np.abs(ro)), 1e-300) + 1e-12


def test_cg_steps_misaligned_vectors_and_odd_k(ctx, oracle):
    """Vectors at different 8-byte residues mod 32 (scalar edge path), entry/col starting at an odd element."""
    from legionsolvers_b200 import kernels as K

    off, val = oracle.benchmark_stencil(3)
    m = oracle.stencil_csr((17, 13, 11), off, val)
    n, nnz, its = m.n_rows, m.nnz, 12
    buf = lambda o: torch.zeros(n + 8, dtype=torch.float64, device="cuda")[o:o + n]  # noqa: E731
    ebuf = torch.zeros(nnz + 4, dtype=torch.float64, device="cuda"); cbuf = torch.zeros(nnz + 4, dtype=torch.int64, device="cuda")
    entry, col = ebuf[1:1 + nnz], cbuf[1:1 + nnz]
    entry.copy_(dev(m.entry)); col.copy_(dev(m.col))
    st = _cg_state(oracle, m, cap=its)
    st.update(entry=entry, col=col, x=buf(1), r=buf(2), p=buf(3), q=buf(0))
    st["r"].fill_(1.0); st["p"].fill_(1.0)
    _cg_steps(ctx, st, its)
    torch.cuda.synchronize()
    opl = oracle.Planner([n], [1]); opl.fill(1, 1.0); opl.add_matrix(m)
    ocg = oracle.CGSolver(opl)
    for _ in range(its):
        ocg.step()
    want = ocg.residual_norm_squared[1:]
    assert np.max(np.abs(st["hist"].cpu().numpy() - want) / want) <= 1e-10
    xo = opl.vector(0)
    assert np.max(np.abs(st["x"].cpu().numpy() - xo)) <= 1e-10 * np.max(np.abs(xo))


@pytest.mark.parametrize("n,off", [(4, 0), (7, 1), (1000, 0), (4099, 3), (300_001, 2), (2_097_152, 0), (2_800_003, 1)])
def test_cg_tail_equals_update_then_direction(ctx, n, off):
    """lsk_cg_tail_f64 (both vector passes of the CG step in one launch) against lsk_cg_update_f64 + lsk_cg_direction_f64:
    x and r bit-identical; r.r within 1e-12 (its fold order differs); p within what that difference does to beta; history,
    rr_cur and the second launch (re-armed tickets, next launch number) included."""
    from legionsolvers_b200 import _abi

    g = torch.Generator(device="cuda").manual_seed(n)
    mk = lambda: (torch.rand(n + 8, dtype=torch.float64, device="cuda", generator=g) - 0.5)[off:off + n]  # noqa: E731
    p0, q0, x0, r0 = mk(), mk(), mk(), mk()
    if not _abi.lib().lsk_cg_direction_supported(n, r0.data_ptr(), p0.data_ptr()):
        pytest.skip("reference pair needs the streamed direction kernel")
    st = {}
    for name in ("pair", "tail"):
        st[name] = dict(rr_cur=torch.tensor([3.25], dtype=torch.float64, device="cuda"), pq=torch.tensor([1.75], dtype=torch.float64, device="cuda"),
                        rr_new=torch.zeros(1, dtype=torch.float64, device="cuda"), hist=torch.zeros(4, dtype=torch.float64, device="cuda"),
                        count=torch.zeros(1, dtype=torch.int64, device="cuda"))
        for k, src in (("p", p0), ("q", q0), ("x", x0), ("r", r0)):  # both states see the same alignment
            buf = torch.zeros(n + 8, dtype=torch.float64, device="cuda")
            buf[off:off + n].copy_(src)
            st[name][k] = buf[off:off + n]
    for launch in range(2):
        a, b = st["pair"], st["tail"]
        ctx.cg_update(a["rr_cur"], a["pq"], a["p"], a["q"], a["x"], a["r"], a["rr_new"])
        ctx.cg_direction(a["rr_cur"], a["rr_new"], a["r"], a["p"], a["hist"], a["count"])
        ctx.cg_tail(b["rr_cur"], b["pq"], b["rr_new"], b["p"], b["q"], b["x"], b["r"], b["hist"], b["count"])
        torch.cuda.synchronize()
        if launch == 0:
            assert torch.equal(a["x"], b["x"]) and torch.equal(a["r"], b["r"])
        rr_a, rr_b = a["rr_new"].item(), b["rr_new"].item()
        assert abs(rr_a - rr_b) <= 1e-12 * abs(rr_a)
        assert b["rr_cur"].item() == rr_b and int(b["count"].item()) == launch + 1 and b["hist"][launch].item() == rr_b
        scale = float(a["p"].abs().max())
        assert float((a["p"] - b["p"]).abs().max()) <= 1e-11 * scale
        for d in (a, b):
            d["pq"].fill_(2.5)


@pytest.mark.parametrize("off", [0, 1, 2, 3])
def test_blas1_tma_streamed_large(ctx, oracle, off):
    """Above 6 M elements scal / axpy / xpay / dot / dot2 / axpy_dot / bicg_p_update take the TMA-streamed, dynamically
    scheduled form: same element-wise arithmetic (bit-exact), dots within 1e-12; ragged edges at every alignment."""
    n = 6_291_456 + 23
    rng = np.random.default_rng(100 + off)
    x0, y0, w0 = rng.standard_normal(n), rng.standard_normal(n), rng.standard_normal(n)
    mk = lambda a: (lambda t: (t.copy_(dev(a)), t)[1])(torch.zeros(n + 8, dtype=torch.float64, device="cuda")[off:off + n])  # noqa: E731
    f = [0.37, -1.9, 0.61]
    alpha = oracle.get_alpha(f)
    terms = scalars(*f)
    out = torch.zeros(1, dtype=torch.float64, device="cuda")
    # scal
    xd = mk(x0); xw = x0.copy()
    oracle.scal(alpha, xw); ctx.scal(terms, xd)
    np.testing.assert_array_equal(xd.cpu().numpy(), xw)
    # axpy, xpay
    xd, yd = mk(x0), mk(y0); yw = y0.copy()
    oracle.axpy(alpha, x0, yw); ctx.axpy(terms, xd, yd)
    np.testing.assert_array_equal(yd.cpu().numpy(), yw)
    oracle.xpay(alpha, x0, yw); ctx.xpay(terms, xd, yd)
    np.testing.assert_array_equal(yd.cpu().numpy(), yw)
    # dot, dot2
    wd = mk(w0)
    ctx.dot(xd, wd, out)
    assert abs(out.item() - oracle.dot(x0, w0)) <= REL * float(np.dot(np.abs(x0), np.abs(w0)))
    o2 = torch.zeros(1, dtype=torch.float64, device="cuda")
    ctx.dot2(xd, wd, out, o2)
    assert abs(out.item() - oracle.dot(x0, w0)) <= REL * float(np.dot(np.abs(x0), np.abs(w0)))
    assert abs(o2.item() - oracle.dot(w0, w0)) <= REL * oracle.dot(w0, w0)
    # axpy_dot, separate w and aliased w = y
    yd = mk(y0); yw = y0.copy()
    oracle.axpy(alpha, x0, yw); ctx.axpy_dot(terms, xd, yd, wd, out)
    np.testing.assert_array_equal(yd.cpu().numpy(), yw)
    assert abs(out.item() - oracle.dot(yw, w0)) <= REL * float(np.dot(np.abs(yw), np.abs(w0)))
    yd = mk(y0)
    ctx.axpy_dot(terms, xd, yd, yd, out)
    assert abs(out.item() - oracle.dot(yw, yw)) <= REL * oracle.dot(yw, yw)
    # bicg_p_update: P += (-omega) V ; P = beta P + R
    rho_new, rho_old, al, om = 1.3, 0.9, 0.41, 0.77
    beta = oracle.scalar("mul", oracle.scalar("div", rho_new, rho_old), oracle.scalar("div", al, om))
    pw = y0.copy()
    oracle.axpy(-om, x0, pw); oracle.xpay(beta, w0, pw)
    pd = mk(y0)
    ctx.bicg_p_update(*scalars(rho_new, rho_old, al, om), xd, wd, pd)
    np.testing.assert_array_equal(pd.cpu().numpy(), pw)
    # and once more: the work counters were re-armed by the last CTA of every launch
    ctx.dot(xd, wd, out)
    assert abs(out.item() - oracle.dot(x0, w0)) <= REL * float(np.dot(np.abs(x0), np.abs(w0)))
