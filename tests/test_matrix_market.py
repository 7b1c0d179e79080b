"""Matrix Market reader / writer and COO -> CSR conversion (host-only code of liblsk.so: runs without a GPU).
scipy.io is the independent check: files written by scipy are read by lsk, files written by lsk are read by scipy."""
import ctypes as C

import numpy as np
import pytest
import scipy.io
import scipy.sparse as sp

from legionsolvers_b200 import solvers as S


def dense_of(rows, cols, entry, row, col):
    a = np.zeros((rows, cols))
    np.add.at(a, (row, col), entry)
    return a


@pytest.mark.parametrize("symmetry", ["general", "symmetric", "skew-symmetric"])
@pytest.mark.parametrize("field", ["real", "integer", "pattern"])
def test_reads_what_scipy_writes(tmp_path, field, symmetry):
    rng = np.random.default_rng(7)
    n = 37
    a = sp.random(n, n, density=0.12, random_state=3, format="coo")
    a.data = np.round(rng.standard_normal(a.nnz) * 100.0) if field != "real" else rng.standard_normal(a.nnz)
    if symmetry == "symmetric":
        a = (a + a.T).tocoo()
    elif symmetry == "skew-symmetric":
        a = sp.triu(a, 1)
        a = (a - a.T).tocoo()
    if field == "pattern":
        if symmetry == "skew-symmetric":
            pytest.skip("a pattern file has no signs to mirror")
        a.data[:] = 1.0
    path = tmp_path / "m.mtx"
    scipy.io.mmwrite(str(path), a, field=field, symmetry=symmetry, comment="written by scipy\nsecond comment line")
    rows, cols, entry, row, col = S.read_matrix_market(path)
    assert (rows, cols) == (n, n)
    want = scipy.io.mmread(str(path)).toarray()
    np.testing.assert_array_equal(dense_of(rows, cols, entry, row, col), want)
    assert entry.size == np.count_nonzero(want) or field == "real" or field == "integer"  # explicit zeros may be stored


def test_rectangular_values_round_trip_bit_exact(tmp_path):
    rng = np.random.default_rng(11)
    rows, cols, nnz = 23, 41, 200
    row, col = rng.integers(0, rows, nnz), rng.integers(0, cols, nnz)
    entry = rng.standard_normal(nnz) * 10.0 ** rng.integers(-200, 200, nnz)  # 17 significant digits: doubles round-trip
    path = tmp_path / "r.mtx"
    S.write_matrix_market(path, rows, cols, entry, row, col)
    r2, c2, e2, row2, col2 = S.read_matrix_market(path)
    assert (r2, c2) == (rows, cols)
    np.testing.assert_array_equal(row2, row)
    np.testing.assert_array_equal(col2, col)
    np.testing.assert_array_equal(e2, entry)          # bit-exact, duplicates and file order kept
    got = scipy.io.mmread(str(path))                   # and scipy reads the same matrix (duplicates summed)
    np.testing.assert_allclose(got.toarray(), dense_of(rows, cols, entry, row, col), rtol=1e-15, atol=0)


def test_coo_to_csr_layout():
    rng = np.random.default_rng(5)
    rows, nnz = 50, 400
    row, col = rng.integers(0, rows, nnz), rng.integers(0, 70, nnz)
    row[row == 17] = 18                                # an empty row
    entry = rng.standard_normal(nnz)
    e, c, rp = S.coo_to_csr(rows, entry, row, col)
    rp2 = np.ascontiguousarray(rp).view(np.int64).reshape(-1, 2)
    assert rp2[17, 0] > rp2[17, 1]                     # empty row: lo > hi
    k = 0
    for r in range(rows):
        idx = np.nonzero(row == r)[0]                  # input order within the row is kept (stable)
        lo, hi = rp2[r]
        assert hi - lo + 1 == idx.size and (idx.size == 0 or lo == k)
        np.testing.assert_array_equal(e[lo:hi + 1], entry[idx])
        np.testing.assert_array_equal(c[lo:hi + 1], col[idx])
        k += idx.size
    assert k == nnz


def test_empty_and_edge_files(tmp_path):
    p = tmp_path / "empty.mtx"
    p.write_text("%%MatrixMarket matrix coordinate real general\n% nothing\n\n5 7 0\n")
    rows, cols, entry, row, col = S.read_matrix_market(p)
    assert (rows, cols, entry.size) == (5, 7, 0)
    p = tmp_path / "blank.mtx"
    p.write_text("%%MatrixMarket MATRIX Coordinate Real General\n3 3 2\n\n1 1 2.5\n% a comment between entries\n3 2 -1e-3\n")
    rows, cols, entry, row, col = S.read_matrix_market(p)
    assert entry.tolist() == [2.5, -1e-3] and row.tolist() == [0, 2] and col.tolist() == [0, 1]


@pytest.mark.parametrize("text,what", [
    ("%%MatrixMarket matrix array real general\n2 2\n1\n2\n3\n4\n", "coordinate only"),
    ("%%MatrixMarket matrix coordinate complex general\n2 2 1\n1 1 1 0\n", "complex"),
    ("%%MatrixMarket matrix coordinate real hermitian\n2 2 1\n1 1 1\n", "hermitian"),
    ("%%MatrixMarket matrix coordinate real general\n2 2 2\n1 1 1.0\n", "announced"),
    ("%%MatrixMarket matrix coordinate real general\n2 2 1\n3 1 1.0\n", "out of range"),
    ("%%MatrixMarket matrix coordinate real general\n2 2 1\n1 1\n", "without a value"),
    ("%%MatrixMarket matrix coordinate real symmetric\n2 3 1\n1 1 1.0\n", "non-square"),
    ("hello\n", "banner"),
])
def test_malformed_files_are_refused(tmp_path, text, what):
    p = tmp_path / "bad.mtx"
    p.write_text(text)
    with pytest.raises(RuntimeError) as e:
        S.read_matrix_market(p)
    assert what in str(e.value)
    with pytest.raises(RuntimeError):
        S.read_matrix_market(tmp_path / "does_not_exist.mtx")
