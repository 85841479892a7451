"""The C-ABI .mtx loader (cc_mtx_load_csr / cc_coo_to_csr, host code, no GPU) against the
oracle's pandas pivot and the reference's golden fixture (test/cell_type_training_test.py:
14-21,32-41).  Integer / index work: bit-exact."""
import os

import numpy as np
import pytest

from oracle import loader_oracle as LO
from cellcomm_b200 import cell_type_training as ctt

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
FIXTURE = os.path.join(GOLDEN, "example_matrix.mtx")

TEST_MATRIX_CONTENT = [   # reference test/cell_type_training_test.py:14-21
    [0, 1, 6, 1, 11],
    [4, 1, 0, 1, 6],
    [1, 1, 1, 1, 6],
    [1, 0, 14, 0, 1],
    [0, 0, 0, 2, 0],
]


def test_fixture_golden_matrix():
    m = ctt.load_matrix(FIXTURE)
    assert m.shape == (5, 5)
    assert m.values.tolist() == TEST_MATRIX_CONTENT
    assert m.values.dtype == np.float64
    assert m.index.tolist() == [1, 2, 3, 4, 5] and m.columns.tolist() == [1, 2, 3, 4, 5]
    # the dims line of the fixture (27998 2405 3399591) is wrong on purpose and ignored
    assert m.nnz == 17


def test_fixture_matches_pandas_and_numpy_oracles():
    m = ctt.load_matrix(FIXTURE)
    df = LO.load_matrix_pandas(FIXTURE)
    assert np.array_equal(m.values, df.values)
    assert m.index.tolist() == df.index.tolist() and m.columns.tolist() == df.columns.tolist()
    dense, rows, cols = LO.load_matrix_numpy(FIXTURE)
    assert np.array_equal(m.values, dense)


def test_sampler_golden_rows():
    m = ctt.load_matrix(FIXTURE)
    tr = ctt.CellTraining.__new__(ctt.CellTraining)
    tr.data, tr.batch_size = m, 3
    s = tr.sample_cell_data(0)
    assert s.shape == (3, 5)
    assert s.values.tolist() == [TEST_MATRIX_CONTENT[2], TEST_MATRIX_CONTENT[0],
                                 TEST_MATRIX_CONTENT[1]]
    # unseeded sampling draws from numpy's global state, like DataFrame.sample
    np.random.seed(5)
    a = m.sample(4).positions.tolist()
    np.random.seed(5)
    assert a == np.random.permutation(5)[:4].tolist()
    df = LO.load_matrix_pandas(FIXTURE)
    for seed in (0, 1, 7, 123):
        assert np.array_equal(m.sample(3, random_state=seed).values,
                              df.sample(3, random_state=seed).values)


def _write_mtx(path, genes, barcodes, vals, header_dims="1 1 1"):
    with open(path, "w") as f:
        f.write("%%MatrixMarket matrix coordinate integer general\n%\n" + header_dims + "\n")
        for g, b, v in zip(genes, barcodes, vals):
            f.write(f"{g} {b} {v}\n")


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_random_coo_with_duplicates_zeros_and_gaps(tmp_path, seed):
    """duplicates are averaged, explicit zeros keep rows/cols alive, absent ids are dropped,
    input order is irrelevant"""
    rng = np.random.default_rng(seed)
    n = 4000
    genes = rng.choice(np.arange(1, 400, 3), n)          # gaps in gene ids
    barcodes = rng.choice(np.r_[np.arange(5, 60), 1000], n)
    vals = rng.integers(0, 20, n)                        # includes explicit zeros + duplicates
    path = tmp_path / "m.mtx"
    _write_mtx(path, genes, barcodes, vals)
    m = ctt.load_matrix(str(path))
    df = LO.load_matrix_pandas(str(path))
    assert m.shape == df.shape
    assert m.index.tolist() == df.index.tolist()
    assert m.columns.tolist() == df.columns.tolist()
    assert np.array_equal(m.values, df.values)           # float64 means, bit-exact
    # float32 CSR values = float32(float64 mean)
    assert np.array_equal(m.values32, m.values64.astype(np.float32))
    # same through the in-memory COO entry point
    m2 = ctt.CellMatrix.from_coo(genes, barcodes, vals.astype(np.float64))
    assert np.array_equal(m2.values, df.values)


def test_ragged_and_edge_inputs(tmp_path):
    # blank lines are skipped (pandas skip_blank_lines), tabs / CRLF tolerated
    p = tmp_path / "a.mtx"
    p.write_text("%%MatrixMarket\n%\n9 9 9\n3 2 5\n\n1\t7\t2\r\n3 2 7\n")
    m = ctt.load_matrix(str(p))
    assert m.shape == (2, 2) and m.values.tolist() == [[0.0, 6.0], [2.0, 0.0]]
    # header only -> empty matrix
    q = tmp_path / "b.mtx"
    q.write_text("%%MatrixMarket\n%\n0 0 0\n")
    assert ctt.load_matrix(str(q)).shape == (0, 0)
    # malformed line -> loud error with the offending text
    r = tmp_path / "c.mtx"
    r.write_text("%%MatrixMarket\n%\n1 1 1\n1 2\n")
    with pytest.raises(Exception, match="malformed line"):
        ctt.load_matrix(str(r))
    with pytest.raises(Exception, match="cannot open"):
        ctt.load_matrix(str(tmp_path / "missing.mtx"))


def test_large_multithreaded_parse_matches_numpy(tmp_path):
    """> 1 MiB body takes the chunked multi-thread parse path"""
    rng = np.random.default_rng(9)
    n = 200_000
    genes = rng.integers(1, 3000, n)
    barcodes = np.sort(rng.integers(1, 800, n))
    vals = rng.integers(1, 300, n)
    path = tmp_path / "big.mtx"
    _write_mtx(path, genes, barcodes, vals)
    assert os.path.getsize(path) > (1 << 20)
    m = ctt.load_matrix(str(path))
    dense, rows, cols = LO.load_matrix_numpy(str(path))
    assert np.array_equal(m.values, dense)
    assert m.index.tolist() == rows.tolist() and m.columns.tolist() == cols.tolist()


def test_convert_roundtrip(tmp_path):
    """`python3 src convert` writes a dense CSV that load_cells reads back (src/__main__.py:76-81)"""
    m = ctt.load_matrix(FIXTURE)
    out = tmp_path / "cells.csv"
    m.to_csv(str(out))
    back = ctt.load_cells(str(out))
    assert back.shape == m.shape and np.array_equal(back.values, m.values)
    assert back.index.tolist() == m.index.tolist()
