"""cells / genes documents from the loader's triplets (cellcomm_b200/intercepts/
import_barcodes.py) against the restated per-line walk of the reference
(oracle/loader_oracle.convert_matrix_walk, src/intercepts/import_barcodes.py:14-50) and the
reference's own goldens (test/db_recorder_test.py:60-107).  Document work: EXACT equality."""
import os

import numpy as np

from cellcomm_b200.intercepts import import_barcodes as ib
from oracle import loader_oracle as LO

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _fixture():
    barcodes = [ln.strip() for ln in open(os.path.join(GOLDEN, "example_barcodes.tsv"))]
    genes = [ln.strip().split("\t") for ln in open(os.path.join(GOLDEN, "example_genes.tsv"))]
    lines = open(os.path.join(GOLDEN, "example_matrix.mtx")).readlines()[3:]
    return barcodes, genes, [ln.strip().split(" ") for ln in lines]


def test_triplets_keep_file_order():
    _, _, rows = _fixture()
    g, b, v = ib.load_triplets(os.path.join(GOLDEN, "example_matrix.mtx"))
    assert g.tolist() == [int(r[0]) for r in rows]
    assert b.tolist() == [int(r[1]) for r in rows]
    assert v.tolist() == [float(r[2]) for r in rows]


def test_fixture_documents_equal_the_reference_walk():
    barcodes, genes, rows = _fixture()
    trip = ib.load_triplets(os.path.join(GOLDEN, "example_matrix.mtx"))
    got = ib.convert_matrix("src-1", barcodes, genes, trip)
    ref = LO.convert_matrix_walk("src-1", barcodes, genes, rows)
    assert got[0] == ref[0] and got[1] == ref[1]
    assert ib.convert_matrix("src-1", barcodes, genes, rows) == (ref[0], ref[1])
    # reference goldens, test/db_recorder_test.py:60-107
    assert got[0][0]["g"] == [{"e": "ENSMUSG00000025902", "m": "Sox17", "v": 11},
                              {"e": "ENSMUSG00000102343", "m": "Gm37381", "v": 6},
                              {"e": "ENSMUSG00000089699", "m": "Gm1992", "v": 1},
                              {"e": "ENSMUSG00000109048", "m": "Rp1", "v": 1}]
    assert got[1][0] == {"sid": "src-1", "e": "ENSMUSG00000089699", "m": "Gm1992", "cids": [1, 2, 3]}


def test_random_matrices_including_unsorted_barcodes_and_shared_ensembl_ids(tmp_path):
    rng = np.random.default_rng(3)
    for trial in range(4):
        n_cells, n_genes, nnz = 40, 25, 600
        barcodes = [f"BC{i:03d}-1" for i in range(n_cells)]
        genes = [[f"ENS{i:05d}", f"sym{i}"] for i in range(n_genes)]
        genes[7][0] = genes[3][0]                     # two gene lines share an ensembl id
        cell = rng.integers(1, n_cells + 1, nnz)
        if trial % 2 == 0:
            cell = np.sort(cell)                      # barcode-sorted like real 10x files
        gene = rng.integers(1, n_genes + 1, nnz)
        val = rng.integers(1, 6, nnz)                 # many ties: the stable order matters
        path = tmp_path / f"m{trial}.mtx"
        with open(path, "w") as f:
            f.write("%%MatrixMarket matrix coordinate integer general\n%\n1 1 1\n")
            for g, c, v in zip(gene, cell, val):
                f.write(f"{g} {c} {v}\n")
        rows = [ln.strip().split(" ") for ln in open(path).readlines()[3:]]
        got = ib.convert_matrix("s", barcodes, genes, ib.load_triplets(str(path)))
        ref = LO.convert_matrix_walk("s", barcodes, genes, rows)
        assert got[0] == ref[0]
        assert got[1] == ref[1]


def test_empty_matrix(tmp_path):
    path = tmp_path / "e.mtx"
    path.write_text("%%MatrixMarket\n%\n0 0 0\n")
    assert ib.convert_matrix("s", [], [], ib.load_triplets(str(path))) == ([], [])
