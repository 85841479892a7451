"""ORACLE — test infrastructure, NOT product code.

CPU restatement of the reference's data path:
  * load_matrix        src/cell_type_training.py:9-17   (pandas read_csv + pivot_table)
  * sample_cell_data   src/cell_type_training.py:37-38  (DataFrame.sample)
  * find_duplicate_ids src/intercepts/db_recorder.py:113-117
  * encits coords      src/intercepts/db_recorder.py:92-101

`load_matrix_pandas` is the reference's own body with the single keyword pandas >= 2.2 removed
(`delim_whitespace=True` -> `sep=r'\\s+'`, same tokenisation).  `load_matrix_numpy` is an
independent numpy restatement of the semantics; the two are checked
against each other and against the reference's golden matrix
(test/cell_type_training_test.py:14-21) in tests/test_oracle_goldens.py.
PARITY STATUS: pinned (loader golden, sampler golden [2,0,1], encits golden).
"""
import numpy as np
import pandas as pd


def load_matrix_pandas(matrix_file):
    df = pd.read_csv(matrix_file, header=None, skiprows=3, sep=r"\s+",
                     names=["gene", "barcode", "p"])
    return df.pivot_table(index="barcode", columns="gene", values="p", fill_value=0)


def load_matrix_numpy(matrix_file):
    """-> (dense float64 [rows, cols], barcode ids ascending, gene ids ascending)."""
    genes, barcodes, vals = [], [], []
    with open(matrix_file) as f:
        for i, line in enumerate(f):
            if i < 3:
                continue
            parts = line.split()
            if not parts:
                continue
            genes.append(int(parts[0]))
            barcodes.append(int(parts[1]))
            vals.append(float(parts[2]))
    genes, barcodes, vals = np.array(genes), np.array(barcodes), np.array(vals, dtype=np.float64)
    row_ids, ri = np.unique(barcodes, return_inverse=True)
    col_ids, ci = np.unique(genes, return_inverse=True)
    sums = np.zeros((len(row_ids), len(col_ids)))
    cnts = np.zeros((len(row_ids), len(col_ids)))
    np.add.at(sums, (ri, ci), vals)
    np.add.at(cnts, (ri, ci), 1)
    dense = np.where(cnts > 0, sums / np.maximum(cnts, 1), 0.0)
    return dense, row_ids, col_ids


def sample_indices(n_rows, batch_size, random_state=None):
    """Row positions DataFrame.sample(batch_size, random_state=...) picks: without replacement,
    `RandomState(seed)` for an int seed, the GLOBAL numpy state for None."""
    rs = np.random.RandomState(random_state) if random_state is not None else np.random
    return rs.permutation(n_rows)[:batch_size]


def find_duplicate_ids(np_coords):
    """src/intercepts/db_recorder.py:113-117 verbatim semantics (O(N*U))."""
    coords = [(c[0], c[1], c[2]) for c in np_coords]
    unique_coords = set(coords)
    all_indices = [[i + 1 for i, x in enumerate(coords) if x == uc] for uc in unique_coords]
    return [ixs for ixs in all_indices if len(ixs) > 1]


def convert_matrix_walk(source_id, barcodes, genes_src, matrix):
    """src/intercepts/import_barcodes.py:14-50 restated: one pass over the `(gene line, barcode
    line, value)` string triples in file order.  A new cell record starts whenever the barcode
    column differs from the previous line's (:22-31); every entry appends `{e, m, v=int(p)}` to
    the current cell (:33-39) and the cell id to the gene record keyed by ensembl id, created
    on first sight (:40-41); finally each cell's gene list is sorted by value, descending,
    with Python's stable sort (:43, :48-50).  Gene records come out in first-seen order (:44)."""
    cells, gene_records, previous = [], {}, None
    for gene_line, barcode_line, p_val in matrix:
        if barcode_line != previous:
            previous = barcode_line
            cells.append({"sid": source_id, "cid": int(barcode_line),
                          "n": barcodes[int(barcode_line) - 1], "g": []})
        ensembl, mgi = genes_src[int(gene_line) - 1][0], genes_src[int(gene_line) - 1][1]
        cells[-1]["g"].append({"e": ensembl, "m": mgi, "v": int(p_val)})
        if ensembl not in gene_records:
            gene_records[ensembl] = {"sid": source_id, "e": ensembl, "m": mgi, "cids": []}
        gene_records[ensembl]["cids"].append(int(barcode_line))
    for cell in cells:
        cell["g"] = sorted(cell["g"], key=lambda g: -g["v"])
    return cells, list(gene_records.values())
