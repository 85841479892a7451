"""cells / genes collections of a 10x source, built from the loader's parsed triplets.

Documents (db_schema.js:1-47), as the reference's src/intercepts/import_barcodes.py:14-81
produces them:

    cells{sid, cid, n, g:[{e, m, v}]}   one per run of equal barcode ids in FILE order;
                                        g sorted by value, descending, stable
    genes{sid, e, m, cids:[...]}        one per ensembl id, in first-seen order; cids in file order

The reference reads the matrix text a second time and walks the nnz triplets one by one in
Python.  Here the matrix is parsed once by the threaded C++ loader (`cc_mtx_load_coo`, the same
parse `load_matrix` uses) and the grouping is three stable numpy sorts over the triplet arrays;
Python only assembles the final documents.  `import_barcodes(...)` keeps the reference's
signature (plus an injectable client factory, as in db_recorder).
"""
import ctypes as C

import numpy as np

from .. import _lib


def read_lines(path, split=None):
    """Stripped lines of a small text file (barcodes.tsv / genes.tsv), optionally split."""
    print(f'loading "{path}" ... ', end='', flush=True)
    with open(path) as f:
        rows = [ln.strip() for ln in f]
    if split is not None:
        rows = [ln.split(split) for ln in rows]
    print(f'{len(rows)} lines read')
    return rows


def load_triplets(matrix_file):
    """(gene, barcode, value) int64 / int64 / float64 arrays in file order."""
    lib = _lib.load()
    h = C.c_void_p()
    _lib.check(lib.cc_mtx_load_coo(str(matrix_file).encode(), C.byref(h)))
    try:
        n = lib.cc_coo_nnz(h)

        def view(ptr, ctype):
            if n == 0:
                return np.zeros(0, dtype=np.dtype(ctype))
            return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ctype)), shape=(n,)).copy()

        return (view(lib.cc_coo_gene(h), C.c_int64), view(lib.cc_coo_barcode(h), C.c_int64),
                view(lib.cc_coo_value(h), C.c_double))
    finally:
        lib.cc_coo_destroy(h)


def convert_matrix(source_id, barcodes, genes_src, matrix):
    """-> (cell documents, gene documents).  `matrix` is either the (gene, barcode, value)
    arrays of `load_triplets` or any sequence of `(gene, barcode, value)` rows / string
    triples (what the reference's callers pass)."""
    print('converting matrix ... ', end='', flush=True)
    if isinstance(matrix, tuple) and len(matrix) == 3 and isinstance(matrix[0], np.ndarray):
        gene_ln, cell_id, val = (np.asarray(a) for a in matrix)
    else:
        rows = list(matrix)
        gene_ln = np.array([int(r[0]) for r in rows], dtype=np.int64)
        cell_id = np.array([int(r[1]) for r in rows], dtype=np.int64)
        val = np.array([float(r[2]) for r in rows], dtype=np.float64)
    nnz = len(val)
    if nnz == 0:
        print('DONE')
        return [], []
    ival = val.astype(np.int64)                 # 'v': int(p_val)
    ensembl = [g[0] for g in genes_src]
    symbol = [g[1] for g in genes_src]
    g0 = gene_ln - 1                            # gene / barcode references are 1-based

    # ---- cells: a new record whenever the barcode column changes; inside a record the genes
    # are ordered by value, descending, ties in file order (list.sort is stable)
    run = np.cumsum(np.r_[0, cell_id[1:] != cell_id[:-1]])
    order = np.lexsort((np.arange(nnz), -ival, run))
    bounds = np.flatnonzero(np.r_[True, run[order][1:] != run[order][:-1], True])
    og, ov = g0[order].tolist(), ival[order].tolist()
    cells = []
    for s, e in zip(bounds[:-1].tolist(), bounds[1:].tolist()):
        cid = int(cell_id[order[s]])
        cells.append({'sid': source_id, 'cid': cid, 'n': barcodes[cid - 1],
                      'g': [{'e': ensembl[g], 'm': symbol[g], 'v': v}
                            for g, v in zip(og[s:e], ov[s:e])]})

    # ---- genes: keyed by ensembl id (two gene lines sharing one id share a record), listed in
    # first-seen order, each with its cell ids in file order
    first_line = {}
    canon = np.array([first_line.setdefault(e, i) for i, e in enumerate(ensembl)], dtype=np.int64)
    key = canon[g0]
    by_gene = np.argsort(key, kind='stable')
    gb = np.flatnonzero(np.r_[True, key[by_gene][1:] != key[by_gene][:-1], True])
    starts = gb[:-1]
    seen_order = np.argsort(by_gene[starts], kind='stable')     # first occurrence in the file
    cids_sorted = cell_id[by_gene].tolist()
    genes = []
    for j in seen_order.tolist():
        g = int(g0[by_gene[starts[j]]])      # the gene line of the id's first entry names it
        genes.append({'sid': source_id, 'e': ensembl[g], 'm': symbol[g],
                      'cids': cids_sorted[gb[j]:gb[j + 1]]})
    print('DONE')
    return cells, genes


def store_documents(cells_genes, mongo_url, mongo_db, cells_collection, genes_collection,
                    client_factory=None):
    from .db_recorder import mongo_client
    cells, genes = cells_genes
    client = (client_factory or mongo_client)(mongo_url)
    try:
        for what, name, docs in (('cells', cells_collection, cells), ('genes', genes_collection, genes)):
            print(f'importing {what} ... ', end='', flush=True)
            if docs:
                client[mongo_db][name].insert_many(docs)
            print('DONE')
    finally:
        client.close()


def import_barcodes(source_id, matrix_file, barcodes_file, genes_file,
                    mongo_url, mongo_db, cells_collection, genes_collection, client_factory=None):
    print('importing barcodes:', barcodes_file)
    triplets = load_triplets(matrix_file)
    print(f'loaded "{matrix_file}": {len(triplets[2])} entries')
    barcodes = read_lines(barcodes_file)
    genes = read_lines(genes_file, split='\t')
    store_documents(convert_matrix(source_id, barcodes, genes, triplets), mongo_url, mongo_db,
                    cells_collection, genes_collection, client_factory)
