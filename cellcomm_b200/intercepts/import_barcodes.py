"""cells / genes collections from a 10x matrix (reference src/intercepts/import_barcodes.py:
14-81; schema db_schema.js:1-47).

    cells{sid, cid, n, g:[{e, m, v}]}   g sorted by value, descending, stable
    genes{sid, e, m, cids:[...]}        in first-seen order, cids in file order

The reference walks the nnz triplets in a Python loop and assumes barcode-sorted input; this
builds the same documents from the file's triplets with numpy grouping (stable sorts), and
tolerates unsorted barcodes by grouping on first appearance like the reference's
`current_id` change detector does for sorted files.
"""
import numpy as np


def load_file(file, converter, skip=0):
    result = []
    print(f'loading "{file}" ... ', end='', flush=True)
    with open(file) as f:
        for line in f.readlines()[skip:]:
            result.append(converter(line))
    print(f'{len(result)} lines read')
    return result


def get_line_num(arr, line_num):
    """ Barcode + genes references are 1-based
    """
    return arr[int(line_num) - 1]


def convert_matrix(source_id, barcodes, genes_src, matrix):
    print('converting matrix ... ', end='', flush=True)
    if len(matrix) == 0:
        print('DONE')
        return [], []
    trip = np.array([[int(t[0]), int(t[1]), int(float(t[2]))] for t in matrix], dtype=np.int64)
    gene_ln, cell_id, val = trip[:, 0], trip[:, 1], trip[:, 2]
    # one cell record per run of equal barcode ids (the reference starts a new record whenever
    # the barcode column changes, :22-31)
    starts = np.flatnonzero(np.r_[True, cell_id[1:] != cell_id[:-1]])
    ends = np.r_[starts[1:], len(cell_id)]
    cells = []
    for s, e in zip(starts, ends):
        order = np.argsort(-val[s:e], kind='stable') + s
        cid = int(cell_id[s])
        cells.append({
            'sid': source_id, 'cid': cid, 'n': barcodes[cid - 1],
            'g': [{'e': genes_src[gene_ln[i] - 1][0], 'm': genes_src[gene_ln[i] - 1][1],
                   'v': int(val[i])} for i in order]})
    # genes in first-seen order, cell ids in file order
    first_seen = {}
    for i, gl in enumerate(gene_ln):
        first_seen.setdefault(int(gl), []).append(int(cell_id[i]))
    genes_list = [{'sid': source_id, 'e': genes_src[gl - 1][0], 'm': genes_src[gl - 1][1],
                   'cids': cids} for gl, cids in first_seen.items()]
    print('DONE')
    return cells, genes_list


def sort_cell_genes_by_value(cells):
    for cell in cells:
        cell['g'].sort(key=lambda gene: gene['v'], reverse=True)


def import_cells(cells_genes, mongo_url, mongo_db, cells_collection, genes_collection,
                 client_factory=None):
    from .db_recorder import mongo_client
    cell_json, genes_json = cells_genes
    print('importing cells ... ', end='', flush=True)
    client = (client_factory or mongo_client)(mongo_url)
    client[mongo_db][cells_collection].insert_many(cell_json)
    print('DONE')
    print('importing genes ... ', end='', flush=True)
    client[mongo_db][genes_collection].insert_many(genes_json)
    print('DONE')
    client.close()


def import_barcodes(source_id, matrix_file, barcodes_file, genes_file,
                    mongo_url, mongo_db, cells_collection, genes_collection, client_factory=None):
    print('importing barcodes:', barcodes_file)
    matrix = load_file(matrix_file, lambda line: line.strip().split(' '), skip=3)
    barcodes = load_file(barcodes_file, lambda line: line.strip())
    genes = load_file(genes_file, lambda line: line.strip().split('\t'))

    cells_genes_data = convert_matrix(source_id, barcodes, genes, matrix)
    import_cells(cells_genes_data, mongo_url, mongo_db, cells_collection, genes_collection,
                 client_factory)
