"""MongoDB encoding recorder: the consumer of the encode-all-cells pass.

API and wire format of the reference's src/intercepts/db_recorder.py:27-121 (documents as in
db_schema.js:1-47): `DbRecorder(run_id, sources).setup()` registers the run in `encs` and
makes sure the source's cells / genes are imported; the interceptor it creates encodes ALL
cells once per iteration (`network.encoding_prediction(trainer.data)`, the hot path), scales
the float32 result by 255 and stores one `encits` document
    {eid, it, cids, ns, xs, ys, zs, ds}
then points `encs.defit` / `encs.showits` at it.  The assertion texts are the reference's.

What differs without changing a document:
  * `find_duplicate_ids` groups equal (x, y, z) triplets with one lexicographic sort instead
    of the reference's O(N * U) scan; groups come out in coordinate order, which is also what
    the reference's golden `[[3, 4], [2, 5]]` shows (test/db_recorder_test.py:120-149), and the
    viewer reads `ds` as a set (cellan/frontend-static/encits.js:169-176);
  * the Mongo client is injectable (`client_factory`): pymongo when it imports, otherwise the
    in-memory stand-in `fake_mongo` (neither pymongo nor mongod exist in this image);
  * all cells of an iteration share ONE BSON document (16 MB cap): an AssertionError says so
    up front instead of the server rejecting the insert.
"""
import os
from datetime import datetime

import numpy as np

from .import_barcodes import import_barcodes

MONGO_URL = 'mongodb://localhost:27017/'
MONGO_DB = 'cellcomm-update'
ENCODINGS_COLLECTION, ITERATIONS_COLLECTION = 'encs', 'encits'
CELLS_COLLECTION, GENES_COLLECTION = 'cells', 'genes'
BSON_MAX_BYTES = 16 << 20
SOURCE_KINDS = ('matrix', 'barcodes', 'genes')


def mongo_client(url=MONGO_URL):
    try:
        from pymongo import MongoClient
    except ImportError:
        from .fake_mongo import MongoClient
    return MongoClient(url)


def get_file_name(full_path):
    return full_path.rsplit('/', 1)[-1]


def check_files(sources):
    missing = [src for src in sources if not os.path.exists(src)]
    assert not missing, f'File not found: {missing[0]}'


def cell_id_from(ix):
    """Cell ids are the 1-based row positions."""
    return ix + 1


def find_duplicate_ids(np_coords):
    """Ids of cells that share exactly the same (x, y, z): one id list per coordinate that
    occurs more than once, ids ascending inside a list."""
    coords = np.asarray(np_coords)
    n = coords.shape[0]
    if n == 0:
        return []
    order = np.lexsort((coords[:, 2], coords[:, 1], coords[:, 0]))       # stable
    ranked = coords[order]
    boundary = np.flatnonzero(np.any(ranked[1:] != ranked[:-1], axis=1)) + 1
    groups = np.split(order, boundary)
    return [[cell_id_from(int(i)) for i in np.sort(g)] for g in groups if len(g) > 1]


class DbRecorder:
    def __init__(self, enc_run_id, sources, m_db=MONGO_DB, client_factory=None):
        self.enc_run_id = enc_run_id
        self.matrix_file, self.barcodes_file, self.genes_file = (sources[k] for k in SOURCE_KINDS)
        check_files((self.matrix_file, self.barcodes_file, self.genes_file))
        self.mongo_db = m_db
        self.source_id = self.barcodes = self.cell_ids = None
        self._client_factory = client_factory or mongo_client
        self._db = self._client_factory(MONGO_URL)[m_db]
        self._seen_iterations = set()

    # ------------------------------------------------------------------ setup
    def setup(self):
        self.store_encoding_run()
        self.load_barcodes()

    def store_encoding_run(self):
        runs = self._db[ENCODINGS_COLLECTION]
        assert runs.find_one({'_id': self.enc_run_id}) is None, \
            f'Encoding run id already exists: {self.enc_run_id}'
        self.source_id = get_file_name(self.barcodes_file)
        files = dict(zip(SOURCE_KINDS, (get_file_name(self.matrix_file), self.source_id,
                                        get_file_name(self.genes_file))))
        runs.insert_one({'_id': self.enc_run_id, 'date': datetime.now(), 'defit': 0,
                         'showits': [], 'srcs': files})

    def load_barcodes(self):
        assert self.source_id, 'Cannot load barcodes without encoding!'
        cells, of_source = self._db[CELLS_COLLECTION], {'sid': self.source_id}
        if not cells.count_documents(of_source):
            import_barcodes(self.source_id, self.matrix_file, self.barcodes_file, self.genes_file,
                            MONGO_URL, self.mongo_db, CELLS_COLLECTION, GENES_COLLECTION,
                            client_factory=self._client_factory)
        self.barcodes = [doc['n'] for doc in cells.find(of_source, {'_id': 0, 'n': 1})]
        self.cell_ids = list(map(cell_id_from, range(len(self.barcodes))))

    # ------------------------------------------------------------------ per iteration
    def _iteration_document(self, it, encodings):
        rows, width = encodings.shape[0], encodings.shape[1]
        assert rows == len(self.barcodes), \
            f'encodings + barcodes have different length: {rows} != {len(self.barcodes)}'
        assert width == 3, f'encodings vector length = {width}, not in x, y, z format'
        estimate = 40 * rows + sum(len(name) + 8 for name in self.barcodes)
        assert estimate < BSON_MAX_BYTES, \
            f'encits document for {rows} cells exceeds the 16 MB BSON limit'
        coords = np.multiply(encodings, 255)            # stays float32 (reference :92)
        xs, ys, zs = (coords[:, k].tolist() for k in range(3))
        return {'eid': self.enc_run_id, 'it': it, 'cids': self.cell_ids, 'ns': self.barcodes,
                'xs': xs, 'ys': ys, 'zs': zs, 'ds': find_duplicate_ids(coords)}

    def create_interceptor(self, trainer):
        assert self.barcodes, 'Cannot store iterations without barcodes!'
        shown = []

        def record(it, _losses):
            assert it not in self._seen_iterations, f'duplicate iteration {it}'
            self._seen_iterations.add(it)
            encodings = trainer.network.encoding_prediction(trainer.data)   # all cells
            self._db[ITERATIONS_COLLECTION].insert_one(self._iteration_document(it, encodings))
            shown.append(it)
            self._db[ENCODINGS_COLLECTION].update_one(
                {'_id': self.enc_run_id}, {'$set': {'defit': it, 'showits': list(shown)}})

        return record
