"""MongoDB encoding recorder (reference src/intercepts/db_recorder.py:27-121; wire format
db_schema.js:1-47): once per iteration, encode ALL cells (the hot path's encode pass),
scale by 255 in float32, and write one `encits` document {eid,it,cids,ns,xs,ys,zs,ds}; keep
`encs` {_id,date,defit,showits,srcs} up to date.  Same assertion texts as the reference.

Differences that do not change the documents:
  * `find_duplicate_ids` is a sort/unique over the (x,y,z) triplets (O(N log N)) instead of
    the reference's O(N*U) Python scan; groups come out in lexicographic coordinate order,
    which reproduces the reference test's golden [[3,4],[2,5]] (test/db_recorder_test.py:
    120-149); the viewer treats `ds` as unordered (cellan/frontend-static/encits.js:169-176).
  * the Mongo client is pluggable: real pymongo when it imports, else the in-memory
    stand-in in fake_mongo.py (pymongo / mongod are not installed in this image).
  * one BSON document holds all cells (16 MB cap): a clear AssertionError is raised instead of
    a server-side failure when the document would exceed it.
"""
import os
from datetime import datetime

import numpy as np

from .import_barcodes import import_barcodes

MONGO_URL = 'mongodb://localhost:27017/'
MONGO_DB = 'cellcomm-update'
ENCODINGS_COLLECTION = 'encs'
ITERATIONS_COLLECTION = 'encits'
CELLS_COLLECTION = 'cells'
GENES_COLLECTION = 'genes'
BSON_MAX_BYTES = 16 * 1024 * 1024


def mongo_client(url=MONGO_URL):
    try:
        from pymongo import MongoClient
    except ImportError:
        from .fake_mongo import MongoClient
    return MongoClient(url)


def get_file_name(full_path):
    return full_path.split('/')[-1]


def check_files(sources):
    for src in sources:
        assert os.path.exists(src), f'File not found: {src}'


class DbRecorder:
    def __init__(self, enc_run_id, sources, m_db=MONGO_DB, client_factory=None):
        self.enc_run_id = enc_run_id
        self.matrix_file = sources['matrix']
        self.barcodes_file = sources['barcodes']
        self.genes_file = sources['genes']
        check_files([self.matrix_file, self.barcodes_file, self.genes_file])
        self.mongo_db = m_db
        self.source_id = None
        self.barcodes = None
        self.cell_ids = None
        self._client_factory = client_factory or mongo_client
        self._db = self._client_factory(MONGO_URL)[self.mongo_db]
        self._processed_its = []

    def setup(self):
        self.store_encoding_run()
        self.load_barcodes()

    def _coll(self, coll_name):
        return self._db[coll_name]

    def store_encoding_run(self):
        assert self._coll(ENCODINGS_COLLECTION).find_one({'_id': self.enc_run_id}) is None, \
            f'Encoding run id already exists: {self.enc_run_id}'

        self.source_id = get_file_name(self.barcodes_file)
        self._coll(ENCODINGS_COLLECTION).insert_one({
            '_id': self.enc_run_id,
            'date': datetime.now(),
            'defit': 0,
            'showits': [],
            'srcs': {
                'matrix': get_file_name(self.matrix_file),
                'barcodes': self.source_id,
                'genes': get_file_name(self.genes_file)
            }
        })

    def load_barcodes(self):
        assert self.source_id, 'Cannot load barcodes without encoding!'
        cells = self._coll(CELLS_COLLECTION)
        query = {'sid': self.source_id}
        if cells.count_documents(query) == 0:
            import_barcodes(
                self.source_id, self.matrix_file, self.barcodes_file, self.genes_file,
                MONGO_URL, self.mongo_db, CELLS_COLLECTION, GENES_COLLECTION,
                client_factory=self._client_factory
            )
        self.barcodes = [cell['n'] for cell in cells.find(query, {'_id': 0, 'n': 1})]
        self.cell_ids = [cell_id_from(i) for i in range(len(self.barcodes))]

    def create_interceptor(self, trainer):
        assert self.barcodes, 'Cannot store iterations without barcodes!'
        show_iterations = []

        def intercept(it, _):
            assert it not in self._processed_its, f'duplicate iteration {it}'
            self._processed_its.append(it)
            encodings = trainer.network.encoding_prediction(trainer.data)
            enc_shape = encodings.shape
            assert enc_shape[0] == len(self.barcodes), \
                f'encodings + barcodes have different length: {enc_shape[0]} != {len(self.barcodes)}'
            assert enc_shape[1] == 3, \
                f'encodings vector length = {enc_shape[1]}, not in x, y, z format'

            coords = np.multiply(encodings, 255)
            est = 40 * len(self.barcodes) + sum(len(b) + 8 for b in self.barcodes)
            assert est < BSON_MAX_BYTES, \
                f'encits document for {len(self.barcodes)} cells exceeds the 16 MB BSON limit'
            self._coll(ITERATIONS_COLLECTION).insert_one({
                'eid': self.enc_run_id,
                'it': it,
                'cids': self.cell_ids,
                'ns': self.barcodes,
                'xs': coords[:, 0].tolist(),
                'ys': coords[:, 1].tolist(),
                'zs': coords[:, 2].tolist(),
                'ds': find_duplicate_ids(coords)
            })

            show_iterations.append(it)
            self._coll(ENCODINGS_COLLECTION).update_one(
                {'_id': self.enc_run_id},
                {'$set': {'defit': it, 'showits': list(show_iterations)}}
            )

        return intercept


def find_duplicate_ids(np_coords):
    """Cell ids (1-based) that share exactly the same (x, y, z): [[ids...], ...], one list per
    coordinate occurring more than once."""
    coords = np.asarray(np_coords)
    if coords.shape[0] == 0:
        return []
    order = np.lexsort((coords[:, 2], coords[:, 1], coords[:, 0]))    # stable
    s = coords[order]
    new_group = np.r_[True, np.any(s[1:] != s[:-1], axis=1)]
    starts = np.flatnonzero(new_group)
    ends = np.r_[starts[1:], len(order)]
    return [[cell_id_from(int(i)) for i in np.sort(order[a:b])]
            for a, b in zip(starts, ends) if b - a > 1]


def cell_id_from(ix):
    return ix + 1
