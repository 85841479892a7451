"""In-memory stand-in for the slice of pymongo the recorder uses (MongoClient()[db][coll] with
insert_one / insert_many / find_one / find / count_documents / update_one, and
drop_database).  pymongo and mongod are not installed in this image; when pymongo imports,
DbRecorder uses the real client instead.  Documents are deep-copied on the way in and out,
like a real server round trip."""
import copy
import itertools

_ids = itertools.count(1)


def _matches(doc, query):
    return all(doc.get(k) == v for k, v in (query or {}).items())


def _project(doc, projection):
    if not projection:
        return copy.deepcopy(doc)
    include = [k for k, v in projection.items() if v and k != '_id']
    if include:
        out = {k: copy.deepcopy(doc[k]) for k in include if k in doc}
        if projection.get('_id', 1):
            out['_id'] = doc.get('_id')
        return out
    return {k: copy.deepcopy(v) for k, v in doc.items() if projection.get(k, 1)}


class Collection:
    def __init__(self):
        self.docs = []

    def insert_one(self, doc):
        d = copy.deepcopy(doc)
        d.setdefault('_id', next(_ids))
        if any(x['_id'] == d['_id'] for x in self.docs):
            raise ValueError(f"duplicate key error: _id {d['_id']!r}")
        doc.setdefault('_id', d['_id'])
        self.docs.append(d)

    def insert_many(self, docs):
        for d in docs:
            self.insert_one(d)

    def find_one(self, query=None, projection=None):
        for d in self.docs:
            if _matches(d, query):
                return _project(d, projection)
        return None

    def find(self, query=None, projection=None):
        return [_project(d, projection) for d in self.docs if _matches(d, query)]

    def count_documents(self, query=None):
        return sum(1 for d in self.docs if _matches(d, query))

    def update_one(self, query, update):
        for d in self.docs:
            if _matches(d, query):
                for k, v in update.get('$set', {}).items():
                    d[k] = copy.deepcopy(v)
                return


class Database(dict):
    def __missing__(self, name):
        self[name] = Collection()
        return self[name]


class MongoClient:
    _servers = {}     # url -> {db name -> Database}: one "server" per URL within the process

    def __init__(self, url='mongodb://localhost:27017/'):
        self._dbs = MongoClient._servers.setdefault(url, {})

    def __getitem__(self, name):
        return self._dbs.setdefault(name, Database())

    def drop_database(self, name):
        self._dbs.pop(name, None)

    def close(self):
        pass
