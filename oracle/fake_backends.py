"""TEST INFRASTRUCTURE ONLY -- in-memory stand-ins for three optional third-party packages the reference imports
(`redis`, `cassandra` (cassandra-driver), `pybloomfilter`; none is installed here), so that oracle/gen_golden.py can run the
reference's OWN storage / LSHBloom code paths (datasketch/storage.py:262-1049, datasketch/lsh_bloom.py:53-380) and record
what they write.  They implement just the calls those code paths make and keep every value they are given.

    install()   # before `import datasketch`: puts the fakes into sys.modules
"""
from __future__ import annotations

import sys
import types

DB = {"hash": {}, "list": {}, "set": {}}      # the one fake Redis database: name -> dict / list / set
BLOOM_ADDS = []                                # (filename or None, value) for every BloomFilter.add
CQL = {"ddl": [], "tables": {}}                # the one fake Cassandra keyspace: executed DDL strings; table -> {(key, value): ts}


def _b(x):
    return x.encode("utf8") if isinstance(x, str) else bytes(x)


class _Pipeline:
    def __init__(self, connection_pool=None, response_callbacks=None, transaction=True, shard_hint=None):
        self.command_stack = []

    def execute_command(self, *args, **kwargs):
        self.command_stack.append(args)

    def execute(self):
        self.command_stack = []
        return []


class _Redis:
    def __init__(self, **params):
        self.connection_pool = object()
        self.response_callbacks = {}

    def hset(self, name, key, value):
        DB["hash"].setdefault(_b(name), {})[_b(key)] = _b(value)

    def hexists(self, name, key):
        return _b(key) in DB["hash"].get(_b(name), {})

    def hkeys(self, name):
        return list(DB["hash"].get(_b(name), {}).keys())

    def hvals(self, name):
        return list(DB["hash"].get(_b(name), {}).values())

    def hlen(self, name):
        return len(DB["hash"].get(_b(name), {}))

    def rpush(self, name, *values):
        DB["list"].setdefault(_b(name), []).extend(_b(v) for v in values)

    def lrange(self, name, a, b):
        return list(DB["list"].get(_b(name), []))

    def sadd(self, name, *values):
        DB["set"].setdefault(_b(name), set()).update(_b(v) for v in values)

    def smembers(self, name):
        return set(DB["set"].get(_b(name), set()))

    def pipeline(self):
        return _Pipeline()


class _BloomFilter:
    def __init__(self, capacity=None, error_rate=None, filename=None):
        self.capacity, self.error_rate, self.filename = capacity, error_rate, filename
        self.items = set()

    def add(self, x):
        BLOOM_ADDS.append((self.filename, int(x)))
        self.items.add(int(x))

    def __contains__(self, x):
        return int(x) in self.items

    def sync(self):
        pass

    @classmethod
    def open(cls, fname):
        return cls(filename=fname)


class _Prepared:
    """A prepared statement: the CQL text, its kind and its table (the reference formats the table name in)."""
    def __init__(self, query):
        import re
        self.query = " ".join(query.split())
        self.kind = self.query.split(" ", 1)[0].upper()
        m = re.search(r"(?:INTO|UPDATE|FROM)\s+(\S+)", self.query)
        self.table = m.group(1) if m else None


class _Row:
    def __init__(self, key, value, ts):
        self.key, self.value, self.ts = key, value, ts


class _Session:
    """cassandra.cluster.Session, as far as datasketch/storage.py:262-650 uses it."""
    keyspace = None

    def execute(self, query, parameters=None):
        if isinstance(query, _Prepared):
            return self._run(query, parameters)
        CQL["ddl"].append(" ".join(query.split()))
        return []

    def set_keyspace(self, keyspace):
        self.keyspace = keyspace

    def prepare(self, query):
        return _Prepared(query)

    def _run(self, st, params):
        rows = CQL["tables"].setdefault(st.table, {})
        if st.kind == "INSERT":                      # (key, value, ts)
            rows[(_b(params[0]), _b(params[1]))] = params[2]
            return []
        if st.kind == "UPDATE":                      # SET ts = ? WHERE key = ? AND value = ?
            rows[(_b(params[1]), _b(params[2]))] = params[0]
            return []
        if st.kind == "SELECT":                      # WHERE key = ?   (clustering order: value DESC)
            key = _b(params[0])
            out = [_Row(k, v, ts) for (k, v), ts in rows.items() if k == key]
            out.sort(key=lambda r: r.value, reverse=True)
            return out[:1] if "LIMIT 1" in st.query else out
        raise NotImplementedError(st.query)


class _Cluster:
    def __init__(self, seeds=None, **kwargs):
        pass

    def connect(self):
        return _Session()


class _MonotonicTimestampGenerator:
    def __init__(self):
        self.t = 0

    def __call__(self):
        self.t += 1
        return self.t


def _execute_concurrent(session, statements_and_parameters, concurrency=100, **kwargs):
    return [(True, session.execute(st, params)) for st, params in statements_and_parameters]


def install():
    cas = types.ModuleType("cassandra")
    cas_cluster = types.ModuleType("cassandra.cluster")
    cas_cluster.Cluster = _Cluster
    cas_cluster.MonotonicTimestampGenerator = _MonotonicTimestampGenerator
    cas_conc = types.ModuleType("cassandra.concurrent")
    cas_conc.execute_concurrent = _execute_concurrent
    cas.cluster, cas.concurrent = cas_cluster, cas_conc
    sys.modules["cassandra"] = cas
    sys.modules["cassandra.cluster"] = cas_cluster
    sys.modules["cassandra.concurrent"] = cas_conc
    redis = types.ModuleType("redis")
    client = types.ModuleType("redis.client")
    client.Pipeline = _Pipeline
    redis.client = client
    redis.Redis = _Redis
    redis.__version__ = "0.0.0"      # the asyncio variant (datasketch/aio/storage.py:33) then stays disabled
    sys.modules["redis"] = redis
    sys.modules["redis.client"] = client
    pbf = types.ModuleType("pybloomfilter")
    pbf.BloomFilter = _BloomFilter
    sys.modules["pybloomfilter"] = pbf


def reset():
    DB["hash"].clear(); DB["list"].clear(); DB["set"].clear()
    del BLOOM_ADDS[:]
    del CQL["ddl"][:]
    CQL["tables"].clear()
