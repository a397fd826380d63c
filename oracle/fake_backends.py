"""TEST INFRASTRUCTURE ONLY -- in-memory stand-ins for two optional third-party packages the reference imports
(`redis`, `pybloomfilter`; neither is installed here), so that oracle/gen_golden.py can run the reference's OWN
storage / LSHBloom code paths (datasketch/storage.py:819-1049, datasketch/lsh_bloom.py:53-380) and record what they
write.  They implement just the calls those code paths make and keep every value they are given.

    install()   # before `import datasketch`: puts the fakes into sys.modules
"""
from __future__ import annotations

import sys
import types

DB = {"hash": {}, "list": {}, "set": {}}      # the one fake Redis database: name -> dict / list / set
BLOOM_ADDS = []                                # (filename or None, value) for every BloomFilter.add


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


def install():
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
