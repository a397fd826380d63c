"""Wire formats of the reference's storage layer, so that an index built on the GPU can be handed to an existing
datasketch deployment unchanged.

``MinHashLSH`` over Redis (datasketch/lsh.py:191-200, datasketch/storage.py:902-1049) lays an index out as

    keys container      name = basename + b"_keys"                      (RedisListStorage)
        HSET  name            <key>  ->  name + <key>                    (storage.py:1002-1004)
        RPUSH name + <key>    H_0, ..., H_{b-1}                          (the document's b band keys, lsh.py:345)
    band i container    name_i = basename + b"_bucket_" + pack(">H", i)  (RedisSetStorage, lsh.py:192-197)
        HSET  name_i          <H>    ->  name_i + <H>                    (storage.py:1042-1045)
        SADD  name_i + <H>    <key>                                      (lsh.py:346-347)

with ``<key>`` = ``pickle.dumps(user key)`` when ``prepickle`` (the default for Redis, lsh.py:182) and ``<H>`` = the
band's r hash values as big-endian uint64 bytes (``_H``, lsh.py:537-538), optionally passed through ``hashfunc``
(:540-543).  The band keys of a whole signature matrix come from one ``dsk_band_keys`` launch; everything else is
dictionary glue on the host.  ``RedisLayout`` holds the three Redis data types as dicts, can replay itself as
commands into any client with the redis-py pipeline interface, and compares equal to the state the reference itself
produces on an in-memory stand-in (fixture tests/golden/storage.npz).

``MinHashLSH`` over Cassandra (datasketch/storage.py:262-819) uses one table per container,

    table   "lsh_" + basename + "_keys"                              rows (key = <key>, value = H_i, ts)   INSERT, storage.py:375, :505-516
    table   "lsh_" + basename + "_bucket_" + hexlify(pack(">H", i))  rows (key = <H>,   value = <key>, ts) UPDATE (upsert), :369-373, :518-534

``(key blob, value blob, ts bigint, PRIMARY KEY (key, value)) WITH CLUSTERING ORDER BY (value DESC)`` (:324-331), ``ts`` from
a per-table monotonic generator (:384) -- so a document's band keys are recovered in band order by sorting on ``ts``, and a
repeated (key, value) pair keeps one row.  ``prepickle`` defaults to False for Cassandra (lsh.py:182): keys must be bytes
unless ``prepickle=True``.  ``CassandraLayout`` holds the rows per table and replays itself into a cassandra-driver session;
it compares equal (rows and their ``ts`` order) to what the reference's own storage code wrote on an in-memory stand-in
(fixture tests/golden/storage_cassandra.npz).
"""
from __future__ import annotations

import binascii
import pickle
import struct
from typing import Callable, Dict, Hashable, Iterator, List, Optional, Sequence, Set, Tuple

import numpy as np


def bucket_container_name(basename: bytes, band: int) -> bytes:
    """lsh.py:192-197."""
    return b"".join([basename, b"_bucket_", struct.pack(">H", band)])


def keys_container_name(basename: bytes) -> bytes:
    """lsh.py:200."""
    return b"".join([basename, b"_keys"])


class RedisLayout:
    """The Redis state of one MinHashLSH index: ``hash`` (name -> {field: value}), ``list`` (name -> [values]),
    ``set`` (name -> {members})."""

    def __init__(self, basename: bytes, b: int):
        self.basename, self.b = bytes(basename), int(b)
        self.hash: Dict[bytes, Dict[bytes, bytes]] = {}
        self.list: Dict[bytes, List[bytes]] = {}
        self.set: Dict[bytes, Set[bytes]] = {}

    def add(self, key: bytes, band_keys: Sequence[bytes]) -> None:
        """What ``MinHashLSH._insert`` does to the database for one document (lsh.py:344-347)."""
        if len(band_keys) != self.b:
            raise ValueError("expected %d band keys, got %d" % (self.b, len(band_keys)))
        kname = keys_container_name(self.basename)
        self.hash.setdefault(kname, {})[key] = kname + key
        self.list.setdefault(kname + key, []).extend(band_keys)
        for i, h in enumerate(band_keys):
            name = bucket_container_name(self.basename, i)
            self.hash.setdefault(name, {})[h] = name + h
            self.set.setdefault(name + h, set()).add(key)

    def commands(self) -> Iterator[Tuple]:
        """The layout as Redis commands ``(b"HSET", name, field, value)``, ``(b"RPUSH", name, *values)``,
        ``(b"SADD", name, *members)`` (deterministic order)."""
        for name in sorted(self.hash):
            for field in sorted(self.hash[name]):
                yield (b"HSET", name, field, self.hash[name][field])
        for name in sorted(self.list):
            yield (b"RPUSH", name) + tuple(self.list[name])
        for name in sorted(self.set):
            yield (b"SADD", name) + tuple(sorted(self.set[name]))

    def write(self, client, batch: int = 50000) -> int:
        """Replay into a redis-py style client (``pipeline()`` with ``hset / rpush / sadd / execute``)."""
        pipe, n = client.pipeline(), 0
        for cmd in self.commands():
            if cmd[0] == b"HSET":
                pipe.hset(cmd[1], cmd[2], cmd[3])
            elif cmd[0] == b"RPUSH":
                pipe.rpush(cmd[1], *cmd[2:])
            else:
                pipe.sadd(cmd[1], *cmd[2:])
            n += 1
            if n % batch == 0:
                pipe.execute()
        pipe.execute()
        return n

    def canonical(self):
        """Order-free, comparable form (what the golden fixture stores)."""
        return {"hash": {k: dict(v) for k, v in self.hash.items()}, "list": {k: list(v) for k, v in self.list.items()},
                "set": {k: sorted(v) for k, v in self.set.items()}}

    def __eq__(self, other):
        return isinstance(other, RedisLayout) and self.canonical() == other.canonical()


def redis_layout(keys: Sequence[Hashable], signatures, b: int, r: int, basename: bytes, prepickle: bool = True,
                 hashfunc: Optional[Callable[[bytes], bytes]] = None) -> RedisLayout:
    """Redis layout of ``MinHashLSH(params=(b, r), storage_config={"type": "redis", "basename": basename})`` after
    ``insert(keys[i], signature i)`` for every row of ``signatures`` ([N, K] uint32/uint64 numpy array or CUDA int32
    tensor).  All N*b band keys come from one ``dsk_band_keys`` launch."""
    from . import codec
    from .lsh import _signature_matrix
    sig = _signature_matrix(signatures) if isinstance(signatures, np.ndarray) or not hasattr(signatures, "is_cuda") else signatures
    n = int(sig.shape[0])
    if len(keys) != n:
        raise ValueError("keys and signatures differ in length")
    layout = RedisLayout(basename, b)
    if n == 0:
        return layout
    width = 8 * r
    raw = codec.band_keys(sig, b, r).cpu().numpy().reshape(n, b * width).tobytes()
    for i, key in enumerate(keys):
        if prepickle:
            key = pickle.dumps(key)
        elif not isinstance(key, bytes):
            raise TypeError(f"prepickle=False requires bytes keys for non-dict storage, got {type(key).__name__}. "
                            "Either pass bytes keys or use prepickle=True for automatic serialization.")
        base = i * b * width
        hs = [raw[base + j * width: base + (j + 1) * width] for j in range(b)]
        if hashfunc is not None:
            hs = [hashfunc(h) for h in hs]
        layout.add(key, hs)
    return layout


# ---- Cassandra ---------------------------------------------------------------------------------------------------------
CASSANDRA_CREATE_TABLE = ("CREATE TABLE IF NOT EXISTS {} ( key blob, value blob, ts bigint, PRIMARY KEY (key, value) ) "
                          "WITH CLUSTERING ORDER BY (value DESC)")                               # storage.py:324-331
CASSANDRA_INSERT = "INSERT INTO {} (key, value, ts) VALUES (?, ?, ?)"                              # storage.py:375


def cassandra_table_name(container_name: bytes) -> str:
    """storage.py:398-402: bucket containers get their 2-byte band number hex-encoded, every table the ``lsh_`` prefix."""
    name = container_name
    if b"bucket" in name:
        basename, _, ret = name.split(b"_", 2)
        name = basename + b"_bucket_" + binascii.hexlify(ret)
    return "lsh_" + name.decode("ascii")


class CassandraLayout:
    """The Cassandra tables of one MinHashLSH index: ``tables[name]`` maps ``(key, value)`` to ``ts`` (1, 2, ... in the
    order the reference's client would have written the rows of that table)."""

    def __init__(self, basename: bytes, b: int):
        self.basename, self.b = bytes(basename), int(b)
        self.keys_table = cassandra_table_name(keys_container_name(self.basename))
        self.bucket_tables = [cassandra_table_name(bucket_container_name(self.basename, i)) for i in range(self.b)]
        self.tables: Dict[str, Dict[Tuple[bytes, bytes], int]] = {t: {} for t in [self.keys_table] + self.bucket_tables}
        self._ts: Dict[str, int] = {t: 0 for t in self.tables}

    def _put(self, table: str, key: bytes, value: bytes) -> None:
        self._ts[table] += 1
        self.tables[table][(key, value)] = self._ts[table]        # PRIMARY KEY (key, value): a repeat overwrites ts

    def add(self, key: bytes, band_keys: Sequence[bytes]) -> None:
        """What ``MinHashLSH._insert`` writes for one document (lsh.py:344-347 over storage.py:771-774, :813-816)."""
        if len(band_keys) != self.b:
            raise ValueError("expected %d band keys, got %d" % (self.b, len(band_keys)))
        for h in band_keys:
            self._put(self.keys_table, key, h)
        for table, h in zip(self.bucket_tables, band_keys):
            self._put(table, h, key)

    def ddl(self) -> List[str]:
        return [CASSANDRA_CREATE_TABLE.format(t) for t in self.bucket_tables + [self.keys_table]]   # lsh.py:191-200 order

    def statements(self) -> Iterator[Tuple[str, Tuple[bytes, bytes, int]]]:
        """``(cql, (key, value, ts))`` per row, tables in creation order, rows in ``ts`` order."""
        for t in self.bucket_tables + [self.keys_table]:
            cql = CASSANDRA_INSERT.format(t)
            for (key, value), ts in sorted(self.tables[t].items(), key=lambda kv: kv[1]):
                yield cql, (key, value, ts)

    def write(self, session) -> int:
        """Create the tables and insert every row through a cassandra-driver style session (``execute``, ``prepare``)."""
        for q in self.ddl():
            session.execute(q)
        prepared, n = {}, 0
        for cql, params in self.statements():
            st = prepared.get(cql)
            if st is None:
                st = prepared[cql] = session.prepare(cql)
            session.execute(st, params)
            n += 1
        return n

    def canonical(self):
        """Comparable form: per table the (key, value) rows in ``ts`` order (what the golden fixture stores)."""
        return {t: [kv for kv, _ in sorted(rows.items(), key=lambda x: x[1])] for t, rows in self.tables.items()}

    def __eq__(self, other):
        return isinstance(other, CassandraLayout) and self.canonical() == other.canonical()


def cassandra_layout(keys: Sequence[Hashable], signatures, b: int, r: int, basename: bytes, prepickle: bool = False,
                     hashfunc: Optional[Callable[[bytes], bytes]] = None) -> CassandraLayout:
    """Cassandra layout of ``MinHashLSH(params=(b, r), storage_config={"type": "cassandra", "basename": basename, ...})``
    after ``insert(keys[i], signature i)`` for every row of ``signatures``; band keys from one ``dsk_band_keys`` launch."""
    from . import codec
    from .lsh import _signature_matrix
    sig = _signature_matrix(signatures) if isinstance(signatures, np.ndarray) or not hasattr(signatures, "is_cuda") else signatures
    n = int(sig.shape[0])
    if len(keys) != n:
        raise ValueError("keys and signatures differ in length")
    layout = CassandraLayout(basename, b)
    if n == 0:
        return layout
    width = 8 * r
    raw = codec.band_keys(sig, b, r).cpu().numpy().reshape(n, b * width).tobytes()
    for i, key in enumerate(keys):
        if prepickle:
            key = pickle.dumps(key)
        elif not isinstance(key, bytes):
            raise TypeError(f"prepickle=False requires bytes keys for non-dict storage, got {type(key).__name__}. "
                            "Either pass bytes keys or use prepickle=True for automatic serialization.")
        base = i * b * width
        hs = [raw[base + j * width: base + (j + 1) * width] for j in range(b)]
        if hashfunc is not None:
            hs = [hashfunc(h) for h in hs]
        layout.add(key, hs)
    return layout
