"""CPU: MinHashLSH host logic (parameter optimiser, dict buckets, removal, merge, sessions,
error behaviour) against fixtures produced by the reference -- no GPU needed because the
signatures come from the fixtures (LeanMinHash(seed, hashvalues) carries its state)."""
import pickle

import numpy as np
import pytest


@pytest.fixture(scope="module")
def dsk():
    import datasketch_b200
    return datasketch_b200


def _lean(dsk, row):
    return dsk.LeanMinHash(seed=1, hashvalues=row.astype(np.uint64))


def test_params_match_reference(dsk, golden):
    g = golden("lsh")
    for thr, k, w0, w1, b, r in g["params"]:
        if k > 128:
            continue
        lsh = dsk.MinHashLSH(threshold=float(thr), num_perm=int(k), weights=(float(w0), float(w1)))
        assert (lsh.b, lsh.r) == (int(b), int(r)) and lsh.h == int(k)
        assert lsh.hashranges == [(i * lsh.r, (i + 1) * lsh.r) for i in range(lsh.b)]
    # test/test_lsh.py:21-28
    l1, l2 = dsk.MinHashLSH(threshold=0.8), dsk.MinHashLSH(threshold=0.8, weights=(0.2, 0.8))
    assert l1.is_empty() and l1.b < l2.b and l1.r > l2.r
    # docstring value of the reference (lsh.py:98-100)
    assert (dsk.MinHashLSH(threshold=0.9, num_perm=128).b, dsk.MinHashLSH(threshold=0.9, num_perm=128).r) == (5, 25)


def test_constructor_errors(dsk):
    L = dsk.MinHashLSH
    for kw in [dict(threshold=1.5), dict(threshold=-0.1), dict(num_perm=1), dict(weights=(1.5, -0.5)),
               dict(weights=(0.3, 0.3)), dict(params=(20, 10), num_perm=128), dict(params=(1, 4), num_perm=16),
               dict(storage_config={"type": "redis", "redis": {}})]:
        with pytest.raises(ValueError):
            L(**kw)


def test_insert_query_candidates_equal_reference(dsk, golden):
    g = golden("lsh")
    sig = g["sig"]
    lsh = dsk.MinHashLSH(threshold=0.8, num_perm=128)
    assert [lsh.b, lsh.r] == g["b_r"].tolist()
    ms = [_lean(dsk, row) for row in sig]
    for i, m in enumerate(ms):
        lsh.insert(i, m)
    assert not lsh.is_empty() and 0 in lsh and 299 in lsh and 300 not in lsh
    keys0 = lsh.keys[0]
    assert all(len(H) == 8 * lsh.r for H in keys0)
    assert np.array_equal(np.frombuffer(b"".join(keys0), np.uint8).reshape(lsh.b, 8 * lsh.r), g["keys_doc0"])
    for i, H in enumerate(lsh.keys[0]):
        assert 0 in lsh.hashtables[i][H]
    ptr, idx = g["query_ptr"], g["query_idx"]
    for i, m in enumerate(ms):
        assert sorted(lsh.query(m)) == idx[ptr[i]:ptr[i + 1]].tolist()
    counts = sorted(c for t in lsh.get_counts() for c in t.values())
    assert counts == g["bucket_counts_sorted"].tolist()
    with pytest.raises(ValueError):
        lsh.insert(0, ms[0])                      # duplicate key (lsh.py:342-343)
    lsh.insert(0, ms[0], check_duplication=False)
    with pytest.raises(ValueError):
        lsh.insert("x", dsk.LeanMinHash(seed=1, hashvalues=np.arange(18, dtype=np.uint64)))
    with pytest.raises(ValueError):
        lsh.query(dsk.LeanMinHash(seed=1, hashvalues=np.arange(18, dtype=np.uint64)))
    # query buffer == query (test/test_lsh.py:109-125)
    lsh.add_to_query_buffer(ms[3])
    assert set(lsh.collect_query_buffer()) == set(lsh.query(ms[3]))
    assert lsh._query_b(ms[3], lsh.b) == set(lsh.query(ms[3]))
    with pytest.raises(ValueError):
        lsh._query_b(ms[3], lsh.b + 1)


def test_abc_pinned_candidates(dsk, golden):
    g = golden("lsh")
    b, r = [int(x) for x in g["abc_b_r"]]
    lsh = dsk.MinHashLSH(threshold=0.5, num_perm=32)
    assert (lsh.b, lsh.r) == (b, r)
    ms = [dsk.LeanMinHash(seed=1, hashvalues=row) for row in g["abc_sig"]]
    for i, m in enumerate(ms):
        lsh.insert(i, m)
    assert set(lsh.query(ms[0])) == {0, 1} == set(g["abc_query0"].tolist())


def test_remove_merge_sessions_pickle(dsk, golden):
    sig = golden("lsh")["sig"]
    ms = [_lean(dsk, row) for row in sig[:60]]
    lsh = dsk.MinHashLSH(threshold=0.8, num_perm=128)
    with lsh.insertion_session() as s:
        for i, m in enumerate(ms[:30]):
            s.insert(i, m)
    other = dsk.MinHashLSH(threshold=0.8, num_perm=128)
    for i, m in enumerate(ms[30:], start=30):
        other.insert(i, m)
    lsh.merge(other)
    assert all(i in lsh for i in range(60))
    with pytest.raises(ValueError):
        lsh.merge(other, check_overlap=True)
    with pytest.raises(ValueError):
        lsh.merge(dsk.MinHashLSH(threshold=0.5, num_perm=128))
    # reference: identical signatures 0 and 2 are mutual candidates
    assert {0, 2} <= set(lsh.query(ms[0]))
    lsh.remove(2)
    assert 2 not in lsh and 2 not in lsh.query(ms[0])
    with pytest.raises(ValueError):
        lsh.remove(2)
    for table in lsh.hashtables:          # no empty buckets left behind (lsh.py:521-523)
        assert all(len(table[H]) > 0 for H in table)
    with lsh.deletion_session() as s:
        s.remove(5)
    assert 5 not in lsh
    sub = lsh.get_subset_counts(0, 1)
    assert len(sub) == lsh.b and all(sum(t.values()) == 2 for t in sub)
    p = pickle.loads(pickle.dumps(lsh))
    assert sorted(p.query(ms[0])) == sorted(lsh.query(ms[0]))


def test_prepickle_hashfunc_and_weighted(dsk, golden):
    sig = golden("lsh")["sig"]
    ms = [_lean(dsk, row) for row in sig[:9]]
    lsh = dsk.MinHashLSH(threshold=0.8, num_perm=128, prepickle=True)
    for i, m in enumerate(ms):
        lsh.insert(("doc", i), m)
    assert ("doc", 0) in lsh and ("doc", 2) in lsh.query(ms[0])
    import hashlib
    lh = dsk.MinHashLSH(threshold=0.8, num_perm=128, hashfunc=lambda b: hashlib.md5(b).digest())
    for i, m in enumerate(ms):
        lh.insert(i, m)
    assert all(len(H) == 16 for H in lh.keys[0]) and sorted(lh.query(ms[0])) == sorted(
        int(k[1]) for k in lsh.query(ms[0]))
    # weighted signatures ((k, t) int64 rows) go through the same banding (test/test_lsh.py:461-551)
    w = dsk.WeightedMinHash(1, np.arange(8, dtype=np.int64).reshape(4, 2))
    lw = dsk.WeightedMinHashLSH(threshold=0.5, num_perm=4, params=(2, 2))
    lw.insert("w", w)
    assert lw.query(w) == ["w"] and all(len(H) == 2 * 2 * 8 for H in lw.keys["w"])


def test_cassandra_layout_equals_what_the_reference_wrote(dsk, golden):
    """storage_export.CassandraLayout / MinHashLSH.export_cassandra against the rows the reference's OWN Cassandra storage
    code (datasketch/storage.py:262-819) wrote on an in-memory cassandra-driver stand-in (oracle/gen_golden.py,
    tests/golden/storage_cassandra.npz): the same tables, the same (key, value) rows, in the same ts order; replayed into a
    session with the driver's interface, the state is the reference's again."""
    from datasketch_b200 import storage_export as se
    sig = golden("lsh")["sig"]
    s = golden("storage_cassandra")
    for name, mk, prepickle in (("pickled", lambda i: ("doc", i), True), ("bytes", lambda i: b"k%04d" % i, False)):
        state = pickle.loads(s[name + "_state"].tobytes())
        ddl = pickle.loads(s[name + "_ddl"].tobytes())
        b, r = (int(x) for x in s[name + "_b_r"])
        keys = [mk(i) for i in range(120)]
        lsh = dsk.MinHashLSH(threshold=0.8, num_perm=128)
        assert (lsh.b, lsh.r) == (b, r)
        for key, row in zip(keys, sig):
            lsh.insert(key, dsk.LeanMinHash(seed=1, hashvalues=row.astype(np.uint64)))
        lay = lsh.export_cassandra(b"gpuidx", prepickle=prepickle)
        assert lay.canonical() == state
        assert sorted(lay.ddl()) == sorted(ddl) and lay.ddl() == ddl
        assert lay.keys_table == "lsh_gpuidx_keys" and lay.bucket_tables[1] == "lsh_gpuidx_bucket_0001"
        # the layout built from explicit band keys (what storage_export.cassandra_layout does after one dsk_band_keys launch)
        lay2 = se.CassandraLayout(b"gpuidx", b)
        for key, row in zip(keys, sig):
            hs = [row[i * r:(i + 1) * r].astype(">u8").tobytes() for i in range(b)]
            lay2.add(pickle.dumps(key) if prepickle else key, hs)
        assert lay2 == lay
        # replay through a driver-style session: the rows land where the reference's code put them
        from oracle import fake_backends as fb
        fb.reset()
        n = lay.write(fb._Session())
        assert n == sum(len(v) for v in state.values())
        got = {t: [kv for kv, _ in sorted(rows.items(), key=lambda x: x[1])] for t, rows in fb.CQL["tables"].items()}
        assert got == state
    lsh = dsk.MinHashLSH(threshold=0.8, num_perm=128)
    lsh.insert("not-bytes", dsk.LeanMinHash(seed=1, hashvalues=sig[0].astype(np.uint64)))
    with pytest.raises(TypeError):
        lsh.export_cassandra(b"x")                      # prepickle defaults to False for Cassandra: keys must be bytes
    assert len(lsh.export_cassandra(b"x", prepickle=True).tables["lsh_x_keys"]) == lsh.b
