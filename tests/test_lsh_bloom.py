"""MinHashLSHBloom and the Redis wire layout (f4).

CPU part: constructor contract of the reference (datasketch/lsh_bloom.py:244-275), the C-ABI entry points on the emulated
library.  GPU part: dsk_band_sums against the keys the reference's own BloomTable.insert produced (tests/golden/bloom.npz)
and against the numpy restatement on random matrices; the device Bloom tables (no false negatives, false-positive rate
within the requested bound, answers equal to exact membership up to that rate); storage_export against the database
state the reference's own Redis storage code wrote (tests/golden/storage.npz)."""
import ctypes
import pickle
import shutil
import warnings

import numpy as np
import pytest

from oracle import oracle_np as o


# ---- CPU -----------------------------------------------------------------------------------------------------------
def test_constructor_validation_like_the_reference():
    import datasketch_b200 as dsk
    for kw in (dict(threshold=1.5, n=10, fp=0.1), dict(threshold=-0.1, n=10, fp=0.1), dict(num_perm=1, n=10, fp=0.1),
               dict(n=None, fp=0.1), dict(n=0, fp=0.1), dict(n=10, fp=None), dict(n=10, fp=1.0), dict(n=10, fp=0.0),
               dict(n=10, fp=0.1, weights=(0.5, 0.6)), dict(n=10, fp=0.1, weights=(-0.1, 1.1)),
               dict(n=10, fp=0.1, params=(20, 10), num_perm=128), dict(n=10, fp=0.1, params=(1, 4))):
        with pytest.raises(ValueError), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            dsk.MinHashLSHBloom(**kw)


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_band_sums_and_bloom_tables_on_the_emulated_library(emu_lib, golden):
    lib = emu_lib
    sig = np.ascontiguousarray(golden("lsh")["sig"])
    g = golden("bloom")
    b, r = (int(x) for x in g["b_r"])
    n, k = sig.shape
    keys = np.zeros((n, b), dtype=np.uint64)
    assert lib.dsk_band_sums(sig.ctypes.data, n, k, b, r, keys.ctypes.data, None) == 0
    assert np.array_equal(keys[:len(g["keys_inserted"])], g["keys_inserted"])
    assert lib.dsk_band_sums(sig.ctypes.data, n, k, 20, 10, keys.ctypes.data, None) != 0          # b * r > num_perm
    n_bits, n_hashes = 20011, 7
    words = (n_bits + 31) // 32
    bits = np.zeros((b, words), dtype=np.uint32)
    n_ins = len(g["keys_inserted"])
    assert lib.dsk_bloom_insert(sig.ctypes.data, n_ins, k, b, r, bits.ctypes.data, words, n_bits, n_hashes, None) == 0
    hit = np.zeros(n, dtype=np.uint8)
    assert lib.dsk_bloom_query(sig.ctypes.data, n, k, b, r, bits.ctypes.data, words, n_bits, n_hashes, hit.ctypes.data, None) == 0
    want = g["query_all"]
    assert hit[want].all()                                    # no false negatives
    assert (hit.astype(bool) & ~want).mean() < 0.05           # a handful of false positives at most
    assert lib.dsk_bloom_query(sig.ctypes.data, n, k, b, r, bits.ctypes.data, words, 0, n_hashes, hit.ctypes.data, None) != 0
    assert lib.dsk_bloom_insert(sig.ctypes.data, n, k, b, r, bits.ctypes.data, 1, n_bits, n_hashes, None) != 0   # table too small


# ---- GPU -----------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_band_sums_vs_reference_keys_and_oracle(golden):
    import datasketch_b200 as dsk
    sig = golden("lsh")["sig"]
    g = golden("bloom")
    b, r = (int(x) for x in g["b_r"])
    got = dsk.codec.band_sums(sig, b, r).cpu().numpy().view(np.uint64)
    assert np.array_equal(got[:len(g["keys_inserted"])], g["keys_inserted"])
    rs = np.random.RandomState(3)
    for k, bb, rr in ((128, 9, 13), (256, 17, 15), (64, 64, 1), (100, 3, 33), (2048, 2, 1024)):
        m = rs.randint(0, 1 << 32, size=(257, k), dtype=np.uint64).astype(np.uint32)
        m[0] = 0xFFFFFFFF
        got = dsk.codec.band_sums(m, bb, rr).cpu().numpy().view(np.uint64)
        want = np.stack([o.bloom_band_keys(row.astype(np.uint64), bb, rr) for row in m])
        assert np.array_equal(got, want), (k, bb, rr)


@pytest.mark.gpu
def test_lsh_bloom_api_and_false_positive_rate(golden, tmp_path):
    import datasketch_b200 as dsk
    sig = golden("lsh")["sig"]
    g = golden("bloom")
    for thr, k, bb, rr in g["params"]:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            l = dsk.MinHashLSHBloom(threshold=float(thr), num_perm=int(k), n=100, fp=0.01)
        assert (l.b, l.r) == (int(bb), int(rr))
    with pytest.warns(RuntimeWarning):
        bl = dsk.MinHashLSHBloom(threshold=0.8, num_perm=128, n=1000, fp=0.001)
    assert (bl.b, bl.r) == tuple(int(x) for x in g["b_r"])
    n_ins = len(g["keys_inserted"])
    for row in sig[:5]:                                            # the reference's single-object API
        bl.insert(dsk.LeanMinHash(seed=1, hashvalues=row.astype(np.uint64)))
    bl.insert_batch(sig[5:n_ins])
    got = bl.query_batch(sig)
    want = g["query_all"]                                          # exact-membership answers of the reference
    assert got[want].all() and (got & ~want).sum() <= 2
    assert bl.query(dsk.LeanMinHash(seed=1, hashvalues=sig[0].astype(np.uint64))) is True
    with pytest.raises(ValueError):
        bl.query(dsk.MinHash(num_perm=64))
    with pytest.warns(RuntimeWarning):
        bl.sync()
    assert len(bl.hashtables) == bl.b
    t = dsk.lsh_bloom.BloomTable(10, 0.01, band_size=3, fname=str(tmp_path / "t.bf"))     # test/test_lshbloom.py:11-48
    t.insert(np.array([2, 3, 31], dtype=np.uint32))
    assert t.query(np.array([2, 3, 31], dtype=np.uint32)) and not t.query(np.array([2, 3, 30], dtype=np.uint32))
    with pytest.raises(RuntimeError):
        t.insert(np.array([2, 2], dtype=np.uint32))
    t.sync()
    assert dsk.lsh_bloom.BloomTable(10, 0.01, band_size=3, fname=str(tmp_path / "t.bf")).query([2, 3, 31])
    # false-positive rate of the tables at the design point: n random documents in, 20 000 random queries
    rs = np.random.RandomState(11)
    n, fp = 20_000, 0.01
    big = dsk.MinHashLSHBloom(threshold=0.8, num_perm=128, n=n, fp=fp, save_dir=str(tmp_path / "idx"))
    inserted = rs.randint(0, 1 << 32, size=(n, 128), dtype=np.uint64).astype(np.uint32)
    big.insert_batch(inserted)
    assert big.query_batch(inserted).all()
    rate = big.query_batch(rs.randint(0, 1 << 32, size=(20_000, 128), dtype=np.uint64).astype(np.uint32)).mean()
    assert rate <= 1.0 - (1.0 - fp) ** big.b + 0.01, rate          # any-of-b bands, each table at <= fp
    big.sync()
    again = dsk.MinHashLSHBloom(threshold=0.8, num_perm=128, n=n, fp=fp, save_dir=str(tmp_path / "idx"))
    assert again.query_batch(inserted[:500]).all()


@pytest.mark.gpu
def test_redis_layout_equals_what_the_reference_wrote(golden):
    import datasketch_b200 as dsk
    from datasketch_b200 import storage_export as se
    sig = golden("lsh")["sig"]
    s = golden("storage")
    for name, mk in (("pickled", lambda i: ("doc", i)), ("bytes", lambda i: b"k%04d" % i)):
        state = pickle.loads(s[name + "_state"].tobytes())
        b, r = (int(x) for x in s[name + "_b_r"])
        keys = [mk(i) for i in range(120)]
        lay = se.redis_layout(keys, sig[:120], b, r, b"gpuidx", prepickle=(name == "pickled"))
        assert lay.canonical() == state
        # the same index built through the drop-in class, then exported
        lsh = dsk.MinHashLSH(threshold=0.8, num_perm=128)
        assert (lsh.b, lsh.r) == (b, r)
        lsh.insert_batch(keys, sig[:120])
        assert lsh.export_redis(b"gpuidx", prepickle=(name == "pickled")).canonical() == state
        # replayed into a client with the redis-py pipeline interface, the reference would read it back
        from oracle import fake_backends as fb
        fb.reset()
        assert lay.write(_FakeClient(fb)) == sum(1 for _ in lay.commands())
        assert {"hash": fb.DB["hash"], "list": fb.DB["list"], "set": {k: sorted(v) for k, v in fb.DB["set"].items()}} == state
    with pytest.raises(TypeError):
        se.redis_layout(["not-bytes"], sig[:1], 9, 13, b"x", prepickle=False)


class _FakeClient:
    def __init__(self, fb):
        self.r = fb._Redis()

    def pipeline(self):
        outer = self

        class P:
            def hset(self, *a): outer.r.hset(*a)
            def rpush(self, *a): outer.r.rpush(*a)
            def sadd(self, *a): outer.r.sadd(*a)
            def execute(self): return []
        return P()
