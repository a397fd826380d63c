"""MinHashLSHForest ("next" row of SURVEY.md 8f): host class vs reference-made fixtures (CPU);
batch key generation and the device-resident forest vs the same fixtures (GPU)."""
import pickle

import numpy as np
import pytest


@pytest.fixture(scope="module")
def dsk():
    import datasketch_b200
    return datasketch_b200


def _lean(dsk, row):
    return dsk.LeanMinHash(seed=1, hashvalues=row.astype(np.uint64))


def _check(g, l, topk, results):
    ptr, idx = g[f"l{l}_k{topk}_ptr"], g[f"l{l}_k{topk}_idx"]
    for q, got in enumerate(results):
        assert sorted(got) == idx[ptr[q]:ptr[q + 1]].tolist(), (l, topk, q)


@pytest.mark.parametrize("l", [8, 32])
def test_forest_host_matches_reference(dsk, golden, l):
    g, sig = golden("forest"), golden("lsh")["sig"]
    f = dsk.MinHashLSHForest(num_perm=128, l=l)
    assert f.is_empty()
    for i, row in enumerate(sig):
        f.add(i, _lean(dsk, row))
    assert f.is_empty() and 5 in f and 1000 not in f
    with pytest.raises(ValueError):
        f.add(0, _lean(dsk, sig[0]))
    f.index()
    assert not f.is_empty()
    for topk in (1, 5, 20):
        _check(g, l, topk, [f.query(_lean(dsk, row), topk) for row in sig[:120]])
    if l == 8:
        assert b"".join(f.keys[0]) == g["keys_doc0"].tobytes()
        assert np.array_equal(f.get_minhash_hashvalues(3), g["hashvalues_doc3"])
    with pytest.raises(ValueError):
        f.query(_lean(dsk, sig[0]), 0)
    with pytest.raises(ValueError):
        f.query(dsk.LeanMinHash(seed=1, hashvalues=np.arange(8, dtype=np.uint64)), 3)
    with pytest.raises(KeyError):
        f.get_minhash_hashvalues("nope")
    p = pickle.loads(pickle.dumps(f))
    assert sorted(p.query(_lean(dsk, sig[0]), 5)) == sorted(f.query(_lean(dsk, sig[0]), 5))
    for bad in [dict(l=0), dict(num_perm=0), dict(num_perm=4, l=8)]:
        with pytest.raises(ValueError):
            dsk.MinHashLSHForest(**bad)


@pytest.mark.gpu
@pytest.mark.parametrize("l", [8, 32])
def test_forest_batch_and_device_match_reference(dsk, golden, l):
    g, sig = golden("forest"), golden("lsh")["sig"]
    f = dsk.MinHashLSHForest(num_perm=128, l=l)
    f.add_batch(list(range(len(sig))), sig)           # keys from one dsk_band_keys launch
    f.index()
    if l == 8:
        assert b"".join(f.keys[0]) == g["keys_doc0"].tobytes()
    dev = dsk.GpuLSHForest(num_perm=128, l=l)
    dev.add(sig[:100])
    dev.add(sig[100:])
    dev.index()
    assert len(dev) == len(sig) and not dev.is_empty()
    for topk in (1, 5, 20):
        _check(g, l, topk, [f.query(_lean(dsk, row), topk) for row in sig[:120]])
        res = dev.query(sig[:120], topk)
        _check(g, l, topk, [[int(x) for x in row if x >= 0] for row in res])


@pytest.mark.gpu
def test_device_forest_vs_host_class_random(dsk):
    # low-entropy signatures: long equal prefixes, large buckets, early exits at every r
    rs = np.random.RandomState(4)
    for k_perm, l in [(64, 4), (128, 16), (40, 40)]:
        n = 1500
        sig = rs.randint(0, 3, size=(n, k_perm)).astype(np.uint32)
        sig[rs.randint(0, n, 100)] = sig[rs.randint(0, n, 100)]
        f = dsk.MinHashLSHForest(num_perm=k_perm, l=l)
        for i, row in enumerate(sig):
            f.add(i, dsk.LeanMinHash(seed=1, hashvalues=row.astype(np.uint64)))
        f.index()
        dev = dsk.GpuLSHForest(num_perm=k_perm, l=l)
        dev.add(sig)
        dev.index()
        q = np.concatenate([sig[:80], rs.randint(0, 3, size=(20, k_perm)).astype(np.uint32)])
        for topk in (1, 7, 50):
            res = dev.query(q, topk)
            for j, row in enumerate(q):
                want = sorted(f.query(dsk.LeanMinHash(seed=1, hashvalues=row.astype(np.uint64)), topk))
                assert sorted(int(x) for x in res[j] if x >= 0) == want, (k_perm, l, topk, j)
