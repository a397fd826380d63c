"""GPU parity: device SHA1 token hashing (the reference's default hashfunc) vs hashlib, and
MinHash.bulk on byte tokens through it vs the reference-made fixtures."""
import hashlib
import struct

import numpy as np
import pytest

from oracle import oracle_np as o

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dsk():
    import datasketch_b200
    return datasketch_b200


def test_sha1_tokens_vs_hashlib(dsk):
    rs = np.random.RandomState(1)
    toks = [b"", b"a", b"abc", b"Hello", b"x" * 55, b"y" * 56, b"z" * 63, b"w" * 64, b"v" * 65, b"u" * 119, b"t" * 120,
            b"s" * 1000] + [bytes(rs.randint(0, 256, size=rs.randint(0, 200)).astype(np.uint8)) for _ in range(2000)]
    h32 = dsk.engine.sha1_hash_tokens_device(toks).cpu().numpy().view(np.uint32)
    want32 = np.array([struct.unpack("<I", hashlib.sha1(t).digest()[:4])[0] for t in toks], dtype=np.uint32)
    assert np.array_equal(h32, want32)
    assert h32[3] == dsk.sha1_hash32(b"Hello") == o.sha1_hash32(b"Hello")
    h64 = dsk.engine.sha1_hash_tokens_device(toks, out_u64=True).cpu().numpy().view(np.uint64)
    want64 = np.array([struct.unpack("<Q", hashlib.sha1(t).digest()[:8])[0] for t in toks], dtype=np.uint64)
    assert np.array_equal(h64, want64) and h64[3] == dsk.sha1_hash64(b"Hello")


def test_bulk_default_hashfunc_bytes_tokens(dsk, golden):
    g = golden("minhash")
    data = [f"token-{i}".encode() for i in range(1000)]
    (m,) = dsk.MinHash.bulk([data], num_perm=256, seed=7)
    assert np.array_equal(m.hashvalues, g["sha1_k256_s7_n1000"]) and m.hashvalues.dtype == np.uint64
    docs = [[b"Hello"], [], data[:3], [b"token1", b"token2", b"token3"]]
    ms = dsk.MinHash.bulk(docs, num_perm=4, seed=1)
    assert ms[0].hashvalues.tolist() == [734825475, 960773806, 359816889, 342714745]
    assert ms[1].is_empty()
    for doc, m in zip(docs, ms):
        ref = dsk.MinHash(num_perm=4, seed=1)
        ref.update_batch(doc)
        assert np.array_equal(ref.hashvalues, m.hashvalues)
    with pytest.raises(TypeError):
        dsk.MinHash.bulk([["not-bytes"]], num_perm=4)      # hashlib raises TypeError for str, so do we


def test_bulk_bytes_like_tokens_and_batches(dsk):
    """bytearray / memoryview tokens, generator documents, several device batches, and a token type
    whose len() is not its byte size (falls back to the per-token hash function): all rows must equal
    the per-document update_batch result."""
    import array
    rs = np.random.RandomState(5)
    raw = [[bytes(rs.randint(0, 256, size=rs.randint(0, 40)).astype(np.uint8)) for _ in range(rs.randint(0, 30))]
           for _ in range(300)]
    docs = []
    for i, d in enumerate(raw):
        if i % 3 == 0:
            docs.append([bytearray(t) for t in d])
        elif i % 3 == 1:
            docs.append(memoryview(t) for t in d)          # a generator of memoryviews
        else:
            docs.append(tuple(d))
    ms = list(dsk.MinHash.generator(docs, batch_docs=64, num_perm=64, seed=3))
    assert len(ms) == len(raw)
    perms = o.init_permutations(64, 3)
    for d, m in zip(raw, ms):
        want = o.update_batch(o.init_hashvalues(64), [o.sha1_hash32(t) for t in d], perms)
        assert np.array_equal(m.hashvalues, want)
    odd = [[array.array("I", [1, 2, 3]), b"abc"]]          # hashlib hashes the 12 raw bytes of the array
    (m,) = dsk.MinHash.bulk(odd, num_perm=16, seed=2)
    want = o.update_batch(o.init_hashvalues(16), [o.sha1_hash32(t) for t in odd[0]], o.init_permutations(16, 2))
    assert np.array_equal(m.hashvalues, want)

