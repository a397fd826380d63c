"""GPU parity: batch band keys through MinHashLSH.insert_batch/query_batch and the device-resident
GpuLSH index vs the reference's query results (fixtures) and the dict oracle."""
import numpy as np
import pytest

from oracle import oracle_np as o

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dsk():
    import datasketch_b200
    return datasketch_b200


def test_insert_batch_query_batch_equal_reference(dsk, golden):
    g = golden("lsh")
    sig = g["sig"]
    lsh = dsk.MinHashLSH(threshold=0.8, num_perm=128)
    lsh.insert_batch(list(range(len(sig))), sig)
    keys0 = lsh.keys[0]
    assert np.array_equal(np.frombuffer(b"".join(keys0), np.uint8).reshape(lsh.b, 8 * lsh.r), g["keys_doc0"])
    ptr, idx = g["query_ptr"], g["query_idx"]
    res = lsh.query_batch(sig)
    for i in range(len(sig)):
        assert sorted(res[i]) == idx[ptr[i]:ptr[i + 1]].tolist()
    # batch-inserted and per-item paths interoperate: same buckets byte for byte
    one = dsk.MinHashLSH(threshold=0.8, num_perm=128)
    for i, row in enumerate(sig):
        one.insert(i, dsk.LeanMinHash(seed=1, hashvalues=row.astype(np.uint64)))
    for t1, t2 in zip(lsh.hashtables, one.hashtables):
        assert t1.itemcounts() == t2.itemcounts()
    with pytest.raises(ValueError):
        lsh.insert_batch([0], sig[:1])          # duplicate key
    with pytest.raises(ValueError):
        lsh.insert_batch([1000], sig[:1, :64])  # wrong signature length


def test_gpu_lsh_candidates_equal_reference(dsk, golden):
    g = golden("lsh")
    sig = g["sig"]
    ix = dsk.GpuLSH(threshold=0.8, num_perm=128, capacity=1000)
    assert [ix.b, ix.r] == g["b_r"].tolist()
    ix.insert(sig[:100])
    ix.insert(sig[100:])                          # incremental insert
    assert len(ix) == len(sig)
    ptr, idx = ix.query(sig)
    gp, gi = g["query_ptr"], g["query_idx"]
    assert ptr.dtype == np.int64 and ptr[0] == 0 and ptr[-1] == len(idx)
    for q in range(len(sig)):
        got = idx[ptr[q]:ptr[q + 1]]
        assert len(set(got.tolist())) == len(got)             # distinct (set union over bands)
        assert sorted(got.tolist()) == gi[gp[q]:gp[q + 1]].tolist()
    with pytest.raises(ValueError):
        dsk.GpuLSH(threshold=0.8, num_perm=128, capacity=10).insert(sig)   # over capacity
    with pytest.raises(ValueError):
        ix.query(sig[:, :64])


@pytest.mark.parametrize("k,params", [(128, (32, 4)), (256, (17, 15)), (64, (40, 1)), (16, (2, 8))])
def test_gpu_lsh_vs_dict_oracle_random(dsk, k, params):
    # low-entropy signatures create big buckets, multi-band matches and > 32 bands
    rs = np.random.RandomState(k)
    n = 3000
    sig = rs.randint(0, 3 if params[1] > 1 else 50, size=(n, k)).astype(np.uint32)
    sig[rs.randint(0, n, 200)] = sig[rs.randint(0, n, 200)]   # exact duplicates
    b, r = params
    ix = dsk.GpuLSH(num_perm=k, params=params, capacity=n)
    ix.insert(sig)
    ref = o.DictLSH(k, b, r)
    for i, row in enumerate(sig):
        ref.insert(i, row.astype(np.uint64))
    q = np.concatenate([sig[:300], rs.randint(0, 3, size=(50, k)).astype(np.uint32)])
    ptr, idx = ix.query(q)
    for j in range(len(q)):
        assert sorted(idx[ptr[j]:ptr[j + 1]].tolist()) == sorted(ref.query(q[j].astype(np.uint64)))
    keys = ix.query_keys(q[:5])
    assert [sorted(x) for x in keys] == [sorted(idx[ptr[j]:ptr[j + 1]].tolist()) for j in range(5)]


def test_exclusive_scan(dsk):
    import torch
    from datasketch_b200 import _native as nv
    for n in [0, 1, 1023, 1024, 1025, 50_000, 1_234_567]:
        x = torch.randint(0, 1000, (max(n, 1),), dtype=torch.int64, device="cuda")
        out = torch.empty((n + 1,), dtype=torch.int64, device="cuda")
        scratch = torch.empty((n // 1024 + 2,), dtype=torch.int64, device="cuda")
        nv.check(nv.load().dsk_exclusive_scan(x.data_ptr(), n, out.data_ptr(), scratch.data_ptr(),
                                              torch.cuda.current_stream().cuda_stream))
        want = torch.cat([torch.zeros(1, dtype=torch.int64, device="cuda"), torch.cumsum(x[:n], 0)])
        assert torch.equal(out, want), n


def test_end_to_end_minhash_to_lsh(dsk):
    # tokens -> bulk signatures (GPU) -> device LSH: planted near-duplicates are found, like the reference flow
    rs = np.random.RandomState(2)
    docs = [rs.randint(0, 2 ** 32, size=rs.randint(30, 90), dtype=np.uint64) for _ in range(500)]
    for i in range(0, 500, 5):
        d = docs[i].copy()
        d[: max(1, len(d) // 20)] = rs.randint(0, 2 ** 32, size=max(1, len(d) // 20), dtype=np.uint64)
        docs[i + 1] = d
    tok, off = dsk.engine.pack_docs([[int(t) for t in d] for d in docs])
    P = o.init_permutations(128, 1)
    sig = dsk.engine.bulk_signatures(tok, off, P)
    ix = dsk.GpuLSH(threshold=0.8, num_perm=128, capacity=500)
    ix.insert(sig, keys=[f"doc{i}" for i in range(500)])
    res = ix.query_keys(sig)
    ref = o.DictLSH(128, ix.b, ix.r)
    for i, row in enumerate(sig):
        ref.insert(f"doc{i}", row.astype(np.uint64))
    for i in range(500):
        assert set(res[i]) == ref.query(sig[i].astype(np.uint64))
    found = sum(1 for i in range(0, 500, 5) if f"doc{i + 1}" in res[i])
    assert found >= 80   # ~95 % similar pairs collide in some band with overwhelming probability


@pytest.mark.parametrize("k,params", [(128, (9, 13)), (256, (17, 15)), (64, (9, 7)), (300, (20, 15))])
def test_fused_insert_from_tokens_equals_the_two_step_flow(dsk, k, params):
    """``GpuLSH.insert_tokens`` (dsk_lsh_insert_tokens: the signature kernel with the fused insert epilogue; num_perm = 300
    and 64-bit tokens take its two-kernel route) against signatures-then-insert and against the reference's dict buckets
    (oracle DictLSH = lsh.py:326-347, :370-432): same candidate sets for every query."""
    import torch
    rs = np.random.RandomState(k)
    n = 1500
    docs = []
    for i in range(n):
        if i % 4 == 1 and i > 8:
            d = docs[int(rs.randint(0, i))].copy()
            if len(d):
                d[rs.randint(0, len(d), size=max(1, len(d) // 30))] = rs.randint(0, 2 ** 32, dtype=np.uint64)
            docs.append(d)
        elif i % 9 == 5:
            docs.append(docs[i - 1].copy())
        else:
            docs.append(rs.randint(0, 2 ** 32, size=int(rs.choice([0, 5, 64, 100, 256, 700, 5000 if i % 500 == 0 else 33])),
                                   dtype=np.uint64).astype(np.uint32))
    docs = [np.asarray(d, dtype=np.uint32) for d in docs]
    P = o.init_permutations(k, 3)
    kw = dict(params=params)
    for u64 in (False, True):
        fused = dsk.GpuLSH(num_perm=k, capacity=n, **kw)
        plain = dsk.GpuLSH(num_perm=k, capacity=n, **kw)
        ref = o.DictLSH(k, fused.b, fused.r)
        sigs = []
        for lo, hi in ((0, 600), (600, n)):        # two batches: the second lands behind the first
            part = docs[lo:hi]
            off = np.zeros(len(part) + 1, dtype=np.int64)
            np.cumsum([len(d) for d in part], out=off[1:])
            tok = np.ascontiguousarray(np.concatenate(part))
            if u64:
                tok = tok.astype(np.uint64) | (np.uint64(1) << np.uint64(40))
            d_tok = torch.from_numpy(tok.view(np.int64 if u64 else np.int32)).cuda()
            d_off = torch.from_numpy(off).cuda()
            fused.insert_tokens(d_tok, d_off, len(tok), P)
            sig = dsk.engine.bulk_signatures_device(d_tok, d_off, len(tok), P)
            plain.insert(sig)
            sigs.append(sig.cpu().numpy().view(np.uint32))
        sig = np.concatenate(sigs)
        assert len(fused) == n and len(plain) == n
        for i, row in enumerate(sig):
            ref.insert(i, row.astype(np.uint64))
        q = sig[::3]
        p1, i1 = fused.query(q)
        p2, i2 = plain.query(q)
        assert np.array_equal(p1, p2)
        multi = 0
        for j in range(len(q)):
            a = sorted(i1[p1[j]:p1[j + 1]].tolist())
            assert a == sorted(i2[p2[j]:p2[j + 1]].tolist()) and a == sorted(ref.query(q[j].astype(np.uint64))), (u64, j)
            multi += len(a) > 1
        assert multi > 50
    with pytest.raises(ValueError):
        fused.insert_tokens(d_tok, d_off, len(tok), o.init_permutations(k + 1, 3))
    with pytest.raises(ValueError):
        fused.insert_tokens(d_tok, d_off, len(tok), P)      # capacity exceeded
