"""GPU parity tests for the bulk signature path: CUDA kernels (through the C-ABI)
vs fixtures produced by the real reference and vs the oracle on seeded inputs."""
import pickle

import numpy as np
import pytest

from oracle import oracle_clib as oc
from oracle import oracle_np as o

pytestmark = pytest.mark.gpu

KERNELS = ["two_phase", "direct", "exact"]


@pytest.fixture(scope="module")
def dsk():
    import datasketch_b200
    return datasketch_b200


def test_reference_absolute_golden(dsk):
    # test/test_minhash.py:109-115 of the reference
    m = dsk.MinHash(4, 1)
    m.update(b"Hello")
    assert m.hashvalues.tolist() == [734825475, 960773806, 359816889, 342714745]
    assert m.hashvalues.dtype == np.uint64


@pytest.mark.parametrize("kernel", KERNELS)
def test_c1_bulk_golden(dsk, golden, kernel):
    g = golden("minhash")
    tok = g["c1_tokens"].reshape(-1)
    off = np.arange(1001, dtype=np.int64) * 64
    P = o.init_permutations(128, 1)
    sig = dsk.engine.bulk_signatures(tok, off, P, kernel=kernel)
    assert sig.dtype == np.uint32 and np.array_equal(sig, g["c1_sig"])
    sig64 = dsk.engine.bulk_signatures(tok, off, P, kernel=kernel, out_u64=True)
    assert sig64.dtype == np.uint64 and np.array_equal(sig64, g["c1_sig"].astype(np.uint64))


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("k", [4, 33, 100, 128, 256])
def test_ragged_golden(dsk, golden, k, kernel):
    g = golden("minhash")
    tok, off, seed = g[f"rag_k{k}_tokens"], g[f"rag_k{k}_offsets"], int(g[f"rag_k{k}_seed"])
    P = o.init_permutations(k, seed)
    sig = dsk.engine.bulk_signatures(tok, off, P, kernel=kernel)
    assert np.array_equal(sig, g[f"rag_k{k}_sig"])


@pytest.mark.parametrize("kernel", KERNELS)
def test_long_and_small_docs(dsk, golden, kernel):
    g = golden("minhash")
    P = o.init_permutations(128, 1)
    sig = dsk.engine.bulk_signatures(g["long_tokens"], np.array([0, 20000]), P, kernel=kernel)
    assert np.array_equal(sig, g["long_sig"])
    sig = dsk.engine.bulk_signatures(g["small_tokens"], np.array([0, 64]), P, kernel=kernel)
    assert np.array_equal(sig, g["small_sig"])


def test_u64_tokens_golden(dsk, golden):
    g = golden("minhash")
    P = o.init_permutations(64, 3)
    for kernel in ("auto", "two_phase", "exact"):   # auto / two_phase: the general variant of the signature kernel
        sig = dsk.engine.bulk_signatures(g["u64_tokens"], g["u64_offsets"], P, out_u64=True, kernel=kernel)
        assert np.array_equal(sig, g["u64_sig"]), kernel
    with pytest.raises(ValueError):  # r = lo32(x) + top3(x) everywhere is not exact for 64-bit hashes
        dsk.engine.bulk_signatures(g["u64_tokens"], g["u64_offsets"], P, kernel="direct")


def test_sha1_update_batch_like_reference_gpu_test(dsk, golden):
    # shapes of test/test_minhash_gpu.py:26-52 of the reference
    g = golden("minhash")
    data = [f"token-{i}".encode() for i in range(1000)]
    m = dsk.MinHash(num_perm=256, seed=7, gpu_mode="always")
    m.update_batch(data)
    assert np.array_equal(m.hashvalues, g["sha1_k256_s7_n1000"])
    m = dsk.MinHash(num_perm=128, seed=7, gpu_mode="detect")
    m.update_batch(data[:500])
    m.update_batch(data[:700])
    assert np.array_equal(m.hashvalues, g["sha1_k128_s7_500_700"])


def test_update_equals_update_batch_equals_bulk(dsk):
    ident = int
    m1 = dsk.MinHash(4, 1, hashfunc=ident)
    m1.update(12)
    m1.update(24)
    m2 = dsk.MinHash(4, 1, hashfunc=ident)
    m2.update_batch([12, 24])
    assert all(m1.hashvalues == m2.hashvalues)
    kwargs = dict(num_perm=4, seed=1, hashfunc=ident)
    b = [[n * 4 for n in range(4)]] * 2
    m = dsk.MinHash(**kwargs)
    m.update_batch(b[0])
    x, y = dsk.MinHash.bulk(b, **kwargs)
    assert np.array_equal(m.hashvalues, x.hashvalues) and np.array_equal(m.hashvalues, y.hashvalues)
    want = o.update_batch(o.init_hashvalues(4), b[0], o.init_permutations(4, 1))
    assert np.array_equal(m.hashvalues, want)


def test_running_state_merge_and_estimators(dsk, golden):
    g = golden("minhash")
    m1 = dsk.MinHash(num_perm=128, seed=1, hashfunc=int)
    m1.update_batch(range(0, 100))
    _ = m1.hashvalues  # flush, then continue on a non-empty state (minhash.py:297)
    m1.update_batch(range(100, 300))
    m2 = dsk.MinHash(num_perm=128, seed=1, hashfunc=int)
    m2.update_batch(range(150, 450))
    assert np.array_equal(m1.hashvalues, g["j_m1"]) and np.array_equal(m2.hashvalues, g["j_m2"])
    assert m1.jaccard(m2) == float(g["j_jaccard"])
    assert m1.count() == float(g["j_count1"])
    assert np.array_equal(dsk.MinHash.union(m1, m2).hashvalues, g["j_union"])
    mm = m1.copy()
    mm.merge(m2)
    assert np.array_equal(mm.hashvalues, g["j_merge"])
    p = pickle.loads(pickle.dumps(m1))
    assert p.seed == m1.seed and np.array_equal(p.hashvalues, m1.hashvalues)
    assert np.array_equal(p.permutations, m1.permutations)


def test_init_matrix_and_broadcast(dsk):
    rs = np.random.RandomState(3)
    P = o.init_permutations(128, 9)
    tok = rs.randint(0, 2 ** 32, size=50 * 40, dtype=np.uint64).astype(np.uint32)
    off = np.arange(51, dtype=np.int64) * 40
    base = dsk.engine.bulk_signatures(tok, off, P)
    tok2 = rs.randint(0, 2 ** 32, size=50 * 40, dtype=np.uint64).astype(np.uint32)
    both = dsk.engine.bulk_signatures(tok2, off, P, init=base)
    want = np.minimum(base, oc.minhash_bulk_u32tok(tok2, off, P))
    assert np.array_equal(both, want)
    row = base[7].astype(np.uint64)
    bc = dsk.engine.bulk_signatures(tok2, off, P, init=row, out_u64=True)
    assert np.array_equal(bc, np.minimum(row[None, :], oc.minhash_bulk_u32tok(tok2, off, P).astype(np.uint64)))


@pytest.mark.parametrize("kernel", KERNELS)
def test_tail_and_alignment_shapes(dsk, kernel):
    # total token counts that are not multiples of 4 / 16 exercise the non-TMA tail path
    rs = np.random.RandomState(5)
    P = o.init_permutations(128, 1)
    for n_tok in [1, 2, 3, 5, 15, 17, 31, 33, 511, 513, 1023, 2049]:
        tok = rs.randint(0, 2 ** 32, size=n_tok, dtype=np.uint64).astype(np.uint32)
        cuts = np.sort(rs.randint(0, n_tok + 1, size=5))
        off = np.concatenate([[0], cuts, [n_tok]]).astype(np.int64)
        sig = dsk.engine.bulk_signatures(tok, off, P, kernel=kernel)
        assert np.array_equal(sig, oc.minhash_bulk_u32tok(tok, off, P)), n_tok


def test_duplicates_and_near_ties_take_slow_path(dsk):
    # small multipliers make many L' values collide / land within the +7 window and < 7
    rs = np.random.RandomState(8)
    a = rs.randint(1, 64, size=128).astype(np.uint64)
    a[::4] |= rs.randint(0, 2 ** 29, size=32).astype(np.uint64) << np.uint64(32)
    b = rs.randint(0, 200, size=128).astype(np.uint64)
    b[1::2] = rs.randint(0, 2 ** 61 - 1, size=64, dtype=np.uint64)
    P = np.stack([a, b])
    from datasketch_b200 import _native as nv
    docs = []
    for i in range(200):
        t = rs.randint(0, 64 if i % 2 else 2 ** 32, size=rs.randint(1, 150), dtype=np.uint64)
        if i % 3 == 0:
            t[rs.randint(0, len(t), size=len(t) // 2)] = t[0]
        docs.append(t.astype(np.uint32))
    off = np.zeros(len(docs) + 1, dtype=np.int64)
    np.cumsum([len(d) for d in docs], out=off[1:])
    tok = np.concatenate(docs)
    want = oc.minhash_bulk_u32tok(tok, off, P)
    h = nv.perm_handle(P)
    kernels = ["auto", "exact", "two_phase"] + (["direct"] if h.n_unsafe == 0 else [])
    for kernel in kernels:
        assert np.array_equal(dsk.engine.bulk_signatures(tok, off, P, kernel=kernel), want), kernel


def test_unsafe_permutations_take_the_general_variant(dsk):
    # x = a*h + b can hit the 36 values where `% (2^61-1)` subtracts: r = lo32(x) + top3(x) would be off by one there, so
    # the handle must flag it; AUTO / TWO_PHASE then run the signature kernel's general variant (conditional subtract in
    # its exact stage, window 8) and stay bit-exact, DIRECT refuses.
    from datasketch_b200 import _native as nv
    p = (1 << 61) - 1
    a = np.array([1, 1, 3, 0, 1 << 40, 5], dtype=np.uint64)
    b = np.array([p - 5, (7 << 61) + p - 9, p - 30, p, 12345, (3 << 61) + p - 2 - 5 * 77], dtype=np.uint64)
    P = np.stack([a, b])
    h = nv.perm_handle(P)
    assert h.n_unsafe >= 4
    tok = np.array([0, 1, 2, 5, 9, 10, 77, 1 << 24, 2 ** 32 - 1], dtype=np.uint32)
    off = np.array([0, 3, 3, 9], dtype=np.int64)
    want = o.bulk_signatures_csr(tok, off, 6, 0, permutations=P).astype(np.uint32)
    assert np.array_equal(dsk.engine.bulk_signatures(tok, off, P), want)
    assert np.array_equal(dsk.engine.bulk_signatures(tok, off, P, kernel="exact"), want)
    assert np.array_equal(dsk.engine.bulk_signatures(tok, off, P, kernel="two_phase"), want)
    with pytest.raises(ValueError):
        dsk.engine.bulk_signatures(tok, off, P, kernel="direct")
    # the values the fast formula would get wrong are really exercised:
    hits = [(int(ai) * int(t) + int(bi)) % (1 << 64) for ai, bi in zip(a, b) for t in tok]
    assert any(((x & p) + (x >> 61)) >= p for x in hits)


def test_seeded_permutations_are_safe(dsk):
    from datasketch_b200 import _native as nv
    for k, seed in [(128, 1), (256, 7), (512, 123)]:
        assert nv.perm_handle(o.init_permutations(k, seed)).n_unsafe == 0


def test_device_buffer_api(dsk):
    import torch
    rs = np.random.RandomState(12)
    P = o.init_permutations(128, 1)
    tok = rs.randint(0, 2 ** 32, size=300 * 77, dtype=np.uint64).astype(np.uint32)
    off = np.arange(301, dtype=np.int64) * 77
    d_tok = torch.from_numpy(tok.view(np.int32)).cuda()
    d_off = torch.from_numpy(off).cuda()
    d_sig = dsk.engine.bulk_signatures_device(d_tok, d_off, tok.size, P)
    torch.cuda.synchronize()
    got = d_sig.cpu().numpy().view(np.uint32)
    assert np.array_equal(got, oc.minhash_bulk_u32tok(tok, off, P))
    # misaligned token pointer is rejected, not silently mis-read
    with pytest.raises(ValueError):
        dsk.engine.bulk_signatures_device(d_tok[1:], d_off, tok.size - 1, P)


@pytest.mark.parametrize("kernel", ["two_phase", "direct"])
def test_medium_batch_sampled_against_oracle(dsk, kernel):
    # 50k docs x 256 tokens (C2 shape, scaled): every 97th document checked against the C oracle
    rs = np.random.RandomState(0)
    n, t = 50_000, 256
    tok = rs.randint(0, 2 ** 32, size=n * t, dtype=np.uint64).astype(np.uint32)
    off = np.arange(n + 1, dtype=np.int64) * t
    P = o.init_permutations(128, 1)
    sig = dsk.engine.bulk_signatures(tok, off, P, kernel=kernel)
    idx = np.arange(0, n, 97)
    sub_off = np.arange(len(idx) + 1, dtype=np.int64) * t
    sub_tok = tok.reshape(n, t)[idx].reshape(-1)
    assert np.array_equal(sig[idx], oc.minhash_bulk_u32tok(sub_tok, sub_off, P))
    # size-independent properties: permuting tokens inside a document, or duplicating them,
    # does not change its signature; splitting a document and min-merging the halves does not either
    perm_tok = tok.reshape(n, t)[:, ::-1].reshape(-1).copy()
    assert np.array_equal(dsk.engine.bulk_signatures(perm_tok, off, P, kernel=kernel), sig)
    half = np.arange(2 * n + 1, dtype=np.int64) * (t // 2)
    parts = dsk.engine.bulk_signatures(tok, half, P, kernel=kernel)
    assert np.array_equal(np.minimum(parts[0::2], parts[1::2]), sig)


@pytest.mark.parametrize("kernel", ["auto", "exact"])
def test_long_documents_are_split_and_merged(dsk, kernel):
    # host pipeline cuts documents above the threshold into pieces (many warps) and min-merges on device
    rs = np.random.RandomState(21)
    P = o.init_permutations(128, 1)
    lens = [5000, 0, 17, 4096, 4097, 70_000, 3, 0, 2_500_000, 129]
    off = np.zeros(len(lens) + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    tok = rs.randint(0, 2 ** 32, size=int(off[-1]), dtype=np.uint64).astype(np.uint32)
    want = oc.minhash_bulk_u32tok(tok, off, P)
    got = dsk.engine.bulk_signatures(tok, off, P, kernel=kernel)
    assert np.array_equal(got, want)
    init = rs.randint(0, 2 ** 32, size=(len(lens), 128), dtype=np.uint64)
    got64 = dsk.engine.bulk_signatures(tok, off, P, kernel=kernel, init=init, out_u64=True)
    assert got64.dtype == np.uint64 and np.array_equal(got64, np.minimum(init, want.astype(np.uint64)))
    # the reference's own GPU benchmark shape: one update_batch of 50 000 tokens (docs/minhash.rst:161-169)
    m = dsk.MinHash(num_perm=128, seed=1, hashfunc=int)
    m.update_batch(int(x) for x in tok[:50_000])
    assert np.array_equal(m.hashvalues, oc.minhash_bulk_u32tok(tok[:50_000], np.array([0, 50_000]), P)[0].astype(np.uint64))


def test_long_documents_on_the_device_buffer_entry(dsk):
    """dsk_minhash_bulk_ws through engine.bulk_signatures_device: long documents are cut into pieces ON THE DEVICE
    (signature_kernel.cu piece mode, caller-owned workspace) -- one 2.5 M-token document, the reference's 50 000-token
    benchmark shape, repeats inside a long document, long documents at the batch's ends; u64 output + running state."""
    import torch
    rs = np.random.RandomState(33)
    lens = [60_000, 5000, 0, 17, 4096, 4097, 50_000, 3, 2_500_000, 129, 20_000]
    off = np.zeros(len(lens) + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    tok = rs.randint(0, 2 ** 32, size=int(off[-1]), dtype=np.uint64).astype(np.uint32)
    tok[off[6]:off[6] + 25_000] = tok[off[6] + 25_000:off[7]]          # a long document made of repeats
    for k in (128, 256, 300):
        P = o.init_permutations(k, 1)
        want = oc.minhash_bulk_u32tok(tok, off, P)
        d_tok = torch.from_numpy(tok.view(np.int32)).cuda()
        d_off = torch.from_numpy(off).cuda()
        got = dsk.engine.bulk_signatures_device(d_tok, d_off, len(tok), P).cpu().numpy().view(np.uint32)
        assert np.array_equal(got, want), k
    init = rs.randint(0, 2 ** 32, size=(len(lens), 300), dtype=np.uint64)
    d_init = torch.from_numpy(init.view(np.int64)).cuda()
    d_out = torch.empty((len(lens), 300), dtype=torch.int64, device="cuda")
    dsk.engine.bulk_signatures_device(d_tok, d_off, len(tok), P, d_out=d_out, d_init=d_init, init_stride=300)
    assert np.array_equal(d_out.cpu().numpy().view(np.uint64), np.minimum(init, want.astype(np.uint64)))
    # timing contract: one 2.5 M-token document within 1.5x of the same tokens as 1000 documents
    one = torch.from_numpy(tok[off[8]:off[9]].view(np.int32).copy()).cuda()
    P = o.init_permutations(128, 1)
    off1 = torch.tensor([0, 2_500_000], dtype=torch.int64, device="cuda")
    offk = torch.arange(0, 2_500_001, 2500, dtype=torch.int64, device="cuda")

    def ms(o_):
        for _ in range(2):
            dsk.engine.bulk_signatures_device(one, o_, 2_500_000, P)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            dsk.engine.bulk_signatures_device(one, o_, 2_500_000, P)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / 5
    t_one, t_many = ms(off1), ms(offk)
    assert t_one < 1.5 * t_many + 0.05, (t_one, t_many)


def test_full_size_c2_properties(dsk):
    """BASELINE.json configs[1] at full size (1M docs x 256 tokens, K=128): size-independent properties
    plus a sampled comparison with the C oracle."""
    n, t, k = 1_000_000, 256, 128
    rng = np.random.default_rng(1234)
    tok = np.empty(n * t, dtype=np.uint32)
    step = 1 << 24
    for i in range(0, tok.size, step):
        tok[i:i + step] = rng.integers(0, 1 << 32, size=min(step, tok.size - i), dtype=np.uint32)
    off = np.arange(n + 1, dtype=np.int64) * t
    P = o.init_permutations(k, 1)
    sig = dsk.engine.bulk_signatures(tok, off, P)
    assert sig.shape == (n, k) and sig.dtype == np.uint32
    # (1) sampled rows equal the oracle
    idx = np.arange(0, n, 9973)
    sub = np.ascontiguousarray(tok.reshape(n, t)[idx]).reshape(-1)
    assert np.array_equal(sig[idx], oc.minhash_bulk_u32tok(sub, np.arange(len(idx) + 1, dtype=np.int64) * t, P))
    # (2) splitting every document in two and min-merging the halves changes nothing (merge rule, minhash.py:359)
    half = np.arange(2 * n + 1, dtype=np.int64) * (t // 2)
    parts = dsk.engine.bulk_signatures(tok, half, P)
    assert np.array_equal(np.minimum(parts[0::2], parts[1::2]), sig)
    del parts
    # (3) the direct kernel (different arithmetic path) produces the identical matrix: checksum of checksums
    sig_d = dsk.engine.bulk_signatures(tok, off, P, kernel="direct")
    assert int(sig_d.astype(np.uint64).sum()) == int(sig.astype(np.uint64).sum()) and np.array_equal(sig_d, sig)
    # (4) no document of 256 uniform tokens keeps the empty value; minima are small as order statistics predict
    assert (sig != 0xFFFFFFFF).all() and float(sig.mean()) < 2 ** 32 / 200


def test_api_fuzz_against_oracle(dsk):
    """Random sequences of update / update_batch / merge / copy / clear / union / LeanMinHash on MinHash objects
    (lazy queue + GPU flush) tracked step by step against the numpy oracle."""
    rs = np.random.RandomState(77)
    for k, seed in [(128, 1), (33, 5), (256, 9)]:
        P = o.init_permutations(k, seed)
        objs = [dsk.MinHash(num_perm=k, seed=seed, hashfunc=int) for _ in range(4)]
        refs = [o.init_hashvalues(k) for _ in range(4)]
        for step in range(120):
            i, j = rs.randint(0, 4), rs.randint(0, 4)
            op = rs.randint(0, 8)
            if op == 0:
                v = int(rs.randint(0, 2 ** 32, dtype=np.uint64))
                objs[i].update(v)
                refs[i] = o.update_one(refs[i], v, P)
            elif op in (1, 2):
                vals = [int(x) for x in rs.randint(0, 2 ** 32 if op == 1 else 50, size=rs.randint(0, 40), dtype=np.uint64)]
                objs[i].update_batch(vals)
                refs[i] = o.update_batch(refs[i], vals, P)
            elif op == 3 and i != j:
                objs[i].merge(objs[j])
                refs[i] = o.merge(refs[i], refs[j])
            elif op == 4:
                objs[i] = objs[j].copy()
                refs[i] = refs[j].copy()
            elif op == 5 and rs.rand() < 0.3:
                objs[i].clear()
                refs[i] = o.init_hashvalues(k)
            elif op == 6:
                u = dsk.MinHash.union(objs[i], objs[j])
                assert np.array_equal(u.hashvalues, np.minimum(refs[i], refs[j]))
                assert dsk.LeanMinHash(u).jaccard(dsk.LeanMinHash(objs[i])) == o.jaccard(np.minimum(refs[i], refs[j]), refs[i])
            else:
                assert objs[i].jaccard(objs[j]) == o.jaccard(refs[i], refs[j])
                assert objs[i].is_empty() == bool((refs[i] == 0xFFFFFFFF).all())
            if step % 7 == 0:
                assert np.array_equal(objs[i].hashvalues, refs[i]) and objs[i].hashvalues.dtype == np.uint64
        for a, b in zip(objs, refs):
            assert np.array_equal(a.hashvalues, b) and len(a) == k
