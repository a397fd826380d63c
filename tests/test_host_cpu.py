"""CPU: the C-ABI library loads and exports every symbol include/dsk.h declares; host-side
logic (permutation generation + safety analysis, CSR packing, API validation) -- no GPU."""
import ctypes
import os
import re
import warnings

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def dsk():
    import datasketch_b200
    return datasketch_b200


def _has_gpu():
    from datasketch_b200 import _native as nv
    return nv.device_count() > 0


def test_library_exports_every_declared_symbol(dsk):
    from datasketch_b200 import _native as nv
    hdr = open(os.path.join(ROOT, "include", "dsk.h")).read()
    declared = set(re.findall(r"DSK_API[^;(]*?\b(dsk_\w+)\s*\(", hdr))
    assert len(declared) >= 10
    lib = ctypes.CDLL(nv.lib_path())
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == set(nv.SIGNATURES), declared ^ set(nv.SIGNATURES)
    assert nv.load().dsk_version() == 100


def test_permutation_generation_matches_reference(dsk, golden):
    from datasketch_b200.minhash import _make_permutations
    g = golden("minhash")
    for k, seed in [(4, 1), (128, 1), (256, 7), (100, 42)]:
        p = _make_permutations(k, seed)
        assert p.dtype == np.uint64 and p.flags.c_contiguous
        assert np.array_equal(p, g[f"perm_k{k}_s{seed}"])
    m = dsk.MinHash(num_perm=128, seed=1)
    assert np.array_equal(m.permutations, g["perm_k128_s1"])
    assert m.is_empty() and len(m) == 128 and m.hashvalues.dtype == np.uint64


def _py_unsafe(a, b):
    """Independent big-int statement of dsk_perm_analyze."""
    p = (1 << 61) - 1
    for j in range(8):
        for lo in range(p - j, p + 1):
            x = (j << 61) | lo
            t = (x - b) % (1 << 64)
            if a == 0:
                if t == 0:
                    return True
                continue
            e = (a & -a).bit_length() - 1
            if t % (1 << e):
                continue
            nb = 64 - e
            if nb < 32:
                return True
            h0 = ((t >> e) * pow(a >> e, -1, 1 << nb)) % (1 << nb)
            if h0 < (1 << 32):
                return True
    return False


def test_perm_analysis(dsk):
    from datasketch_b200 import _native as nv
    rs = np.random.RandomState(0)
    p = (1 << 61) - 1
    # crafted: pick a token h and a value x in the subtract set, solve for b
    a_list, b_list = [], []
    for _ in range(200):
        a = int(rs.randint(0, 2 ** 63, dtype=np.uint64)) * 2 + int(rs.randint(0, 2))
        if rs.randint(0, 4) == 0:
            a <<= int(rs.randint(1, 40))
            a %= 1 << 64
        h = int(rs.randint(0, 2 ** 32, dtype=np.uint64))
        j = int(rs.randint(0, 8))
        x = (j << 61) | (p - int(rs.randint(0, j + 1)))
        b = (x - a * h) % (1 << 64)
        a_list.append(a)
        b_list.append(b)
    P = np.array([a_list, b_list], dtype=np.uint64)
    assert nv.perm_analyze(P).all()
    # random permutations: agree with the independent implementation (and are almost surely safe)
    a = rs.randint(1, p, size=300, dtype=np.uint64)
    b = rs.randint(0, p, size=300, dtype=np.uint64)
    a[::7] <<= np.uint64(20)
    a[::11] = np.uint64(0)
    b[::11] = np.uint64(p)
    got = nv.perm_analyze(np.stack([a, b]))
    want = np.array([_py_unsafe(int(x), int(y)) for x, y in zip(a, b)])
    assert np.array_equal(got, want)
    assert got[::11].all() and not got[1::11].any()


def test_pack_docs_and_token_dtypes(dsk):
    tok, off = dsk.engine.pack_docs([[1, 2, 3], [], [2 ** 32 - 1]])
    assert tok.dtype == np.uint32 and off.tolist() == [0, 3, 3, 4]
    tok, off = dsk.engine.pack_docs([[1], [2 ** 40]])
    assert tok.dtype == np.uint64
    tok, off = dsk.engine.pack_docs([[], []])
    assert tok.size == 0 and off.tolist() == [0, 0, 0]
    with pytest.raises(OverflowError):
        dsk.engine.pack_docs([[-1]])


def test_constructor_validation_matches_reference(dsk):
    MinHash = dsk.MinHash
    with pytest.raises(ValueError):
        MinHash(hashfunc=123)
    with pytest.raises(ValueError):
        MinHash(num_perm=4, hashvalues=[1, 2, 3, 4], permutations=np.zeros((2, 5), dtype=np.uint64))
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        MinHash(num_perm=4, hashobj=object())
        assert any(issubclass(x.category, DeprecationWarning) for x in w)
    m = MinHash(hashvalues=[5, 6, 7])
    assert m.num_perm == 3 and m.hashvalues.tolist() == [5, 6, 7] and m.hashvalues.dtype == np.uint64
    m1, m2 = MinHash(4, 1), MinHash(4, 2)
    with pytest.raises(ValueError):
        m1.jaccard(m2)
    with pytest.raises(ValueError):
        m1.merge(MinHash(8, 1))
    with pytest.raises(ValueError):
        MinHash.union(m1)
    assert m1.jaccard(MinHash(4, 1)) == 1.0
    assert m1 == MinHash(4, 1) and m1 != m2
    m1.update_batch([])  # empty batch is a no-op even without a device (minhash.py:265-266)
    assert m1.is_empty()


def test_no_device_raises_runtimeerror_not_fallback(dsk):
    if _has_gpu():
        pytest.skip("GPU present")
    m = dsk.MinHash(num_perm=64, seed=1, gpu_mode="always")
    with pytest.raises(RuntimeError):  # test/test_minhash_gpu.py:73-79 of the reference
        m.update_batch([b"a", b"b"])
    m.update(b"a")
    with pytest.raises(RuntimeError):
        _ = m.hashvalues
    with pytest.raises(RuntimeError):
        dsk.engine.bulk_signatures(np.zeros(4, np.uint32), np.array([0, 4]), m.permutations)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "datasketch_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "oracle" not in src.replace("the oracle", ""), os.path.join(dp, f)


def test_bulk_signatures_validates_every_offset(dsk):
    """engine.bulk_signatures checks the ends of the offsets BEFORE the library indexes host memory with them:
    offsets[0] > 0 past the end, a negative start, a reversed range (interior offsets: the C pass that reads them,
    tests/test_capi_emulation_cpu.py::test_interior_offsets_are_validated_by_the_library)."""
    P = dsk.minhash._make_permutations(16, 1)
    tok = np.arange(10, dtype=np.uint32)
    for off in ([5, 13], [-1, 3], [0, 11], [8, 2]):
        with pytest.raises(ValueError):
            dsk.engine.bulk_signatures(tok, np.array(off, dtype=np.int64), P)
