"""b-bit MinHash ("next" row of SURVEY.md 8f): host class vs fixtures produced by the reference (CPU),
batch pack / unpack kernels vs the host class (GPU)."""
import pickle

import numpy as np
import pytest


@pytest.fixture(scope="module")
def dsk():
    import datasketch_b200
    return datasketch_b200


def _pair(dsk, g):
    return (dsk.LeanMinHash(seed=1, hashvalues=g["hv1"]), dsk.LeanMinHash(seed=1, hashvalues=g["hv2"]))


def test_bbit_host_matches_reference(dsk, golden):
    g = golden("bbit")
    m1, m2 = _pair(dsk, g)
    for b in [int(x) for x in g["bs"]]:
        x, y = dsk.bBitMinHash(m1, b), dsk.bBitMinHash(m2, b)
        assert x.hashvalues.dtype == np.uint32 and np.array_equal(x.hashvalues, g[f"hv_b{b}"])
        assert bytes(x.__getstate__()) == g[f"state_b{b}"].tobytes()
        assert x.bytesize() == int(g[f"size_b{b}"])
        assert x.jaccard(y) == float(g[f"jac_b{b}"])
        z = pickle.loads(pickle.dumps(x))
        assert z == x and z.b == b and z.seed == 1
    assert dsk.bBitMinHash(m1, 4, r=0.3).jaccard(dsk.bBitMinHash(m2, 4, r=0.1)) == float(g["jac_b4_r"])
    with pytest.raises(ValueError):
        dsk.bBitMinHash(m1, 33)
    with pytest.raises(ValueError):
        dsk.bBitMinHash(m1, 4, r=1.5)
    with pytest.raises(ValueError):
        dsk.bBitMinHash(m1, 4).jaccard(dsk.bBitMinHash(m1, 5))
    # test/test_minhash.py:148-175 of the reference
    e1, e2 = dsk.bBitMinHash(dsk.MinHash(4, 1)), dsk.bBitMinHash(dsk.MinHash(4, 1))
    assert e1 == e2 and e1.jaccard(e2) == 1.0 and e1 != dsk.bBitMinHash(dsk.MinHash(8, 1))


@pytest.mark.gpu
@pytest.mark.parametrize("b", [1, 2, 3, 4, 5, 8, 12, 16, 27, 32])
def test_bbit_pack_kernel_matches_host_class(dsk, golden, b):
    g = golden("bbit")
    rs = np.random.RandomState(b)
    for k in (128, 100, 7):
        sig = rs.randint(0, 2 ** 32, size=(57, k), dtype=np.uint64).astype(np.uint32)
        if k == 128:
            sig[0] = g["hv1"].astype(np.uint32)
        blocks = dsk.codec.bbit_pack(sig, b).cpu().numpy().view(np.uint64)
        for i in (0, 13, 56):
            x = dsk.bBitMinHash(dsk.LeanMinHash(seed=1, hashvalues=sig[i].astype(np.uint64)), b)
            state = bytes(x.__getstate__())
            assert blocks[i].tobytes() == state[21:]            # payload after the '<qBdi' header
        if k == 128:
            assert blocks[0].tobytes() == g[f"state_b{b}"].tobytes()[21:]
        back = dsk.codec.bbit_unpack(blocks, k, b).cpu().numpy().view(np.uint32)
        slot = 1 if b == 1 else 2 if b == 2 else 4 if b <= 4 else 8 if b <= 8 else 16 if b <= 16 else 32
        mask = np.uint32((1 << b) - 1 if b < 32 else 0xFFFFFFFF)
        assert np.array_equal(back, sig & mask) and slot >= b
