"""GPU tests written AFTER the round-1 GPU budget was spent: their CUDA paths have passed the CPU emulation suite
(tests/test_kernel_emulation_cpu.py) and the host logic has passed on CPU, but they have not yet run on a B200.  The file
name sorts last so that, under ``pytest -x``, everything validated on hardware runs first.  Move the tests next to their
siblings (test_sha1_gpu.py, test_lshensemble.py) once they have passed on a GPU."""
import numpy as np
import pytest

from oracle import oracle_np as o

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dsk():
    import datasketch_b200
    return datasketch_b200


def test_ensemble_index_query_match_reference_gpu(dsk, golden):
    from test_lshensemble import index_and_check
    index_and_check(dsk, golden("ensemble"))


def test_xxh32_murmur3_device_hashes_and_bulk(dsk, golden):
    """dsk_hash_tokens against the values the `xxhash` package produced (fixtures) and the MurmurHash3 oracle, and
    MinHash.bulk with those hash functions (device route) against per-document update_batch (host hash route)."""
    g = golden("hashes")
    blob, off = g["blob"], g["off"]
    toks = [bytes(blob[off[i]:off[i + 1]]) for i in range(len(off) - 1)]
    for seed in (0, 1, 0x9747B28C):
        got = dsk.engine.hash_tokens_device(toks, 1, seed).cpu().numpy().view(np.uint32)
        assert np.array_equal(got, g[f"xxh32_seed{seed}"])
        got = dsk.engine.hash_tokens_device(toks, 2, seed).cpu().numpy().view(np.uint32)
        assert got.tolist() == [o.murmur3_32(t, seed) for t in toks]
    docs = [toks[:50], [], toks[50:53], toks[100:400]]
    for hf in (dsk.xxh32_hash32, dsk.murmur3_hash32):
        ms = dsk.MinHash.bulk(docs, num_perm=64, seed=3, hashfunc=hf)
        for d, m in zip(docs, ms):
            ref = dsk.MinHash(num_perm=64, seed=3, hashfunc=hf)
            ref.update_batch(d)
            assert np.array_equal(ref.hashvalues, m.hashvalues)
            want = o.update_batch(o.init_hashvalues(64), [hf(t) for t in d], o.init_permutations(64, 3))
            assert np.array_equal(m.hashvalues, want)
    with pytest.raises(ValueError):
        dsk.engine.hash_tokens_device(toks[:3], 7)
