"""MinHashLSHEnsemble against fixtures produced by the reference (oracle/gen_golden.py:gen_ensemble):
partition bounds, (b, r) tables, and the query results of an indexed corpus.

The CPU tests inject the oracle's band keys into ``MinHashLSH._batch_band_keys`` (this container has no GPU);
the GPU run of the same check (product path: one ``dsk_band_keys`` launch per partition and r) lives in
tests/test_zz_first_gpu_run.py."""
import numpy as np
import pytest

from oracle import oracle_np as o


@pytest.fixture(scope="module")
def dsk():
    import datasketch_b200
    return datasketch_b200


def test_partitions_match_reference(dsk, golden):
    from datasketch_b200.lshensemble import optimal_partitions
    g = golden("ensemble")
    for i in range(int(g["n_part_cases"][0])):
        got = optimal_partitions(g[f"part{i}_sizes"], g[f"part{i}_counts"], int(g[f"part{i}_num_part"][0]))
        assert np.array_equal(np.array(got, dtype=np.int64), g[f"part{i}_bounds"]), i


def test_param_tables_match_reference(dsk, golden):
    g = golden("ensemble")
    for i in range(3):
        thr, k, m, w0, w1 = g[f"par{i}_cfg"]
        e = dsk.MinHashLSHEnsemble(threshold=float(thr), num_perm=int(k), num_part=2, m=int(m), weights=(float(w0), float(w1)))
        assert np.array_equal(e.params, g[f"par{i}_params"]) and np.allclose(e.xqs, g[f"par{i}_xqs"], rtol=0, atol=0)
        assert set(e.indexes[0]) == {int(r) for _, r in e.params} and len(e.indexes) == 2
        for r, lsh in e.indexes[1].items():
            assert (lsh.b, lsh.r, lsh.h) == (int(k) // r, r, int(k))


def test_constructor_and_index_errors(dsk):
    E = dsk.MinHashLSHEnsemble
    for kw in [dict(threshold=1.5), dict(threshold=-0.1), dict(num_perm=1), dict(num_part=0), dict(m=1),
               dict(m=200, num_perm=128), dict(weights=(1.5, -0.5)), dict(weights=(0.3, 0.3))]:
        with pytest.raises(ValueError):
            E(**kw)
    e = E(threshold=0.5, num_perm=16, num_part=2, m=2)
    assert e.is_empty() and "x" not in e
    with pytest.raises(ValueError):
        e.index([])
    m = dsk.MinHash(num_perm=16, hashvalues=np.arange(16))
    with pytest.raises(ValueError):
        e.index(iter([("a", m, 0)]))                 # "Set size must be positive" (checked for non-list input)


def index_and_check(dsk, g):
    thr, k, num_part, m = g["e2e_cfg"]
    sig, sizes = g["e2e_sig"], g["e2e_sizes"]
    mhs = [dsk.MinHash(num_perm=int(k), seed=1, hashvalues=row.astype(np.uint64)) for row in sig]
    ens = dsk.MinHashLSHEnsemble(threshold=float(thr), num_perm=int(k), num_part=int(num_part), m=int(m))
    ens.index((i, mhs[i], int(sizes[i])) for i in range(len(mhs)))        # generator input, like the reference's tests
    assert [-1 if x is None else int(x) for x in ens.lowers] == g["e2e_lowers"].tolist()
    assert [-1 if x is None else int(x) for x in ens.uppers] == g["e2e_uppers"].tolist()
    assert not ens.is_empty() and 0 in ens and len(mhs) - 1 in ens and len(mhs) not in ens
    ptr, idx = g["e2e_qptr"], g["e2e_qidx"]
    for i, mh in enumerate(mhs):
        assert sorted(set(ens.query(mh, int(sizes[i])))) == idx[ptr[i]:ptr[i + 1]].tolist(), i
    with pytest.raises(ValueError):
        ens.index([(0, mhs[0], 5)])                                          # callable once
    return ens


def test_index_query_match_reference_with_oracle_band_keys(dsk, golden, monkeypatch):
    def oracle_keys(self, sig):
        return [o.lsh_band_keys(row.astype(np.uint64), self.b, self.r) for row in sig]
    monkeypatch.setattr(dsk.MinHashLSH, "_batch_band_keys", oracle_keys)
    index_and_check(dsk, golden("ensemble"))

