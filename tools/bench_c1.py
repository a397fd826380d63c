"""Config C1 through the drop-in API: ``MinHash.bulk`` over byte-token documents with the default
hash function (SHA1-32), i.e. the call a datasketch user makes (datasketch/minhash.py:464-489).

Timed: datasketch_b200.MinHash.bulk (device SHA1 + signature kernel, host packing included) against
the CPU restatement of the reference loop (hashlib SHA1 per token + numpy update_batch per document,
oracle/oracle_np.py) on a bounded sample.  Every compared row must be identical.  Prints JSON lines.
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def make_docs(n_docs, n_tok, seed=7):
    rng = np.random.default_rng(seed)
    ids = rng.integers(0, 1 << 40, size=(n_docs, n_tok))
    return [[b"shingle-%d" % v for v in row] for row in ids]


def main():
    import torch
    from datasketch_b200 import MinHash
    from oracle import oracle_np

    out = []
    for n_docs, n_tok in ((1000, 64), (10_000, 256), (100_000, 256)):
        docs = make_docs(n_docs, n_tok)
        MinHash.bulk(docs[:64], num_perm=128)  # warm-up: library load, permutation upload
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        mhs = MinHash.bulk(docs, num_perm=128)
        t_gpu = time.perf_counter() - t0

        sample = min(n_docs, 300)
        perms = np.asarray(mhs[0].permutations, dtype=np.uint64)
        t0 = time.perf_counter()
        ref_rows = []
        for d in docs[:sample]:
            hv = np.array([oracle_np.sha1_hash32(t) for t in d], dtype=np.uint64)
            ref_rows.append(oracle_np.update_batch(oracle_np.init_hashvalues(128), hv, perms))
        t_cpu = (time.perf_counter() - t0) * n_docs / sample
        same = all(np.array_equal(mhs[i].hashvalues, ref_rows[i]) for i in range(sample))
        out.append({"config": "C1-style MinHash.bulk, bytes tokens, sha1_hash32", "docs": n_docs, "tokens": n_tok,
                    "num_perm": 128, "bulk_s": round(t_gpu, 4), "docs_per_s": round(n_docs / t_gpu, 1),
                    "cpu_reference_loop_s_extrapolated": round(t_cpu, 2), "cpu_sample_docs": sample,
                    "speedup_vs_one_core": round(t_cpu / t_gpu, 1), "rows_identical": bool(same)})
        print(json.dumps(out[-1]), flush=True)
    assert all(o["rows_identical"] for o in out)


if __name__ == "__main__":
    main()
