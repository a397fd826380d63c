"""Codec kernels vs the HBM roofline (algorithmic bytes / CUDA-event time).  GPU only."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import datasketch_b200 as dsk  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
peak = 6650.0
pp = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(pp):
    peak = float(json.load(open(pp))["hbm_gbs"])


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


out = []
for n, k, b, r in [(2_000_000, 128, 9, 13), (1_000_000, 256, 17, 15)]:
    sig = torch.randint(-2 ** 31, 2 ** 31 - 1, (n, k), dtype=torch.int32, device="cuda")
    rec = torch.empty((n, 12 + 4 * k), dtype=torch.uint8, device="cuda")
    ms = timeit(lambda: dsk.codec.lean_pack(sig, 1, out=rec))
    by = n * (4 * k + 12 + 4 * k)
    out.append({"kernel": "lean_pack", "n": n, "k": k, "ms": ms, "GBps": by / ms / 1e6, "frac": by / ms / 1e6 / peak})
    ms = timeit(lambda: dsk.codec.lean_unpack(rec, k, 1))
    out.append({"kernel": "lean_unpack(+status sync)", "n": n, "k": k, "ms": ms, "GBps": by / ms / 1e6, "frac": by / ms / 1e6 / peak})
    ms = timeit(lambda: dsk.codec.band_keys(sig, b, r))
    by = n * (4 * b * r + 8 * b * r)
    out.append({"kernel": "band_keys", "n": n, "k": k, "b": b, "r": r, "ms": ms, "GBps": by / ms / 1e6, "frac": by / ms / 1e6 / peak})
    ms = timeit(lambda: dsk.codec.band_fingerprints(sig, b, r))
    by = n * (4 * b * r + 8 * b)
    out.append({"kernel": "band_fingerprints", "n": n, "k": k, "b": b, "r": r, "ms": ms, "GBps": by / ms / 1e6, "frac": by / ms / 1e6 / peak})
    ms = timeit(lambda: dsk.codec.band_sums(sig, b, r))
    by = n * (4 * b * r + 8 * b)
    out.append({"kernel": "band_sums (LSHBloom keys)", "n": n, "k": k, "b": b, "r": r, "ms": ms, "GBps": by / ms / 1e6, "frac": by / ms / 1e6 / peak})
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        bl = dsk.MinHashLSHBloom(num_perm=k, n=n, fp=0.001, params=(b, r))
    ms = timeit(lambda: bl.insert_batch(sig), iters=3, warm=1)
    out.append({"kernel": "bloom insert (keys + %d probes x %d bands)" % (bl.n_hashes, b), "n": n, "k": k, "ms": ms, "docs_per_s": n / ms * 1e3,
                "table_MB": bl.b * bl.words_per_table * 4 / 1e6})
    ms = timeit(lambda: bl.query_batch(sig, to_host=False), iters=3, warm=1)
    out.append({"kernel": "bloom query", "n": n, "k": k, "ms": ms, "docs_per_s": n / ms * 1e3})
    del sig, rec, bl
for o in out:
    print(json.dumps(o))
