"""Round-2 starting point: validate and time the experimental re-scan variant (DSK_RESCAN=1) of the two-phase
kernel on documents with repeated tokens.  NOT part of tests/: the variant was written after the round-1 GPU budget
was spent and has only been compiled, never run (its SASS is separate; the default kernels are byte-identical).

    gpurun -- 'python tools/check_rescan.py && python tools/bench_duplicates.py; DSK_RESCAN=1 python tools/bench_duplicates.py'
"""
import os
import sys

os.environ["DSK_RESCAN"] = "1"   # must be set before the first launch (read once)

import numpy as np  # noqa: E402

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import datasketch_b200 as dsk  # noqa: E402
from datasketch_b200.minhash import _make_permutations  # noqa: E402
from oracle import oracle_clib as oc  # noqa: E402  (checker only)

rs = np.random.RandomState(5)
bad = 0
for k in (128, 256, 100, 64, 16):
    perms = _make_permutations(k, 1)
    for share in (0.0, 0.05, 0.5, 0.95):
        n = 3000
        lens = rs.randint(0, 700, size=n)
        lens[:4] = [0, 1, 15, 17]
        off = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(lens, out=off[1:])
        tok = rs.randint(0, 2 ** 32, size=int(off[-1]), dtype=np.uint64).astype(np.uint32)
        for d in range(n):                       # repeats of earlier tokens of the same document
            a, b = off[d], off[d + 1]
            if b - a > 1:
                rep = np.nonzero(rs.uniform(size=b - a) < share)[0]
                rep = rep[rep > 0]
                tok[a + rep] = tok[a + (rs.uniform(size=len(rep)) * rep).astype(np.int64)]
        got = dsk.engine.bulk_signatures(tok, off, perms)
        want = oc.minhash_bulk_u32tok(tok, off, perms)
        ok = np.array_equal(got, want)
        bad += not ok
        print(f"num_perm={k:4d} repeat_share={share:4.2f} docs={n} identical={ok}", flush=True)
# small values / near-wrap tokens exercise the m < 7 rule
tok = np.concatenate([np.arange(0, 4096, dtype=np.uint32), np.arange(2 ** 32 - 4096, 2 ** 32, dtype=np.uint64).astype(np.uint32)])
off = np.arange(0, len(tok) + 1, 64, dtype=np.int64)
perms = _make_permutations(128, 3)
ok = np.array_equal(dsk.engine.bulk_signatures(tok, off, perms), oc.minhash_bulk_u32tok(tok, off, perms))
bad += not ok
print("structured tokens identical =", ok)
sys.exit(1 if bad else 0)
