"""Throughput of the signature kernel on documents with REPEATED tokens.

A repeated token makes two 16-token blocks tie on the cheap phase-1 value L', which sends that
permutation -- and with it the whole warp -- through the two-phase kernel's exact slow path for the
document ("Known limitations").  This tool measures the penalty as a function of the share
of a document's tokens that are repeats of earlier ones, and checks every configuration against the
C oracle on a sample.  Round-1 result: profiles/r1z_repeated_tokens_penalty.jsonl.

    gpurun -- 'python tools/bench_duplicates.py > gpurun_out/dups.jsonl'
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import datasketch_b200 as dsk  # noqa: E402
from datasketch_b200.minhash import _make_permutations  # noqa: E402
from oracle import oracle_clib as oc  # noqa: E402  (checker only)

n, t, k = 500_000, 256, 128
dev = torch.device("cuda", 0)
perms = _make_permutations(k, 1)
off = torch.arange(0, (n + 1) * t, t, dtype=torch.int64, device=dev)
g = torch.Generator(device="cuda").manual_seed(1)
for share in (0.0, 0.01, 0.1, 0.5):
    tok = torch.randint(-2 ** 31, 2 ** 31 - 1, (n, t), dtype=torch.int32, device=dev, generator=g)
    if share > 0:  # position i repeats a token from an earlier position of the same document
        rep = torch.rand((n, t), device=dev, generator=g) < share
        src = (torch.rand((n, t), device=dev, generator=g) * torch.arange(t, device=dev)).long().clamp_(min=0)
        rep[:, 0] = False
        tok = torch.where(rep, torch.gather(tok, 1, src), tok)
    sig = torch.empty((n, k), dtype=torch.int32, device=dev)
    flat = tok.view(-1)
    for _ in range(3):
        dsk.engine.bulk_signatures_device(flat, off, n * t, perms, d_out=sig)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        dsk.engine.bulk_signatures_device(flat, off, n * t, perms, d_out=sig)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    idx = np.linspace(0, n - 1, 200).astype(np.int64)
    sub = tok[torch.from_numpy(idx).to(dev)].cpu().numpy().view(np.uint32).reshape(-1)
    want = oc.minhash_bulk_u32tok(sub, np.arange(len(idx) + 1, dtype=np.int64) * t, perms)
    got = sig[torch.from_numpy(idx).to(dev)].cpu().numpy().view(np.uint32)
    print(json.dumps({"repeat_share": share, "ms": ms, "signatures_per_s": n / ms * 1e3,
                      "rows_identical": bool(np.array_equal(got, want))}), flush=True)
