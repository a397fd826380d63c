"""Signature-kernel throughput over document SHAPES (aligned / ragged / short / long / K=256) and over the share of
REPEATED tokens, each configuration checked against the C oracle on a sample of rows.  The kernel is whatever the
process environment selects (default: minhash_sig_kernel; DSK_TWO_PHASE_V1=1: the round-1 kernel; DSK_SIG_OCC=5), so
an A/B is two runs:

    gpurun -- 'python tools/bench_shapes.py > gpurun_out/shapes_new.jsonl; DSK_TWO_PHASE_V1=1 python tools/bench_shapes.py > gpurun_out/shapes_v1.jsonl'
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

dev = torch.device("cuda", 0)
g = torch.Generator(device="cuda").manual_seed(1)
only = set(sys.argv[1:])


def run(name, lens, k, repeat_share=0.0, iters=5, u64=False, unsafe=False, kernel="auto"):
    if only and name not in only:
        return
    lens = torch.as_tensor(lens, dtype=torch.int64, device=dev)
    n = int(lens.numel())
    off = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    off[1:] = torch.cumsum(lens, 0)
    nt = int(off[-1].item())
    tok = torch.randint(-2 ** 31, 2 ** 31 - 1, (nt + 4,), dtype=torch.int32, device=dev, generator=g)[:nt]
    perms = _make_permutations(k, 1)
    if unsafe:      # one user-supplied permutation that reaches the conditional subtract of `% (2^61-1)` (token 3 -> x = p):
        perms = perms.copy()    # the handle is flagged and the kernel's general u32 variant (GEN = 1) runs
        perms[0, 5], perms[1, 5] = 1, (1 << 61) - 1 - 3
    if u64:         # 64-bit hash values: a random high word on every token (GEN = 2)
        hi = torch.randint(0, 1 << 32, (nt,), dtype=torch.int64, device=dev, generator=g)
        tok = hi.mul_(1 << 32).add_(tok.to(torch.int64) & 0xFFFFFFFF)
    if repeat_share > 0:   # position i of a document repeats a token from an earlier position of the same document
        doc = torch.repeat_interleave(torch.arange(n, device=dev), lens)
        pos = torch.arange(nt, device=dev) - off[doc]
        rep = (torch.rand(nt, device=dev, generator=g) < repeat_share) & (pos > 0)
        src = off[doc] + (torch.rand(nt, device=dev, generator=g) * pos).long()
        tok = torch.where(rep, tok[src], tok)
    sig = torch.empty((n, k), dtype=torch.int32, device=dev)
    for _ in range(3):
        dsk.engine.bulk_signatures_device(tok, off, nt, perms, d_out=sig, kernel=kernel)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        dsk.engine.bulk_signatures_device(tok, off, nt, perms, d_out=sig, kernel=kernel)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    idx = np.unique(np.linspace(0, n - 1, 150).astype(np.int64))
    h_off = off.cpu().numpy()
    h_tok = tok.cpu().numpy().view(np.uint64 if u64 else np.uint32)
    sub_off = np.zeros(len(idx) + 1, dtype=np.int64)
    np.cumsum(h_off[idx + 1] - h_off[idx], out=sub_off[1:])
    sub = np.concatenate([h_tok[h_off[i]:h_off[i + 1]] for i in idx]) if len(idx) else np.zeros(0, h_tok.dtype)
    want = (oc.minhash_bulk_u64tok if u64 else oc.minhash_bulk_u32tok)(np.ascontiguousarray(sub), sub_off, perms)
    got = sig[torch.from_numpy(idx).to(dev)].cpu().numpy().view(np.uint32)
    evals = float(nt) * k
    print(json.dumps({"shape": name, "tokens_are": "u64" if u64 else "u32", "kernel": kernel, "docs": n, "tokens": nt, "num_perm": k, "repeat_share": repeat_share, "ms": round(ms, 4),
                      "signatures_per_s": n / ms * 1e3, "evaluations_per_s": evals / ms * 1e3,
                      "frac_of_imad_floor": evals / ms * 1e3 / (148 * 64 * 1.965e9),
                      "rows_identical": bool(np.array_equal(got, want))}), flush=True)


rs = np.random.RandomState(0)
run("c2_aligned_1Mx256", np.full(1_000_000, 256), 128)
run("ragged_1M_128to384", rs.randint(128, 385, size=1_000_000), 128)
run("short_aligned_4Mx64", np.full(4_000_000, 64), 128)
run("short_ragged_4M_16to112", rs.randint(16, 113, size=4_000_000), 128)
run("long_20k_x12800", np.full(20_000, 12_800), 128)
run("long_ragged_40k_2kto10k", rs.randint(2_000, 10_001, size=40_000), 128)
run("one_doc_2.5M_tokens", np.array([2_500_000]), 128, iters=10)
run("same_tokens_as_1000_docs", np.full(1000, 2500), 128, iters=10)
run("reference_bench_shape_1x50k", np.array([50_000]), 128, iters=20)
run("k256_2Mx128", np.full(2_000_000, 128), 256)
run("k64_2Mx256", np.full(2_000_000, 256), 64)
run("k192_2Mx128", np.full(2_000_000, 128), 192)
for share in (0.01, 0.1, 0.5):
    run("repeats_500kx256", np.full(500_000, 256), 128, repeat_share=share)
run("repeats_ragged_500k", rs.randint(128, 385, size=500_000), 128, repeat_share=0.1)
# general variants of the signature kernel next to round 1's EXACT kernel (what these inputs were routed to before)
run("u64_tokens_1Mx256", np.full(1_000_000, 256), 128, u64=True)
run("u64_tokens_1Mx256_exact_kernel", np.full(1_000_000, 256), 128, u64=True, kernel="exact", iters=3)
run("u64_tokens_ragged_1M_128to384", rs.randint(128, 385, size=1_000_000), 128, u64=True)
run("u64_tokens_repeats_500kx256", np.full(500_000, 256), 128, repeat_share=0.1, u64=True)
run("u64_tokens_long_20k_x12800", np.full(20_000, 12_800), 128, u64=True)
run("unsafe_permutation_1Mx256", np.full(1_000_000, 256), 128, unsafe=True)
run("unsafe_permutation_1Mx256_exact_kernel", np.full(1_000_000, 256), 128, unsafe=True, kernel="exact", iters=3)
