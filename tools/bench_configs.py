"""Secondary workloads of BASELINE.json (configs[2..4]) on ONE GPU, with the oracle's CPU numbers beside them.
These are not bench.py lines; they are evidence for the other rows of SURVEY.md section 8.

    python tools/bench_configs.py [--c3-docs 2000000] [--c4-vecs 20000] [--c5-rows 100000]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import datasketch_b200 as dsk  # noqa: E402
from datasketch_b200.minhash import _make_permutations  # noqa: E402
from oracle import oracle_np as o  # noqa: E402


def ev_time(fn, iters=3, warm=1):
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


def c3(n_docs, t=128, k=256, n_query=100_000, fused_leg=False):
    """10M x 128 tokens, K=256 + LSH(0.8) insert + query (here n_docs on one GPU)."""
    perms = _make_permutations(k, 1)
    g = torch.Generator(device="cuda").manual_seed(3)
    tok = torch.randint(-2 ** 31, 2 ** 31 - 1, (n_docs, t), dtype=torch.int32, device="cuda", generator=g)
    # plant near-duplicates: 10 % of documents copy another one with ~10 % tokens resampled
    ndup = n_docs // 10
    src = torch.randint(0, n_docs, (ndup,), device="cuda", generator=g)
    dst = torch.randint(0, n_docs, (ndup,), device="cuda", generator=g)
    tok[dst] = tok[src]
    mask = torch.rand((ndup, t), device="cuda", generator=g) < 0.1
    tok[dst] = torch.where(mask, torch.randint(-2 ** 31, 2 ** 31 - 1, (ndup, t), dtype=torch.int32, device="cuda", generator=g), tok[dst])
    off = torch.arange(0, (n_docs + 1) * t, t, dtype=torch.int64, device="cuda")
    sig = torch.empty((n_docs, k), dtype=torch.int32, device="cuda")
    ms_sig = ev_time(lambda: dsk.engine.bulk_signatures_device(tok.view(-1), off, n_docs * t, perms, d_out=sig))
    warm = dsk.GpuLSH(threshold=0.8, num_perm=k, capacity=4096)
    warm.insert(sig[:4096])                      # first launch of the kernels (module load, attribute set) is not timed
    warm.query(sig[:64], to_host=False)
    del warm
    lsh = dsk.GpuLSH(threshold=0.8, num_perm=k, capacity=n_docs)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    lsh.insert(sig)
    e1.record()
    torch.cuda.synchronize()
    ms_ins = e0.elapsed_time(e1)
    # fused: tokens -> signatures in the index's storage + bucket updates in ONE launch (GpuLSH.insert_tokens); three fresh
    # indexes, the first one warms the kernel up; candidates of the same queries must be identical to the two-step index
    ms_fused = []
    fused = None
    for rep in range(3 if fused_leg else 0):
        del fused
        fused = dsk.GpuLSH(threshold=0.8, num_perm=k, capacity=n_docs)
        torch.cuda.synchronize()
        e0.record()
        fused.insert_tokens(tok.view(-1), off, n_docs * t, perms)
        e1.record()
        torch.cuda.synchronize()
        ms_fused.append(e0.elapsed_time(e1))
    q = sig[torch.randint(0, n_docs, (n_query,), device="cuda", generator=g)]
    same = None
    if fused_leg:
        pf, xf = fused.query(q[:20_000], to_host=False)
        pp, xp = lsh.query(q[:20_000], to_host=False)
        same = bool(torch.equal(pf, pp)) and bool(torch.equal(
            torch.sort(xf.long() + torch.repeat_interleave(torch.arange(20_000, device="cuda"), pf[1:] - pf[:-1]) * n_docs)[0],
            torch.sort(xp.long() + torch.repeat_interleave(torch.arange(20_000, device="cuda"), pp[1:] - pp[:-1]) * n_docs)[0]))
        del fused
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ptr, idx = lsh.query(q, to_host=False)
    torch.cuda.synchronize()
    ms_q = (time.perf_counter() - t0) * 1e3
    # CPU oracle on a small sample: dict insert + query per document
    ns = 20_000
    hs = sig[:ns].cpu().numpy().view(np.uint32).astype(np.uint64)
    ref = o.DictLSH(k, lsh.b, lsh.r)
    t0 = time.perf_counter()
    for i, row in enumerate(hs):
        ref.insert(i, row, check_duplication=False)
    cpu_ins = (time.perf_counter() - t0) / ns
    t0 = time.perf_counter()
    for row in hs[:5000]:
        ref.query(row)
    cpu_q = (time.perf_counter() - t0) / 5000
    return {"config": "C3 (1 GPU)", "docs": n_docs, "tokens": t, "num_perm": k, "b": lsh.b, "r": lsh.r,
            "signature_ms": ms_sig, "signatures_per_s": n_docs / ms_sig * 1e3,
            "lsh_insert_ms": ms_ins, "lsh_insert_docs_per_s": n_docs / ms_ins * 1e3,
            "fused_signatures_plus_insert_ms": min(ms_fused[1:]) if ms_fused else None, "fused_runs_ms": ms_fused,
            "two_step_signatures_plus_insert_ms": ms_sig + ms_ins, "fused_candidates_identical": same,
            "lsh_query_ms": ms_q, "queries": n_query, "lsh_queries_per_s": n_query / ms_q * 1e3,
            "candidates_total": int(ptr[-1].item()),
            "cpu_oracle_insert_us_per_doc": cpu_ins * 1e6, "cpu_oracle_query_us_per_doc": cpu_q * 1e6}


def c4(n_vec, dim=4096, ss=128):
    gen = dsk.WeightedMinHashGenerator(dim, ss, 1)
    g = torch.Generator(device="cuda").manual_seed(4)
    v = torch.rand((n_vec, dim), device="cuda", generator=g) * 10
    v[:, ::7] = 0
    gen.minhash_batch(v[:8])  # create the handle
    from datasketch_b200 import _native as nv
    h = gen._handle(0)
    out = torch.empty((n_vec, ss, 2), dtype=torch.int64, device="cuda")
    st = torch.empty((n_vec,), dtype=torch.int32, device="cuda")
    ms = ev_time(lambda: nv.check(nv.load().dsk_wmh_minhash(h, v.data_ptr(), n_vec, out.data_ptr(), st.data_ptr(), 0,
                                                             torch.cuda.current_stream().cuda_stream)), iters=2)
    par = o.wmh_params(dim, ss, 1)
    vs = v[:20].cpu().numpy()
    t0 = time.perf_counter()
    want = np.stack([o.wmh_minhash(x, *par) for x in vs])
    cpu = (time.perf_counter() - t0) / len(vs)
    got = out[:20].cpu().numpy()
    return {"config": "C4", "vectors": n_vec, "dim": dim, "sample_size": ss, "ms": ms,
            "vectors_per_s": n_vec / ms * 1e3, "evals_per_s": n_vec * dim * ss / ms * 1e3,
            "cpu_oracle_ms_per_vector": cpu * 1e3, "sampled_rows_identical": bool(np.array_equal(got, want)),
            "sampled_mismatching_samples": int((got != want).any(axis=2).sum())}


def c5(n_rows, k=128, topk=10):
    """All-pairs top-k: the exact kernel and the fingerprint-prefilter variant on the same matrix (identical lists
    required), on signatures without similarity (the C2 shape: every count is 0, ties by index) plus planted duplicates."""
    g = torch.Generator(device="cuda").manual_seed(5)
    sig = torch.randint(-2 ** 31, 2 ** 31 - 1, (n_rows, k), dtype=torch.int32, device="cuda", generator=g)
    sig[1::2][: n_rows // 4] = sig[0::2][: n_rows // 4]       # planted duplicates
    res = {}
    for name, pf in (("exact", False), ("prefilter", True)):
        dsk.codec.jaccard_topk(sig[:256], sig[:4096], topk=topk, self_base=0, to_host=False, prefilter=pf)   # first launch
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        res[name] = dsk.codec.jaccard_topk(sig, sig, topk=topk, self_base=0, to_host=False, prefilter=pf)
        e1.record()
        torch.cuda.synchronize()
        res[name + "_ms"] = e0.elapsed_time(e1)
    same = bool(torch.equal(res["exact"][0], res["prefilter"][0]) and torch.equal(res["exact"][1], res["prefilter"][1]))
    # brute-force check of sampled rows
    rows = torch.linspace(0, n_rows - 1, 64, device="cuda").long()
    cnts = (sig[rows][:, None, :] == sig[None, :, :]).sum(dim=2)
    cnts[torch.arange(len(rows)), rows] = -1
    wrong = 0
    for i, r in enumerate(rows.tolist()):
        c = cnts[i]
        order = torch.argsort(-c * (n_rows + 1) + torch.arange(n_rows, device="cuda"))[:topk]   # count desc, index asc
        wrong += int(not torch.equal(order, res["prefilter"][1][r]))
    a = sig[:2000].cpu().numpy().view(np.uint32).astype(np.uint64)
    t0 = time.perf_counter()
    for i in range(0, 2000, 2):
        o.jaccard(a[i], a[i + 1])
    cpu = (time.perf_counter() - t0) / 1000
    ms = res["prefilter_ms"]
    return {"config": "C5 (1 GPU)", "rows": n_rows, "num_perm": k, "topk": topk, "ms": ms, "ms_exact_kernel": res["exact_ms"],
            "speedup_vs_exact_kernel": res["exact_ms"] / ms, "lists_identical_to_exact_kernel": same, "rows_wrong": wrong,
            "pairs_per_s": n_rows * n_rows / ms * 1e3, "compares_per_s": n_rows * n_rows * k / ms * 1e3,
            "cpu_oracle_us_per_pair": cpu * 1e6}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--c3-docs", type=int, default=2_000_000)
    ap.add_argument("--c4-vecs", type=int, default=20_000)
    ap.add_argument("--c5-rows", type=int, default=100_000)
    ap.add_argument("--c3-fused", action="store_true",
                    help="also time GpuLSH.insert_tokens (fused signatures + insert); this leg has not completed a GPU run yet "
                         "(profiles/README.md r2w_*), tools/diag_fused.py holds the fused timing of record")
    a = ap.parse_args()
    if a.c3_docs > 0:
        print(json.dumps(c3(a.c3_docs, fused_leg=a.c3_fused)), flush=True)
    for fn, arg in ((c4, a.c4_vecs), (c5, a.c5_rows)):
        if arg > 0:
            print(json.dumps(fn(arg)), flush=True)
