"""Config C3 (10M docs x 128 tokens, K=256, MinHashLSH(0.8) insert + query) sharded over the GPUs of one
node: each rank builds the signatures of its documents, inserts them into its shard of the index
(ShardedLSH: no communication) and all ranks query (all-gather of queries + all-to-all of answers).

    python -m torch.distributed.run --nproc-per-node N tools/bench_c3_multi.py --docs-per-gpu 1250000
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import datasketch_b200 as dsk  # noqa: E402
from datasketch_b200.distributed import ShardedLSH  # noqa: E402
from datasketch_b200.minhash import _make_permutations  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--docs-per-gpu", type=int, default=1_250_000)
ap.add_argument("--tokens", type=int, default=128)
ap.add_argument("--num-perm", type=int, default=256)
ap.add_argument("--queries-per-gpu", type=int, default=12_500)
a = ap.parse_args()

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ["NCCL_DEBUG"] = "WARN"
    dist.init_process_group("nccl", device_id=dev)

n, t, k = a.docs_per_gpu, a.tokens, a.num_perm
perms = _make_permutations(k, 1)
g = torch.Generator(device="cuda").manual_seed(100 + rank)
tok = torch.randint(-2 ** 31, 2 ** 31 - 1, (n, t), dtype=torch.int32, device=dev, generator=g)
# a common pool of documents shared by all ranks -> cross-rank near-duplicates exist
gp = torch.Generator(device="cuda").manual_seed(7)
pool = torch.randint(-2 ** 31, 2 ** 31 - 1, (n // 20, t), dtype=torch.int32, device=dev, generator=gp)
tok[: n // 20] = pool
mask = torch.rand((n // 20, t), device=dev, generator=g) < 0.05
tok[: n // 20] = torch.where(mask, torch.randint(-2 ** 31, 2 ** 31 - 1, (n // 20, t), dtype=torch.int32, device=dev, generator=g), tok[: n // 20])
off = torch.arange(0, (n + 1) * t, t, dtype=torch.int64, device=dev)
sig = torch.empty((n, k), dtype=torch.int32, device=dev)


def sync():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


dsk.engine.bulk_signatures_device(tok.view(-1), off, n * t, perms, d_out=sig)  # warm-up
sync()
t0 = time.perf_counter()
dsk.engine.bulk_signatures_device(tok.view(-1), off, n * t, perms, d_out=sig)
sync()
t_sig = time.perf_counter() - t0

ix = ShardedLSH(threshold=0.8, num_perm=k, capacity=n, device=local)
sync()
t0 = time.perf_counter()
ix.insert(sig)
sync()
t_ins = time.perf_counter() - t0

q = sig[: a.queries_per_gpu].contiguous()      # pool documents: their near-duplicates live on every rank
ptr, idx = ix.query(q)                          # warm-up (NCCL channels)
sync()
t0 = time.perf_counter()
ptr, idx = ix.query(q)
sync()
t_q = time.perf_counter() - t0

# check: every query finds itself, and finds its pool twin on the other ranks
self_ok = bool(((idx[ptr[:-1]] >= 0)).all().item()) if idx.numel() else False
nq = q.shape[0]
cnt = (ptr[1:] - ptr[:-1]).float()
own = torch.arange(nq, device=dev) + ix.base
found_self = 0
ph, ih = ptr.cpu().numpy(), idx.cpu().numpy()
for i in range(0, nq, max(1, nq // 200)):
    found_self += int(own[i].item() in ih[ph[i]:ph[i + 1]])
stats = torch.tensor([cnt.mean().item(), float(found_self)], device=dev)
if world > 1:
    dist.all_reduce(stats)
if rank == 0:
    print(json.dumps({"config": "C3 sharded", "n_gpus": world, "docs_total": n * world, "tokens": t, "num_perm": k,
                      "b": ix.index.b, "r": ix.index.r,
                      "signature_s": t_sig, "signatures_per_s": n * world / t_sig,
                      "lsh_insert_s": t_ins, "lsh_insert_docs_per_s": n * world / t_ins,
                      "lsh_query_s": t_q, "queries_total": nq * world, "lsh_queries_per_s": nq * world / t_q,
                      "mean_candidates_per_query": stats[0].item() / world,
                      "sampled_queries_that_found_themselves": int(stats[1].item()),
                      "sampled_queries": len(range(0, nq, max(1, nq // 200))) * world}), flush=True)
if world > 1:
    dist.destroy_process_group()
