"""Config C5 (all-pairs Jaccard over 1M signatures, num_perm=128, top-10) sharded over the GPUs of one node:
all-gather of the signature matrix (the path's one exchange), then every rank ranks the whole corpus for its
own rows (datasketch_b200.distributed.sharded_jaccard_topk).  A sample of rows per rank is checked exactly
against a brute-force torch count on the device.

    python -m torch.distributed.run --nproc-per-node N tools/bench_c5_multi.py --rows-per-gpu 125000
"""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from datasketch_b200.distributed import gather_signature_blocks, sharded_jaccard_topk  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows-per-gpu", type=int, default=125_000)
ap.add_argument("--num-perm", type=int, default=128)
ap.add_argument("--topk", type=int, default=10)
ap.add_argument("--check-rows", type=int, default=64)
a = ap.parse_args()

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ["NCCL_DEBUG"] = "WARN"
    dist.init_process_group("nccl", device_id=dev)

n, k, topk = a.rows_per_gpu, a.num_perm, a.topk
g = torch.Generator(device="cuda").manual_seed(500 + rank)
sig = torch.randint(-2 ** 31, 2 ** 31 - 1, (n, k), dtype=torch.int32, device=dev, generator=g)
# rows [0, n/50) of every rank are perturbed copies of one common pool: near-duplicates across ranks
gp = torch.Generator(device="cuda").manual_seed(9)
npool = max(n // 50, 1)
pool = torch.randint(-2 ** 31, 2 ** 31 - 1, (npool, k), dtype=torch.int32, device=dev, generator=gp)
keep = torch.rand((npool, k), device=dev, generator=g) < 0.8
sig[:npool] = torch.where(keep, pool, sig[:npool])


def sync():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


sharded_jaccard_topk(sig[:1024].contiguous(), topk=topk)  # warm-up (NCCL channels, kernel load)
sync()
t0 = time.perf_counter()
cnt, idx = sharded_jaccard_topk(sig, topk=topk)
sync()
t_all = time.perf_counter() - t0

# exact check of a sample of this rank's rows against a brute-force count over the gathered matrix
full, base, counts = gather_signature_blocks(sig)
n_total = full.shape[0]
rows = torch.linspace(0, n - 1, a.check_rows, device=dev).long().unique()
bad = 0
for i in rows.tolist():
    c = (full == sig[i][None, :]).sum(dim=1).to(torch.int64)
    c[base + i] = -1                                              # a row never lists itself
    key = c * n_total + (n_total - 1 - torch.arange(n_total, device=dev))   # count desc, index asc
    top = torch.topk(key, topk).indices
    bad += int(not (torch.equal(top, idx[i]) and torch.equal(c[top].to(torch.int32), cnt[i])))
pool_hits = int((cnt[:npool, 0] > k // 2).sum().item())         # pool rows must see their twins on other ranks
stats = torch.tensor([float(bad), float(len(rows)), float(pool_hits), float(npool)], device=dev)
if world > 1:
    dist.all_reduce(stats)
if rank == 0:
    print(json.dumps({"config": "C5 sharded", "n_gpus": world, "rows_total": n_total, "num_perm": k, "topk": topk,
                      "seconds": t_all, "element_compares_per_s": n_total * n_total * k / t_all,
                      "row_pairs_per_s": n_total * n_total / t_all,
                      "rows_checked_exactly": int(stats[1].item()), "rows_wrong": int(stats[0].item()),
                      "pool_rows_with_cross_rank_twin_found": int(stats[2].item()), "pool_rows": int(stats[3].item())}),
          flush=True)
assert bad == 0
if world > 1:
    dist.destroy_process_group()
