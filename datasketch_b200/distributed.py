"""Multi-GPU layer: one process per GPU (torch.distributed, NCCL over NVLink/NVSwitch).

Documents shard embarrassingly: a document's signature depends only on its own tokens and the
replicated (a, b) permutation table (2-4 KB), so each rank builds the signatures of one contiguous,
token-balanced range of documents with no data-path collective.  The only exchange step of the
path is assembling the full [N, K] signature matrix on every rank for LSH bucketing / all-pairs
Jaccard: one all-gather of the per-rank blocks (rule: signatures of disjoint document ranges simply
concatenate; a *single* set split across ranks would combine by element-wise min,
datasketch/minhash.py:337-359 -- not needed here).

The ``compute`` argument exists so the sharding / gather logic can be exercised on CPU with the
gloo backend (tests inject a checker); the default is the GPU engine and there is no CPU fallback.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import numpy as np


def shard_bounds(offsets: np.ndarray, world: int) -> List[Tuple[int, int]]:
    """Contiguous document ranges [d0, d1) per rank, balanced by token count (CSR offsets)."""
    offsets = np.asarray(offsets, dtype=np.int64)
    n = len(offsets) - 1
    base, total = int(offsets[0]), int(offsets[-1] - offsets[0])
    cuts = [0]
    for r in range(1, world):
        target = base + (total * r) // world
        cuts.append(int(np.searchsorted(offsets[:-1], target, side="left")))
    cuts.append(n)
    for i in range(1, len(cuts)):       # monotone (empty shards are allowed)
        cuts[i] = max(cuts[i], cuts[i - 1])
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def _gpu_compute(tokens: np.ndarray, offsets: np.ndarray, permutations: np.ndarray, device: int):
    import torch
    from . import engine
    n = len(offsets) - 1
    k = permutations.shape[1]
    if n == 0:
        return torch.empty((0, k), dtype=torch.int32, device=torch.device("cuda", device))
    tok, is64 = engine._as_tokens(tokens)
    if is64:
        raise ValueError("sharded bulk path takes 32-bit token hashes")
    d_tok = torch.from_numpy(tok.view(np.int32)).cuda(device)
    d_off = torch.from_numpy(np.ascontiguousarray(offsets, dtype=np.int64)).cuda(device)
    if d_tok.numel() == 0:
        d_tok = torch.zeros(4, dtype=torch.int32, device=d_off.device)
    return engine.bulk_signatures_device(d_tok, d_off, int(tok.size), permutations)


def sharded_bulk_signatures(tokens: np.ndarray, offsets: np.ndarray, permutations: np.ndarray, group=None,
                            gather: bool = True, compute: Optional[Callable] = None, device: Optional[int] = None):
    """Every rank passes the same CSR batch (or at least its own slice of it); rank r builds the
    signatures of its token-balanced document range and, with ``gather=True``, one all-gather
    assembles the full [N, K] int32 (u32 bit pattern) matrix on every rank.

    Returns ``(signatures, (d0, d1))``: the full matrix (gather) or this rank's block.
    """
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    offsets = np.asarray(offsets, dtype=np.int64)
    bounds = shard_bounds(offsets, world)
    d0, d1 = bounds[rank]
    t0, t1 = int(offsets[d0]), int(offsets[d1])
    loc_off = offsets[d0:d1 + 1] - t0
    loc_tok = np.asarray(tokens)[t0:t1]
    if compute is None:
        if device is None:
            device = torch.cuda.current_device()
        local = _gpu_compute(loc_tok, loc_off, permutations, device)
    else:
        local = compute(loc_tok, loc_off, permutations)
    if not gather or world == 1:
        return local, (d0, d1)
    k = local.shape[1]
    rows = max(b - a for a, b in bounds)
    padded = torch.zeros((rows, k), dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    full = torch.empty((world * rows, k), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(full, padded, group=group)
    n = len(offsets) - 1
    if all(b - a == rows for a, b in bounds):
        return full[:n], (d0, d1)
    out = torch.empty((n, k), dtype=local.dtype, device=local.device)
    for r, (a, b) in enumerate(bounds):
        out[a:b] = full[r * rows: r * rows + (b - a)]
    return out, (d0, d1)
