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


class ShardedLSH:
    """MinHashLSH over documents sharded across ranks (config C3: insert + query on 8 GPUs).

    Every rank owns a device-resident ``GpuLSH`` over *its* documents (all b bands).  Buckets never
    span ranks' documents, so insert needs no communication.  ``query`` all-gathers the query
    signatures (the one exchange of the path), every rank answers all queries against its shard, and
    the answers return to the asking rank with two all-to-alls (per-query counts, then document
    numbers).  Document numbers are global: rank r's i-th inserted document is ``base[r] + i``
    where ``base`` is the exclusive prefix of the per-rank document counts.

    ``index_factory`` exists for the CPU/gloo tests; the default is ``lsh.GpuLSH`` (no CPU fallback).
    """

    def __init__(self, threshold: float = 0.9, num_perm: int = 128, params=None, capacity: int = 1 << 20,
                 group=None, index_factory: Optional[Callable] = None, device: Optional[int] = None):
        import torch
        import torch.distributed as dist
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        if index_factory is None:
            from .lsh import GpuLSH
            if device is None:
                device = torch.cuda.current_device()
            self.index = GpuLSH(threshold=threshold, num_perm=num_perm, params=params, capacity=capacity, device=device)
        else:
            self.index = index_factory(threshold=threshold, num_perm=num_perm, params=params, capacity=capacity)
        self.h = num_perm
        self.n_local = 0
        self.base = 0
        self.counts = [0] * self.world

    def insert(self, sig_local) -> None:
        """Insert this rank's signature block; renumbers the global document ids (collective)."""
        import torch
        import torch.distributed as dist
        self.index.insert(sig_local)
        self.n_local += int(sig_local.shape[0])
        if self.world > 1:
            t = torch.tensor([self.n_local], dtype=torch.int64, device=sig_local.device)
            allc = [torch.zeros_like(t) for _ in range(self.world)]
            dist.all_gather(allc, t, group=self.group)
            self.counts = [int(x.item()) for x in allc]
        else:
            self.counts = [self.n_local]
        self.base = sum(self.counts[: self.rank])

    def query(self, sig_q_local):
        """Candidates (global document numbers) of this rank's queries: ``(ptr int64[Q+1], idx int64[total])``
        as tensors on the queries' device.  Collective: every rank must call it."""
        import torch
        import torch.distributed as dist
        dev = sig_q_local.device
        nq_me = int(sig_q_local.shape[0])
        if self.world == 1:
            ptr, idx = self.index.query(sig_q_local, to_host=False)
            return ptr, idx.to(torch.int64)
        # 1. all-gather the queries (padded to the largest per-rank count)
        t = torch.tensor([nq_me], dtype=torch.int64, device=dev)
        allq = [torch.zeros_like(t) for _ in range(self.world)]
        dist.all_gather(allq, t, group=self.group)
        nq = [int(x.item()) for x in allq]
        rows = max(max(nq), 1)
        padded = torch.zeros((rows, self.h), dtype=sig_q_local.dtype, device=dev)
        padded[:nq_me] = sig_q_local
        gathered = torch.empty((self.world * rows, self.h), dtype=sig_q_local.dtype, device=dev)
        dist.all_gather_into_tensor(gathered, padded, group=self.group)
        sel = torch.cat([torch.arange(r * rows, r * rows + nq[r], device=dev) for r in range(self.world)])
        q_all = gathered[sel].contiguous()
        # 2. answer every query against this rank's shard
        ptr_all, idx_all = self.index.query(q_all, to_host=False)
        idx_all = idx_all.to(torch.int64) + self.base
        cnt_all = (ptr_all[1:] - ptr_all[:-1]).contiguous()
        # 3. return answers to the asking ranks: counts, then payload
        qoff = np.concatenate([[0], np.cumsum(nq)]).astype(np.int64)
        cnt_back = torch.empty((self.world * nq_me,), dtype=torch.int64, device=dev)
        dist.all_to_all_single(cnt_back, cnt_all, [nq_me] * self.world, nq, group=self.group)
        ptr_host = ptr_all.cpu().numpy()
        send_sizes = [int(ptr_host[qoff[r + 1]] - ptr_host[qoff[r]]) for r in range(self.world)]
        cnt_back2 = cnt_back.view(self.world, nq_me)
        recv_sizes = [int(x) for x in cnt_back2.sum(dim=1).cpu().tolist()]
        payload = torch.empty((sum(recv_sizes),), dtype=torch.int64, device=dev)
        dist.all_to_all_single(payload, idx_all, recv_sizes, send_sizes, group=self.group)
        # 4. merge: a query's candidates = concatenation over source ranks (disjoint document sets)
        tot = cnt_back2.sum(dim=0)
        ptr = torch.zeros((nq_me + 1,), dtype=torch.int64, device=dev)
        ptr[1:] = torch.cumsum(tot, 0)
        out = torch.empty((int(ptr[-1].item()),), dtype=torch.int64, device=dev)
        before = torch.cumsum(cnt_back2, 0) - cnt_back2            # candidates of earlier source ranks, per query
        pos0 = 0
        for s in range(self.world):
            c = cnt_back2[s]
            n_s = recv_sizes[s]
            if n_s:
                src_excl = torch.cumsum(c, 0) - c
                dest0 = ptr[:-1] + before[s]
                shift = torch.repeat_interleave(dest0 - src_excl, c)
                out[shift + torch.arange(n_s, device=dev)] = payload[pos0:pos0 + n_s]
            pos0 += n_s
        return ptr, out


def gather_signature_blocks(sig_local, group=None):
    """All-gather per-rank [n_r, K] signature blocks (possibly of different heights) into the full
    [sum n_r, K] matrix in rank order, on every rank.  Returns ``(full, base, counts)`` where ``base`` is this
    rank's first global row.  This is the one exchange step of the path (see module docstring)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n_me = int(sig_local.shape[0])
    if world == 1:
        return sig_local, 0, [n_me]
    dev = sig_local.device
    t = torch.tensor([n_me], dtype=torch.int64, device=dev)
    allc = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(allc, t, group=group)
    counts = [int(x.item()) for x in allc]
    rows = max(max(counts), 1)
    if min(counts) == rows:
        padded = sig_local.contiguous()
    else:
        padded = torch.zeros((rows, sig_local.shape[1]), dtype=sig_local.dtype, device=dev)
        padded[:n_me] = sig_local
    gathered = torch.empty((world * rows, sig_local.shape[1]), dtype=sig_local.dtype, device=dev)
    dist.all_gather_into_tensor(gathered, padded, group=group)
    if min(counts) == rows:
        full = gathered
    else:
        full = torch.cat([gathered[r * rows: r * rows + counts[r]] for r in range(world)])
    return full, sum(counts[:rank]), counts


def sharded_jaccard_topk(sig_local, topk: int = 10, group=None, topk_fn: Optional[Callable] = None):
    """All-pairs top-k Jaccard over signatures sharded across ranks (config C5).

    Every rank all-gathers the signature matrix, then ranks the *whole* corpus for its own rows only
    (``dsk_jaccard_topk`` with ``self_base`` = its first global row, so a row never lists itself):
    the quadratic work splits evenly and the answers need no second exchange.  Returns
    ``(count int32 [n_local, topk], index int64 [n_local, topk])`` device tensors with global row
    numbers; count / K is the reference's ``jaccard`` (datasketch/minhash.py:324).

    ``topk_fn(queries, db, topk, self_base)`` exists for the CPU/gloo tests; default = the GPU kernel."""
    full, base, _ = gather_signature_blocks(sig_local, group)
    if topk_fn is None:
        from . import codec
        return codec.jaccard_topk(sig_local, full, topk=topk, self_base=base, to_host=False)
    return topk_fn(sig_local, full, topk, base)


class FusedGather:
    """Signature build fused with the all-gather (NVLink peer stores instead of a separate NCCL pass).

    Holds a symmetric-memory [n_total, K] int32 matrix (``torch.distributed._symmetric_memory``:
    every rank's buffer is mapped into every other rank's address space over NVLink/NVSwitch).
    ``build`` runs ``dsk_minhash_bulk_gather``: the kernel stores each finished signature row into
    the same row of all ranks' matrices, so the transfer overlaps the integer math document by
    document; a stream sync + barrier then makes the full matrix valid on every rank.
    """

    def __init__(self, n_total: int, num_perm: int, group=None, device=None):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        if self.world > 8:
            raise ValueError("FusedGather supports up to 8 ranks (one NVSwitch domain)")
        if device is None:
            device = torch.cuda.current_device()
        self.device = torch.device("cuda", device) if isinstance(device, int) else device
        self.n_total, self.k = int(n_total), int(num_perm)
        self.buf = symm_mem.empty((self.n_total, self.k), dtype=torch.int32, device=self.device)
        self.hdl = symm_mem.rendezvous(self.buf, self.group)
        self.ptrs = [int(p) for p in self.hdl.buffer_ptrs]

    def build(self, d_tokens, d_offsets, n_tokens: int, permutations: np.ndarray, row_offset: int,
              kernel: str = "auto", sync: bool = True):
        """This rank's documents -> rows [row_offset, row_offset + n_docs) of every rank's matrix."""
        import ctypes
        import torch
        from . import _native as nv
        from .engine import KERNELS
        n = d_offsets.numel() - 1
        if row_offset < 0 or row_offset + n > self.n_total:
            raise ValueError("row range exceeds the gathered matrix")
        h = nv.perm_handle(permutations, self.device.index)
        if h.num_perm != self.k:
            raise ValueError("num_perm mismatch")
        arr = (ctypes.c_void_p * self.world)(*self.ptrs)
        with torch.cuda.device(self.device):
            st = torch.cuda.current_stream(self.device)
            nv.check(nv.load().dsk_minhash_bulk_gather(h.handle, d_tokens.data_ptr() if n_tokens else None,
                                                        int(d_tokens.element_size() == 8), d_offsets.data_ptr(), n,
                                                        n_tokens, ctypes.cast(arr, ctypes.c_void_p), self.world,
                                                        row_offset, 0, KERNELS[kernel], st.cuda_stream))
            if sync:
                self.finish()
        return self.buf

    def finish(self) -> None:
        """All ranks' peer stores have landed: stream sync, then a cross-rank barrier."""
        import torch
        torch.cuda.current_stream(self.device).synchronize()
        self.hdl.barrier()
