"""CPU, world_size 2 over gloo: the N>1 host logic (token-balanced document sharding and the
all-gather that assembles the signature matrix).  The per-rank compute is injected (the oracle,
as a checker stand-in) because this container has no GPU; the product default is the CUDA engine."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oracle_clib as oc
from oracle import oracle_np as o


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, tok, off, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from datasketch_b200.distributed import shard_bounds, sharded_bulk_signatures
    P = o.init_permutations(64, 5)

    def compute(t, f, perms):
        return torch.from_numpy(oc.minhash_bulk_u32tok(t, f, perms).view(np.int32).copy())

    full, (d0, d1) = sharded_bulk_signatures(tok, off, P, compute=compute)
    local, _ = sharded_bulk_signatures(tok, off, P, compute=compute, gather=False)
    assert (d0, d1) == shard_bounds(off, world)[rank] and local.shape[0] == d1 - d0
    ret[rank] = (full.numpy().view(np.uint32).copy(), d0, d1)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("ragged", [False, True])
def test_sharded_signatures_gloo_world2(ragged):
    rs = np.random.RandomState(3)
    n = 101
    lens = rs.randint(0, 60, size=n) if ragged else np.full(n, 32)
    if ragged:
        lens[:40] = 0        # leading empty documents -> unbalanced document counts per shard
    off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    tok = rs.randint(0, 2 ** 32, size=int(off[-1]), dtype=np.uint64).astype(np.uint32)
    want = oc.minhash_bulk_u32tok(tok, off, o.init_permutations(64, 5))
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), tok, off, ret), nprocs=2, join=True)
    (f0, a0, b0), (f1, a1, b1) = ret[0], ret[1]
    assert a0 == 0 and b0 == a1 and b1 == n
    assert np.array_equal(f0, want) and np.array_equal(f1, want)


def test_shard_bounds_properties():
    from datasketch_b200.distributed import shard_bounds
    rs = np.random.RandomState(0)
    for world in (1, 2, 3, 8):
        for _ in range(20):
            n = rs.randint(0, 50)
            off = np.zeros(n + 1, dtype=np.int64)
            np.cumsum(rs.randint(0, 100, size=n), out=off[1:])
            b = shard_bounds(off, world)
            assert len(b) == world and b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1)) and all(x <= y for x, y in b)
            if n and off[-1] > 0:
                per = [off[y] - off[x] for x, y in b]
                assert max(per) <= off[-1] / world + np.diff(off).max()


class _OracleIndex:
    """Stand-in for GpuLSH in the CPU test: the dict oracle behind the same insert/query interface."""

    def __init__(self, threshold, num_perm, params, capacity):
        self.b, self.r = params
        self.ix = o.DictLSH(num_perm, self.b, self.r)
        self.n = 0

    def insert(self, sig):
        for row in sig.numpy().view(np.uint32).astype(np.uint64):
            self.ix.insert(self.n, row)
            self.n += 1

    def query(self, sig, to_host=False):
        ptr, idx = [0], []
        for row in sig.numpy().view(np.uint32).astype(np.uint64):
            idx.extend(sorted(self.ix.query(row)))
            ptr.append(len(idx))
        return torch.tensor(ptr, dtype=torch.int64), torch.tensor(idx, dtype=torch.int32)


def _lsh_worker(rank, world, port, sig, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from datasketch_b200.distributed import ShardedLSH
    n = len(sig)
    cut = [0, 130, n]                                   # uneven shards
    mine = torch.from_numpy(sig[cut[rank]:cut[rank + 1]].view(np.int32).copy())
    ix = ShardedLSH(num_perm=sig.shape[1], params=(8, 4), capacity=1000, index_factory=_OracleIndex)
    ix.insert(mine)
    assert ix.base == cut[rank] and ix.counts == [130, n - 130]
    qsel = np.arange(rank, n, 3)                        # each rank asks about different documents
    ptr, idx = ix.query(torch.from_numpy(sig[qsel].view(np.int32).copy()))
    ret[rank] = (qsel, ptr.numpy().copy(), idx.numpy().copy())
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_lsh_gloo_world2():
    rs = np.random.RandomState(9)
    n, k = 300, 32
    sig = rs.randint(0, 3, size=(n, k)).astype(np.uint32)
    sig[rs.randint(0, n, 40)] = sig[rs.randint(0, n, 40)]
    ref = o.DictLSH(k, 8, 4)
    for i, row in enumerate(sig):
        ref.insert(i, row.astype(np.uint64))
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_lsh_worker, args=(2, _free_port(), sig, ret), nprocs=2, join=True)
    for rank in (0, 1):
        qsel, ptr, idx = ret[rank]
        assert ptr[0] == 0 and ptr[-1] == len(idx)
        for j, qi in enumerate(qsel):
            assert sorted(idx[ptr[j]:ptr[j + 1]].tolist()) == sorted(ref.query(sig[qi].astype(np.uint64)))


def _np_topk(q, db, topk, self_base):
    """Checker for all-pairs top-k: count of equal positions, best first, ties -> lower index."""
    qn, dbn = q.numpy().view(np.uint32), db.numpy().view(np.uint32)
    out_c = np.full((len(qn), topk), -1, np.int32)
    out_i = np.full((len(qn), topk), -1, np.int64)
    for i, row in enumerate(qn):
        cnt = (dbn == row[None, :]).sum(axis=1).astype(np.int64)
        order = np.lexsort((np.arange(len(dbn)), -cnt))
        order = order[order != self_base + i][:topk]
        out_c[i, :len(order)] = cnt[order]
        out_i[i, :len(order)] = order
    return torch.from_numpy(out_c), torch.from_numpy(out_i)


def _topk_worker(rank, world, port, sig, cut, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from datasketch_b200.distributed import gather_signature_blocks, sharded_jaccard_topk
    mine = torch.from_numpy(sig[cut[rank]:cut[rank + 1]].view(np.int32).copy())
    full, base, counts = gather_signature_blocks(mine)
    assert base == cut[rank] and counts == [cut[1] - cut[0], cut[2] - cut[1]]
    assert np.array_equal(full.numpy().view(np.uint32), sig)
    cnt, idx = sharded_jaccard_topk(mine, topk=5, topk_fn=_np_topk)
    ret[rank] = (cnt.numpy().copy(), idx.numpy().copy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("cut", [[0, 90, 200], [0, 100, 200]])       # uneven (padded gather) and even shards
def test_sharded_jaccard_topk_gloo_world2(cut):
    rs = np.random.RandomState(11)
    n, k = 200, 32
    sig = rs.randint(0, 3, size=(n, k)).astype(np.uint32)
    sig[rs.randint(0, n, 30)] = sig[rs.randint(0, n, 30)]
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_topk_worker, args=(2, _free_port(), sig, cut, ret), nprocs=2, join=True)
    full = torch.from_numpy(sig.view(np.int32))
    wc, wi = _np_topk(full, full, 5, 0)                              # single-process answer over the whole corpus
    got_c = np.concatenate([ret[0][0], ret[1][0]])
    got_i = np.concatenate([ret[0][1], ret[1][1]])
    assert np.array_equal(got_c, wc.numpy()) and np.array_equal(got_i, wi.numpy())
