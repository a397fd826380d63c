"""Batch codecs over device-resident signature matrices (torch CUDA tensors as buffers).

``lean_pack`` / ``lean_unpack``: [N, K] signatures <-> LeanMinHash records
(``struct`` "<bo> q i {K}I", datasketch/lean_minhash.py:174-175, :201-214).
``band_keys``: the big-endian byte keys ``MinHashLSH._H`` builds per band
(datasketch/lsh.py:344, :537-538).  ``band_fingerprints``: 64-bit bucket ids.
"""
from __future__ import annotations

import struct

import numpy as np

from . import _native as nv

_BIG = {">": 1, "!": 1, "<": 0, "=": 0, "@": 0}


def _torch():
    import torch
    return torch


def _stream(t, stream):
    torch = _torch()
    return torch.cuda.current_stream(t.device).cuda_stream if stream is None else stream


def _as_device_sig(sig, device=0):
    """numpy or torch, u32/u64 -> contiguous CUDA tensor (int32/int64 storage) + is_u64 flag."""
    torch = _torch()
    if isinstance(sig, np.ndarray):
        if sig.dtype == np.uint32:
            t = torch.from_numpy(np.ascontiguousarray(sig).view(np.int32))
        elif sig.dtype == np.uint64:
            t = torch.from_numpy(np.ascontiguousarray(sig).view(np.int64))
        else:
            raise TypeError("signature matrix must be uint32 or uint64")
        nv.require_device(device)
        sig = t.cuda(device)
    if not sig.is_cuda or sig.dim() != 2:
        raise ValueError("signature matrix must be a 2-D CUDA tensor")
    return sig.contiguous(), int(sig.element_size() == 8)


def lean_record_size(num_perm: int, byteorder: str = "@") -> int:
    return struct.calcsize("%sqi%dI" % (byteorder, num_perm))


def lean_pack(sig, seed: int, byteorder: str = "@", out=None, stream=None):
    """[N, K] signatures -> [N, 12 + 4K] uint8 tensor of LeanMinHash records (one kernel)."""
    torch = _torch()
    d_sig, is64 = _as_device_sig(sig)
    n, k = d_sig.shape
    if lean_record_size(k, byteorder) != 12 + 4 * k:
        raise ValueError("native struct layout on this platform is not the packed q|i|I layout")
    if out is None:
        out = torch.empty((n, 12 + 4 * k), dtype=torch.uint8, device=d_sig.device)
    with torch.cuda.device(d_sig.device):
        nv.check(nv.load().dsk_lean_pack(d_sig.data_ptr(), is64, n, k, int(seed), _BIG[byteorder], out.data_ptr(),
                                         _stream(d_sig, stream)))
    return out


def lean_unpack(rec, num_perm: int, seed: int, byteorder: str = "@", out_u64: bool = False, stream=None):
    """[N, 12 + 4K] uint8 records -> [N, K] signatures; raises ValueError if any header differs."""
    torch = _torch()
    if isinstance(rec, np.ndarray):
        nv.require_device(0)
        rec = torch.from_numpy(np.ascontiguousarray(rec)).cuda()
    rec = rec.contiguous()
    n = rec.shape[0]
    if rec.numel() != n * (12 + 4 * num_perm):
        raise ValueError("record buffer size does not match num_perm")
    out = torch.empty((n, num_perm), dtype=torch.int64 if out_u64 else torch.int32, device=rec.device)
    status = torch.zeros(1, dtype=torch.int32, device=rec.device)
    with torch.cuda.device(rec.device):
        nv.check(nv.load().dsk_lean_unpack(rec.data_ptr(), n, num_perm, int(seed), _BIG[byteorder], out.data_ptr(),
                                           int(out_u64), status.data_ptr(), _stream(rec, stream)))
    if int(status.item()) != 0:
        raise ValueError("LeanMinHash record header (seed / num_perm) mismatch")
    return out


def band_keys(sig, b: int, r: int, stream=None):
    """[N, K] u32 signatures -> [N, b, 8r] uint8 big-endian band keys (lsh.py:537-538)."""
    torch = _torch()
    d_sig, is64 = _as_device_sig(sig)
    if is64:
        raise TypeError("band_keys takes the 32-bit signature matrix")
    n, k = d_sig.shape
    out = torch.empty((n, b, 8 * r), dtype=torch.uint8, device=d_sig.device)
    with torch.cuda.device(d_sig.device):
        nv.check(nv.load().dsk_band_keys(d_sig.data_ptr(), n, k, b, r, out.data_ptr(), _stream(d_sig, stream)))
    return out


def band_fingerprints(sig, b: int, r: int, stream=None):
    """[N, K] u32 signatures -> [N, b] int64 tensor holding the uint64 band fingerprints."""
    torch = _torch()
    d_sig, is64 = _as_device_sig(sig)
    if is64:
        raise TypeError("band_fingerprints takes the 32-bit signature matrix")
    n, k = d_sig.shape
    out = torch.empty((n, b), dtype=torch.int64, device=d_sig.device)
    with torch.cuda.device(d_sig.device):
        nv.check(nv.load().dsk_band_fingerprints(d_sig.data_ptr(), n, k, b, r, out.data_ptr(), _stream(d_sig, stream)))
    return out


def band_sums(sig, b: int, r: int, stream=None):
    """[N, K] u32 signatures -> [N, b] int64 tensor holding the uint64 Bloom keys of MinHashLSHBloom:
    ``sum(hashvalues[j*r:(j+1)*r]) % (2**61 - 1)`` per band (datasketch/lsh_bloom.py:105, :116)."""
    torch = _torch()
    d_sig, is64 = _as_device_sig(sig)
    if is64:
        raise TypeError("band_sums takes the 32-bit signature matrix")
    n, k = d_sig.shape
    out = torch.empty((n, b), dtype=torch.int64, device=d_sig.device)
    with torch.cuda.device(d_sig.device):
        nv.check(nv.load().dsk_band_sums(d_sig.data_ptr(), n, k, b, r, out.data_ptr(), _stream(d_sig, stream)))
    return out


def jaccard_pairs(sig, i, j, stream=None):
    """Jaccard estimates of the row pairs (i[p], j[p]) of one signature matrix: float64 array of
    ``count_equal / K`` exactly as ``MinHash.jaccard`` (minhash.py:324) computes it."""
    torch = _torch()
    d_sig, is64 = _as_device_sig(sig)
    if is64:
        raise TypeError("jaccard_pairs takes the 32-bit signature matrix")
    n, k = d_sig.shape
    di = torch.as_tensor(np.asarray(i, dtype=np.int64) if not hasattr(i, "is_cuda") else i).to(d_sig.device).contiguous()
    dj = torch.as_tensor(np.asarray(j, dtype=np.int64) if not hasattr(j, "is_cuda") else j).to(d_sig.device).contiguous()
    if di.shape != dj.shape:
        raise ValueError("pair index arrays differ in length")
    m = di.numel()
    out = torch.empty((max(m, 1),), dtype=torch.int32, device=d_sig.device)
    with torch.cuda.device(d_sig.device):
        nv.check(nv.load().dsk_jaccard_pairs(d_sig.data_ptr(), n, k, di.data_ptr(), dj.data_ptr(), m, out.data_ptr(),
                                             _stream(d_sig, stream)))
    cnt = out[:m].cpu().numpy()
    if (cnt < 0).any():
        raise IndexError("pair index out of range")
    return cnt.astype(np.float64) / float(k)


def jaccard_topk(queries, db, topk: int = 10, self_base: int = -1, to_host: bool = True, stream=None, prefilter: bool = True):
    """For every query row: the ``topk`` database rows with the highest Jaccard estimate
    (count of equal positions / K), best first, ties -> lower index.  ``self_base >= 0`` means
    query i *is* database row ``self_base + i`` and is left out of its own list.

    Returns ``(jaccard float64 [Q, topk], index int64 [Q, topk])`` (index -1 where the database
    has fewer than ``topk`` other rows)."""
    torch = _torch()
    d_q, q64 = _as_device_sig(queries)
    d_db, db64 = _as_device_sig(db)
    if q64 or db64:
        raise TypeError("jaccard_topk takes 32-bit signature matrices")
    if d_q.shape[1] != d_db.shape[1]:
        raise ValueError("Cannot compute Jaccard given MinHash with different numbers of permutation functions")
    nq, k = d_q.shape
    cnt = torch.empty((nq, topk), dtype=torch.int32, device=d_q.device)
    idx = torch.empty((nq, topk), dtype=torch.int64, device=d_q.device)
    with torch.cuda.device(d_q.device):
        # the fingerprint prefilter needs a workspace for the bit planes of both matrices (library-sized, owned here)
        ws_bytes = int(nv.load().dsk_jaccard_topk_workspace_size(nq, d_db.shape[0], k)) if prefilter else 0
        ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=d_q.device) if ws_bytes else None
        nv.check(nv.load().dsk_jaccard_topk_ws(d_q.data_ptr(), nq, d_db.data_ptr(), d_db.shape[0], k, topk, self_base,
                                               cnt.data_ptr(), idx.data_ptr(), ws.data_ptr() if ws is not None else None,
                                               ws_bytes, _stream(d_q, stream)))
        if ws is not None and stream is not None:
            ws.record_stream(torch.cuda.ExternalStream(stream))
    if not to_host:
        return cnt, idx
    c = cnt.cpu().numpy()
    return np.where(c >= 0, c, 0).astype(np.float64) / float(k), idx.cpu().numpy()


def _bbit_geometry(num_perm: int, b: int):
    if b < 0 or b > 32:
        raise ValueError("b must be an integer in [0, 32]")
    slot = 1 if b == 1 else 2 if b == 2 else 4 if b <= 4 else 8 if b <= 8 else 16 if b <= 16 else 32
    per = 64 // slot
    return slot, per, -(-num_perm // per)


def bbit_pack(sig, b: int, stream=None):
    """[N, K] u32 signatures -> [N, ceil(K / (64/slot))] int64 tensor holding the uint64 blocks a
    ``bBitMinHash`` pickles (datasketch/b_bit_minhash.py:78-92), for every row at once."""
    torch = _torch()
    d_sig, is64 = _as_device_sig(sig)
    if is64:
        raise TypeError("bbit_pack takes the 32-bit signature matrix")
    n, k = d_sig.shape
    _, _, nblk = _bbit_geometry(k, b)
    out = torch.empty((n, nblk), dtype=torch.int64, device=d_sig.device)
    with torch.cuda.device(d_sig.device):
        nv.check(nv.load().dsk_bbit_pack(d_sig.data_ptr(), n, k, int(b), out.data_ptr(), _stream(d_sig, stream)))
    return out


def bbit_unpack(blocks, num_perm: int, b: int, stream=None):
    """Inverse of :func:`bbit_pack`: [N, nblocks] uint64 blocks -> [N, K] masked values (int32 storage)."""
    torch = _torch()
    if isinstance(blocks, np.ndarray):
        nv.require_device(0)
        blocks = torch.from_numpy(np.ascontiguousarray(blocks).view(np.int64)).cuda()
    blocks = blocks.contiguous()
    _, _, nblk = _bbit_geometry(num_perm, b)
    if blocks.dim() != 2 or blocks.shape[1] != nblk:
        raise ValueError("block matrix does not match num_perm / b")
    out = torch.empty((blocks.shape[0], num_perm), dtype=torch.int32, device=blocks.device)
    with torch.cuda.device(blocks.device):
        nv.check(nv.load().dsk_bbit_unpack(blocks.data_ptr(), blocks.shape[0], num_perm, int(b), out.data_ptr(),
                                           _stream(blocks, stream)))
    return out
