"""Batch-native entry points over the C-ABI (the calls MinHash / MinHashLSH make).

Host code stays Python; every function here ends in one call into
libdsk_b200.so.  torch is used only to own device buffers and streams.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import threading

import numpy as np

from . import _native as nv

KERNELS = {"auto": nv.KERNEL_AUTO, "two_phase": nv.KERNEL_TWO_PHASE, "direct": nv.KERNEL_DIRECT,
           "exact": nv.KERNEL_EXACT}


def _as_tokens(tokens) -> Tuple[np.ndarray, int]:
    """Contiguous u32 (32-bit hash contract, minhash.py:69-70) or u64 token array."""
    t = np.asarray(tokens)
    if t.dtype == np.uint32:
        return np.ascontiguousarray(t), 0
    if t.dtype == np.uint64:
        return np.ascontiguousarray(t), 1
    if t.dtype.kind in "iu":
        if t.size and (t.min() < 0):
            raise OverflowError("token hash values must be non-negative")  # numpy raises the same on uint64 cast
        if t.size == 0 or int(t.max()) < (1 << 32):
            return np.ascontiguousarray(t.astype(np.uint32)), 0
        return np.ascontiguousarray(t.astype(np.uint64)), 1
    raise TypeError("token hashes must be an integer array, got dtype %s" % t.dtype)


def pack_docs(docs: Sequence[Sequence[int]]) -> Tuple[np.ndarray, np.ndarray]:
    """List of integer-hash lists -> (tokens, offsets) CSR.  u32 storage when every hash fits 32 bits."""
    lens = np.fromiter((len(d) for d in docs), dtype=np.int64, count=len(docs))
    offsets = np.zeros(len(docs) + 1, dtype=np.int64)
    np.cumsum(lens, out=offsets[1:])
    flat = [h for d in docs for h in d]
    if not flat:
        return np.zeros(0, dtype=np.uint32), offsets
    try:
        tok = np.array(flat, dtype=np.uint64)  # same cast (and OverflowError on negatives) as minhash.py:294
    except OverflowError:
        raise
    if int(tok.max()) < (1 << 32):
        tok = tok.astype(np.uint32)
    return tok, offsets


def bulk_signatures(tokens, offsets, permutations: np.ndarray, init: Optional[np.ndarray] = None,
                    out_u64: bool = False, kernel: str = "auto", device: int = 0,
                    out: Optional[np.ndarray] = None) -> np.ndarray:
    """HOST arrays in, HOST [N, K] signature matrix out (u32, or u64 = the reference's dtype).

    One call to ``dsk_minhash_bulk_host``: slices of documents are pipelined
    H2D -> kernel -> D2H on internal streams.  ``init`` is ``None`` (empty
    state), one row (broadcast) or an [N, K] matrix of running signatures.
    """
    tok, is64 = _as_tokens(tokens)
    off = np.ascontiguousarray(offsets, dtype=np.int64)
    if off.ndim != 1 or off.size < 1:
        raise ValueError("offsets must be a 1-D array of length n_docs + 1")
    n = off.size - 1
    # the library indexes the token array with these offsets absolutely: the ends are checked here, every interior
    # offset (non-decreasing, hence inside [off[0], off[-1]]) by dsk_minhash_bulk_host in the pass that reads them anyway
    if n and (int(off[0]) < 0 or int(off[-1]) > tok.size or int(off[-1]) < int(off[0])):
        raise ValueError("offsets must be non-decreasing and lie inside the token array")
    nv.require_device(device)
    h = nv.perm_handle(permutations, device)
    k = h.num_perm
    dt = np.uint64 if out_u64 else np.uint32
    if out is None:
        out = np.empty((n, k), dtype=dt)
    elif out.dtype != dt or out.shape != (n, k) or not out.flags.c_contiguous:
        raise ValueError("out must be a C-contiguous [n_docs, num_perm] array of dtype %s" % dt)
    init_p, init_stride, init64 = None, 0, 0
    if init is not None:
        init = np.ascontiguousarray(init)
        if init.dtype not in (np.uint32, np.uint64):
            init = init.astype(np.uint64)
        init64 = int(init.dtype == np.uint64)
        if init.ndim == 1:
            if init.shape[0] != k:
                raise ValueError("init row length mismatch")
            init_stride = 0
        else:
            if init.shape != (n, k):
                raise ValueError("init must be [num_perm] or [n_docs, num_perm]")
            init_stride = k
        init_p = init.ctypes.data
    if n == 0:
        return out
    nv.check(nv.load().dsk_minhash_bulk_host(h.handle, tok.ctypes.data if tok.size else None, is64, off.ctypes.data,
                                              n, init_p, init_stride, init64, out.ctypes.data, int(out_u64),
                                              KERNELS[kernel]))
    return out


def release_host_pipeline(device: int = -1) -> None:
    """Free the streams and staging buffers ``bulk_signatures`` keeps per device between calls
    (``dsk_release_host_pipeline``; -1 = all devices).  The next call re-creates them."""
    nv.check(nv.load().dsk_release_host_pipeline(int(device)))


def bulk_signatures_device(d_tokens, d_offsets, n_tokens: int, permutations: np.ndarray, d_out=None, d_init=None,
                           init_stride: int = 0, kernel: str = "auto", stream: Optional[int] = None, workspace=None):
    """DEVICE buffers in (torch CUDA tensors: uint32/int32 or uint64/int64 tokens, int64 offsets),
    DEVICE [N, K] signature tensor out.  Asynchronous on ``stream`` (default: torch's current stream)."""
    import torch
    dev = d_tokens.device.index if d_tokens.is_cuda else None
    if dev is None or not d_offsets.is_cuda:
        raise ValueError("bulk_signatures_device needs CUDA tensors")
    is64 = int(d_tokens.element_size() == 8)
    n = d_offsets.numel() - 1
    h = nv.perm_handle(permutations, dev)
    k = h.num_perm
    if d_out is None:
        d_out = torch.empty((n, k), dtype=torch.int32, device=d_tokens.device)
    out64 = int(d_out.element_size() == 8)
    if stream is None:
        stream = torch.cuda.current_stream(d_tokens.device).cuda_stream
    with torch.cuda.device(dev):
        # long documents (> 16384 tokens) are cut into pieces on the device: that needs a piece table, sized by the
        # library and owned here (a torch tensor, kept alive until the launch was enqueued on a torch-known stream)
        ws_bytes = int(nv.load().dsk_minhash_bulk_workspace_size(n, n_tokens)) if workspace is None else 0
        if workspace is None and ws_bytes:
            workspace = torch.empty((ws_bytes,), dtype=torch.uint8, device=d_tokens.device)
        nv.check(nv.load().dsk_minhash_bulk_ws(h.handle, d_tokens.data_ptr() if n_tokens else None, is64,
                                                d_offsets.data_ptr(), n, n_tokens,
                                                d_init.data_ptr() if d_init is not None else None, init_stride,
                                                int(d_init.element_size() == 8) if d_init is not None else 0,
                                                d_out.data_ptr(), out64, KERNELS[kernel],
                                                workspace.data_ptr() if workspace is not None else None,
                                                workspace.numel() * workspace.element_size() if workspace is not None else 0,
                                                stream))
        if workspace is not None:
            workspace.record_stream(torch.cuda.ExternalStream(stream) if stream else torch.cuda.default_stream(d_tokens.device))
    return d_out


_stage: dict = {}
_stage_lock = threading.Lock()   # the pinned staging buffers are shared: one bulk_signatures_sha1 call at a time


def _pinned_stage(name: str, nbytes: int):
    """Grow-only pinned host staging buffer (uint8 tensor), one per purpose per process."""
    import torch
    t = _stage.get(name)
    if t is None or t.numel() < nbytes:
        t = torch.empty((max(int(nbytes) * 2, 1 << 24) + 7) // 8 * 8, dtype=torch.uint8, pin_memory=True)
        _stage[name] = t
    return t


def _sha1_blob_device(blob, boff: np.ndarray, n: int, device: int = 0, out_u64: bool = False):
    """SHA1-32/64 of ``n`` byte strings stored back to back in ``blob`` (``boff`` = n+1 byte offsets)."""
    import torch
    nv.require_device(device)
    dev = torch.device("cuda", device)
    if not isinstance(blob, bytearray):
        blob = bytearray(blob)  # torch.frombuffer wants a writable buffer
    d_bytes = (torch.frombuffer(blob, dtype=torch.uint8) if len(blob) else torch.zeros(1, dtype=torch.uint8)).to(dev)
    d_boff = torch.from_numpy(np.ascontiguousarray(boff, dtype=np.int64)).to(dev)
    out = torch.empty((max(n, 4),), dtype=torch.int64 if out_u64 else torch.int32, device=dev)
    with torch.cuda.device(device):
        nv.check(nv.load().dsk_sha1_tokens(d_bytes.data_ptr(), d_boff.data_ptr(), n, out.data_ptr(), int(out_u64),
                                            torch.cuda.current_stream(dev).cuda_stream))
    return out[:n]     # (the allocation is padded to 4 elements; the result is not)


def sha1_hash_tokens_device(flat_tokens, device: int = 0, out_u64: bool = False):
    """Device-side ``sha1_hash32`` / ``sha1_hash64`` (datasketch/hashfunc.py:5-28) of a flat list of
    byte strings -> CUDA tensor of hashes (int32 / int64 storage of the unsigned values)."""
    n = len(flat_tokens)
    lens = np.fromiter(map(len, flat_tokens), dtype=np.int64, count=n)
    blob = b"".join(flat_tokens)  # TypeError for non-bytes tokens, like hashlib would raise
    boff = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(lens, out=boff[1:])
    if int(boff[-1]) != len(blob):
        raise TypeError("tokens must be byte strings (item size 1)")
    return _sha1_blob_device(blob, boff, n, device, out_u64)


def hash_tokens_device(flat_tokens, kind: int, seed: int = 0, device: int = 0):
    """Device-side XXH32 (``kind`` 1) / MurmurHash3 x86_32 (``kind`` 2) of a flat list of byte strings -> CUDA int32
    tensor holding the unsigned 32-bit hashes (``hashfunc.xxh32_hash32`` / ``murmur3_hash32`` per token)."""
    import torch
    nv.require_device(device)
    n = len(flat_tokens)
    lens = np.fromiter(map(len, flat_tokens), dtype=np.int64, count=n)
    blob = bytearray().join(flat_tokens)
    boff = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(lens, out=boff[1:])
    if int(boff[-1]) != len(blob):
        raise TypeError("tokens must be byte strings (item size 1)")
    dev = torch.device("cuda", device)
    d_bytes = (torch.frombuffer(blob, dtype=torch.uint8) if len(blob) else torch.zeros(1, dtype=torch.uint8)).to(dev)
    d_boff = torch.from_numpy(boff).to(dev)
    out = torch.empty((max(n, 1),), dtype=torch.int32, device=dev)
    with torch.cuda.device(device):
        nv.check(nv.load().dsk_hash_tokens(d_bytes.data_ptr(), d_boff.data_ptr(), n, int(kind), int(seed) & 0xFFFFFFFF,
                                            out.data_ptr(), torch.cuda.current_stream(dev).cuda_stream))
    return out[:n]


def bulk_signatures_sha1(docs: Sequence[Sequence[bytes]], permutations: np.ndarray, init: Optional[np.ndarray] = None,
                         device: int = 0, hash_kind: int = 0) -> np.ndarray:
    """``MinHash.bulk`` for byte tokens under a hash function the library has on device: the token hash
    (``hash_kind`` 0 = SHA1-32 via ``dsk_sha1_tokens``, the reference's default; ``DSK_HASH_XXH32`` = 1 /
    ``DSK_HASH_MURMUR3_32`` = 2 via ``dsk_hash_tokens``) feeds the signature kernel without the hashes ever
    visiting the host.  Returns the [N, K] uint64 matrix (the reference's dtype).

    Host packing is one C-level pass per document (``bytes.join`` while the document's tokens are
    cache-hot) plus one ``map(len, ...)`` over all tokens; a token that is not bytes-like raises
    TypeError before anything reaches the device (the caller then takes the per-token hashfunc
    route, which raises what the reference raises)."""
    with _stage_lock:
        import itertools
        import torch
        nv.require_device(device)
        dev = torch.device("cuda", device)
        n = len(docs)
        doc_lens = np.fromiter(map(len, docs), dtype=np.int64, count=n)
        off = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(doc_lens, out=off[1:])
        n_tok = int(off[-1])
        lens = np.fromiter(map(len, itertools.chain.from_iterable(docs)), dtype=np.int64, count=n_tok)
        # grow-only pinned staging (no fresh-page faults per call, full-rate H2D); safe to reuse because this
        # function returns only after the device->host copy of the result, which is ordered after both uploads
        h_boff = _pinned_stage("boff", (n_tok + 1) * 8).view(torch.int64).numpy()
        h_boff[0] = 0
        np.cumsum(lens, out=h_boff[1:n_tok + 1])
        total = int(h_boff[n_tok])
        h_blob = _pinned_stage("blob", total)
        mv = memoryview(h_blob.numpy())
        pos = 0
        for d in docs:
            piece = b"".join(d)  # TypeError for a token that is not bytes-like
            end = pos + len(piece)
            if end > total:
                break
            mv[pos:end] = piece
            pos = end
        else:
            end = pos
        if end != total:  # e.g. array('I') tokens: len() counts items, not bytes
            raise TypeError("tokens must be byte strings (item size 1)")
        d_bytes = torch.empty((max(total, 1),), dtype=torch.uint8, device=dev)
        d_bytes[:total].copy_(h_blob[:total], non_blocking=True)
        d_boff = torch.empty((n_tok + 1,), dtype=torch.int64, device=dev)
        d_boff.copy_(_pinned_stage("boff", 0).view(torch.int64)[:n_tok + 1], non_blocking=True)
        d_hash = torch.empty((max(n_tok, 4),), dtype=torch.int32, device=dev)
        with torch.cuda.device(device):
            st = torch.cuda.current_stream(dev).cuda_stream
            if hash_kind == 0:
                nv.check(nv.load().dsk_sha1_tokens(d_bytes.data_ptr(), d_boff.data_ptr(), n_tok, d_hash.data_ptr(), 0, st))
            else:
                nv.check(nv.load().dsk_hash_tokens(d_bytes.data_ptr(), d_boff.data_ptr(), n_tok, int(hash_kind), 0,
                                                    d_hash.data_ptr(), st))
        d_off = torch.from_numpy(off).to(dev)
        d_out = torch.empty((n, permutations.shape[1]), dtype=torch.int64, device=dev)
        d_init, stride = None, 0
        if init is not None:
            init = np.ascontiguousarray(init, dtype=np.uint64)
            d_init = torch.from_numpy(init.view(np.int64)).to(dev)
            stride = 0 if init.ndim == 1 else init.shape[1]
        if n:
            bulk_signatures_device(d_hash, d_off, n_tok, permutations, d_out=d_out, d_init=d_init, init_stride=stride)
        return d_out.cpu().numpy().view(np.uint64)
