"""TEST INFRASTRUCTURE ONLY -- ctypes loader for the plain-C oracle (oracle/oracle_c.c).

Only tests/, __graft_entry__ and bench.py's cpu_baseline legs may import this.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "oracle_c.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = ctypes.CDLL(_SO)
    return _lib


def _p(x):
    return x.ctypes.data_as(ctypes.c_void_p)


def minhash_bulk_u32tok(tokens: np.ndarray, offsets: np.ndarray, perms: np.ndarray,
                        d0: int = 0, d1: int | None = None, out: np.ndarray | None = None) -> np.ndarray:
    tokens = np.ascontiguousarray(tokens, dtype=np.uint32)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    a = np.ascontiguousarray(perms[0], dtype=np.uint64)
    b = np.ascontiguousarray(perms[1], dtype=np.uint64)
    n = len(offsets) - 1
    k = len(a)
    if d1 is None:
        d1 = n
    if out is None:
        out = np.empty((n, k), dtype=np.uint32)
    lib().oracle_minhash_bulk_u32tok(_p(tokens), _p(offsets), ctypes.c_int64(d0), ctypes.c_int64(d1),
                                     _p(a), _p(b), ctypes.c_int(k), _p(out))
    return out


def minhash_bulk_u64tok(tokens: np.ndarray, offsets: np.ndarray, perms: np.ndarray) -> np.ndarray:
    tokens = np.ascontiguousarray(tokens, dtype=np.uint64)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    a = np.ascontiguousarray(perms[0], dtype=np.uint64)
    b = np.ascontiguousarray(perms[1], dtype=np.uint64)
    n = len(offsets) - 1
    k = len(a)
    out = np.empty((n, k), dtype=np.uint64)
    lib().oracle_minhash_bulk_u64tok(_p(tokens), _p(offsets), ctypes.c_int64(n), _p(a), _p(b),
                                     ctypes.c_int(k), _p(out))
    return out


def jaccard_pairs_u32(sig: np.ndarray, ia: np.ndarray, ib: np.ndarray) -> np.ndarray:
    sig = np.ascontiguousarray(sig, dtype=np.uint32)
    ia = np.ascontiguousarray(ia, dtype=np.int64)
    ib = np.ascontiguousarray(ib, dtype=np.int64)
    out = np.empty(len(ia), dtype=np.int32)
    lib().oracle_jaccard_pairs_u32(_p(sig), _p(ia), _p(ib), ctypes.c_int64(len(ia)),
                                   ctypes.c_int(sig.shape[1]), _p(out))
    return out


def lean_pack_le(sig: np.ndarray, seed: int) -> np.ndarray:
    sig = np.ascontiguousarray(sig, dtype=np.uint32)
    n, k = sig.shape
    out = np.empty((n, 12 + 4 * k), dtype=np.uint8)
    lib().oracle_lean_pack_le(_p(sig), ctypes.c_int64(n), ctypes.c_int(k), ctypes.c_int64(seed), _p(out))
    return out


def band_keys_be(sig: np.ndarray, b: int, r: int) -> np.ndarray:
    sig = np.ascontiguousarray(sig, dtype=np.uint32)
    n, k = sig.shape
    out = np.empty((n, b, 8 * r), dtype=np.uint8)
    lib().oracle_band_keys_be(_p(sig), ctypes.c_int64(n), ctypes.c_int(k), ctypes.c_int(b), ctypes.c_int(r), _p(out))
    return out
