"""The whole library on the CPU: ``tests/emu/emu_capi.cpp`` compiles every kernel file AND the C-ABI layer
(``dsk_api.cu``) with g++ on the emulation shim and exports the same C-ABI as ``libdsk_b200.so`` ("device" memory is host
memory, one emulated device with two SMs, streams synchronous).  Test infrastructure only -- the product never loads it.

Two uses:
* the C-ABI itself through ctypes: versions / device info, the permutation analysis, argument validation and error
  codes, the host pipeline of ``dsk_minhash_bulk_host`` with tiny slices (many slices, long documents cut into pieces
  and min-merged, running-state merge);
* swapped in for the real library inside a fixture, the Python layer's host-buffer paths: the parity tests of
  ``tests/test_minhash_gpu.py`` that need no torch CUDA tensor run unchanged (lazy update queue, bulk, estimators, the
  API fuzz against the oracle), as do LeanMinHash and the dict-bucket MinHashLSH on top of them.
"""
import ctypes
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest

from oracle import oracle_clib as oc
from oracle import oracle_np as o

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tests", "emu")
sys.path.insert(0, os.path.join(ROOT, "tests"))

pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")


@pytest.fixture()
def dsk_on_emu(emu_lib, monkeypatch):
    """datasketch_b200 with the emulated library swapped in for libdsk_b200.so (host-buffer paths only)."""
    import datasketch_b200 as dsk
    from datasketch_b200 import _native as nv
    nv._perm_cache.clear()
    monkeypatch.setattr(nv, "_lib", emu_lib)
    yield dsk
    nv._perm_cache.clear()                    # handles created on the emulated library die on it


# ---- the C-ABI directly -----------------------------------------------------------------------------------------------
def test_version_device_and_error_plumbing(emu_lib):
    lib = emu_lib
    assert lib.dsk_version() > 0 and lib.dsk_device_count() == 1
    sm, ma, mi, mem = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_size_t()
    assert lib.dsk_device_info(0, ctypes.byref(sm), ctypes.byref(ma), ctypes.byref(mi), ctypes.byref(mem)) == 0
    assert (sm.value, ma.value) == (2, 10)
    assert lib.dsk_device_info(3, ctypes.byref(sm), ctypes.byref(ma), ctypes.byref(mi), ctypes.byref(mem)) != 0
    assert b"device" in lib.dsk_last_error().lower()
    h = ctypes.c_void_p()
    a = np.arange(1, 5, dtype=np.uint64)
    assert lib.dsk_perm_create(a.ctypes.data, None, 4, 0, ctypes.byref(h)) != 0          # null b
    assert lib.dsk_perm_create(a.ctypes.data, a.ctypes.data, 0, 0, ctypes.byref(h)) != 0  # num_perm = 0
    assert lib.dsk_minhash_bulk_host(None, None, 0, None, 0, None, 0, 0, None, 0, 0) != 0
    assert lib.dsk_hash_tokens(None, None, 1, 7, 0, None, None) != 0 and b"kind" in lib.dsk_last_error()


def test_host_pipeline_tiny_slices_long_documents_and_init(emu_lib, monkeypatch):
    """dsk_minhash_bulk_host with 4096-token slices: dozens of slices over three pipeline slots, documents far above the
    piece length cut into pieces and min-merged by seg_min_kernel, running-state merge in both dtypes."""
    lib = emu_lib
    monkeypatch.setenv("DSK_SLICE_TOKENS", "4096")
    rs = np.random.RandomState(21)
    k = 100
    P = o.init_permutations(k, 4)
    a, b = np.ascontiguousarray(P[0]), np.ascontiguousarray(P[1])
    h = ctypes.c_void_p()
    assert lib.dsk_perm_create(a.ctypes.data, b.ctypes.data, k, 0, ctypes.byref(h)) == 0
    n, bad, dev = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    assert lib.dsk_perm_info(h, ctypes.byref(n), ctypes.byref(bad), ctypes.byref(dev)) == 0 and (n.value, bad.value) == (k, 0)
    lens = [700, 0, 17, 4096, 4097, 30_000, 3, 0, 9_000, 129] + list(rs.randint(0, 300, size=150))
    off = np.zeros(len(lens) + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    tok = rs.randint(0, 1 << 32, size=int(off[-1]), dtype=np.uint64).astype(np.uint32)
    want = oc.minhash_bulk_u32tok(tok, off, P)
    for flags in (0, 1, 2, 3):                                             # auto, two-phase, direct, exact
        out = np.zeros((len(lens), k), dtype=np.uint32)
        assert lib.dsk_minhash_bulk_host(h, tok.ctypes.data, 0, off.ctypes.data, len(lens), None, 0, 0, out.ctypes.data, 0,
                                         flags) == 0, lib.dsk_last_error()
        assert np.array_equal(out, want), flags
    init = rs.randint(0, 1 << 32, size=(len(lens), k), dtype=np.uint64)
    out64 = np.zeros((len(lens), k), dtype=np.uint64)
    assert lib.dsk_minhash_bulk_host(h, tok.ctypes.data, 0, off.ctypes.data, len(lens), init.ctypes.data, k, 1,
                                     out64.ctypes.data, 1, 0) == 0
    assert np.array_equal(out64, np.minimum(init, want.astype(np.uint64)))
    row = init[3].astype(np.uint32)
    out = np.zeros((len(lens), k), dtype=np.uint32)
    assert lib.dsk_minhash_bulk_host(h, tok.ctypes.data, 0, off.ctypes.data, len(lens), row.ctypes.data, 0, 0,
                                     out.ctypes.data, 0, 0) == 0
    assert np.array_equal(out, np.minimum(row[None, :], want))
    tok64 = tok.astype(np.uint64) << np.uint64(7)
    out = np.zeros((len(lens), k), dtype=np.uint32)
    assert lib.dsk_minhash_bulk_host(h, tok64.ctypes.data, 1, off.ctypes.data, len(lens), None, 0, 0, out.ctypes.data, 0, 0) == 0
    assert np.array_equal(out, oc.minhash_bulk_u64tok(tok64, off, P))
    assert lib.dsk_minhash_bulk_host(h, tok64.ctypes.data, 1, off.ctypes.data, len(lens), None, 0, 0, out.ctypes.data, 0, 1) != 0
    lib.dsk_perm_destroy(h)


# ---- the Python layer's host-buffer paths on the emulated library ----------------------------------------------------
def test_minhash_parity_tests_run_on_the_emulated_library(dsk_on_emu, golden):
    import test_minhash_gpu as t
    dsk = dsk_on_emu
    t.test_reference_absolute_golden(dsk)
    for kernel in t.KERNELS:
        t.test_c1_bulk_golden(dsk, golden, kernel)
        t.test_long_and_small_docs(dsk, golden, kernel)
        t.test_tail_and_alignment_shapes(dsk, kernel)
        for k in (4, 33, 100, 128, 256):
            t.test_ragged_golden(dsk, golden, k, kernel)
    t.test_u64_tokens_golden(dsk, golden)
    t.test_sha1_update_batch_like_reference_gpu_test(dsk, golden)
    t.test_update_equals_update_batch_equals_bulk(dsk)
    t.test_running_state_merge_and_estimators(dsk, golden)
    t.test_init_matrix_and_broadcast(dsk)
    t.test_duplicates_and_near_ties_take_slow_path(dsk)
    t.test_unsafe_permutations_route_to_exact(dsk)
    t.test_seeded_permutations_are_safe(dsk)
    for kernel in ("auto", "exact"):
        t.test_long_documents_are_split_and_merged(dsk, kernel)      # a 2.5M-token document cut into pieces


def test_api_fuzz_runs_on_the_emulated_library(dsk_on_emu):
    import test_minhash_gpu as t
    t.test_api_fuzz_against_oracle(dsk_on_emu)


def test_device_entry_points_and_fused_gather(emu_lib):
    """dsk_minhash_bulk (device-pointer entry: alignment check, init) and dsk_minhash_bulk_gather: every signature row
    must land at row_offset + i of EVERY peer matrix (the peers are plain host buffers here)."""
    lib = emu_lib
    rs = np.random.RandomState(3)
    k, n = 128, 90
    P = o.init_permutations(k, 1)
    a, b = np.ascontiguousarray(P[0]), np.ascontiguousarray(P[1])
    h = ctypes.c_void_p()
    assert lib.dsk_perm_create(a.ctypes.data, b.ctypes.data, k, 0, ctypes.byref(h)) == 0
    lens = rs.randint(0, 120, size=n)
    off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    tok = rs.randint(0, 1 << 32, size=int(off[-1]) + 4, dtype=np.uint64).astype(np.uint32)
    want = oc.minhash_bulk_u32tok(tok[:off[-1]], off, P)
    out = np.zeros((n, k), dtype=np.uint32)
    assert lib.dsk_minhash_bulk(h, tok.ctypes.data, 0, off.ctypes.data, n, int(off[-1]), None, 0, 0, out.ctypes.data, 0, 0,
                                None) == 0, lib.dsk_last_error()
    assert np.array_equal(out, want)
    assert lib.dsk_minhash_bulk(h, tok.ctypes.data + 4, 0, off.ctypes.data, n, int(off[-1]) - 1, None, 0, 0, out.ctypes.data,
                                0, 0, None) != 0                      # token pointer not 16-byte aligned
    assert b"align" in lib.dsk_last_error().lower()
    n_total, row0 = n + 37, 20
    peers = [np.full((n_total, k), 0xABCDEF01, dtype=np.uint32) for _ in range(3)]
    ptrs = (ctypes.c_void_p * 3)(*[p.ctypes.data for p in peers])
    assert lib.dsk_minhash_bulk_gather(h, tok.ctypes.data, 0, off.ctypes.data, n, int(off[-1]), ptrs, 3, row0, 0, 0,
                                       None) == 0, lib.dsk_last_error()
    for p in peers:
        assert np.array_equal(p[row0:row0 + n], want)
        assert (p[:row0] == 0xABCDEF01).all() and (p[row0 + n:] == 0xABCDEF01).all()   # nothing else touched
    assert lib.dsk_minhash_bulk_gather(h, tok.ctypes.data, 0, off.ctypes.data, n, int(off[-1]), ptrs, 9, row0, 0, 0, None) != 0
    lib.dsk_perm_destroy(h)
