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
    # 64-bit tokens: AUTO / TWO_PHASE = the general variant of the signature kernel, EXACT = round 1's kernel; DIRECT refuses
    for flags in (1, 3):
        out = np.zeros((len(lens), k), dtype=np.uint32)
        assert lib.dsk_minhash_bulk_host(h, tok64.ctypes.data, 1, off.ctypes.data, len(lens), None, 0, 0, out.ctypes.data, 0, flags) == 0
        assert np.array_equal(out, oc.minhash_bulk_u64tok(tok64, off, P)), flags
    assert lib.dsk_minhash_bulk_host(h, tok64.ctypes.data, 1, off.ctypes.data, len(lens), None, 0, 0, out.ctypes.data, 0, 2) != 0
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
    t.test_unsafe_permutations_take_the_general_variant(dsk)
    t.test_seeded_permutations_are_safe(dsk)
    for kernel in ("auto", "exact"):
        t.test_long_documents_are_split_and_merged(dsk, kernel)      # a 2.5M-token document cut into pieces


def test_general_variant_gpu_tests_run_on_the_emulated_library(dsk_on_emu):
    """The host-buffer GPU tests of the general variants (u32 tokens + unsafe permutations, u64 tokens) through the Python
    layer and the C-ABI on the emulated library: routing (dsk_perm_create's analysis -> BulkParams::gen) included."""
    import test_signature_kernel_gpu as ts
    ts.test_general_variants_structured_tokens_wraps_and_extreme_parameters(False)
    ts.test_general_variants_structured_tokens_wraps_and_extreme_parameters(True)


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


def test_remaining_entry_points_through_the_c_abi(emu_lib, golden):
    """One pass over every other C-ABI entry with host buffers standing in for device buffers: results against the oracle
    / fixtures, and the argument checks each entry makes."""
    lib = emu_lib
    P64, I64 = ctypes.c_void_p, ctypes.c_int64
    rs = np.random.RandomState(10)
    ptr = lambda x: P64(x.ctypes.data)                                              # noqa: E731
    # permutation analysis (host-only entry)
    P = o.init_permutations(64, 2)
    bad = np.ones(64, dtype=np.uint8)
    assert lib.dsk_perm_analyze(ptr(np.ascontiguousarray(P[0])), ptr(np.ascontiguousarray(P[1])), 64, ptr(bad)) == 0
    assert not bad.any()
    # merge, lean codec, band keys / fingerprints, b-bit
    k, n = 128, 50
    sig = rs.randint(0, 1 << 32, size=(n, k), dtype=np.uint64).astype(np.uint32)
    other = rs.randint(0, 1 << 32, size=(n, k), dtype=np.uint64).astype(np.uint32)
    out = np.zeros_like(sig)
    assert lib.dsk_sig_merge_min(ptr(sig), ptr(other), I64(n * k), ptr(out), None) == 0
    assert np.array_equal(out, np.minimum(sig, other))
    rec = np.zeros((n, 12 + 4 * k), dtype=np.uint8)
    assert lib.dsk_lean_pack(ptr(sig), 0, I64(n), k, I64(-5), 0, ptr(rec), None) == 0
    assert np.array_equal(rec, oc.lean_pack_le(sig, -5))
    back = np.zeros_like(sig)
    st = np.zeros(1, dtype=np.int32)
    assert lib.dsk_lean_unpack(ptr(rec), I64(n), k, I64(-5), 0, ptr(back), 0, ptr(st), None) == 0
    assert st[0] == 0 and np.array_equal(back, sig)
    keys = np.zeros((n, 9, 8 * 13), dtype=np.uint8)
    assert lib.dsk_band_keys(ptr(sig), I64(n), k, 9, 13, ptr(keys), None) == 0
    assert np.array_equal(keys, oc.band_keys_be(sig, 9, 13))
    assert lib.dsk_band_keys(ptr(sig), I64(n), k, 20, 13, ptr(keys), None) != 0           # b * r > num_perm
    fp = np.zeros((n, 9), dtype=np.uint64)
    assert lib.dsk_band_fingerprints(ptr(sig), I64(n), k, 9, 13, ptr(fp), None) == 0 and len(np.unique(fp)) == fp.size
    blocks = np.zeros((n, k // 16), dtype=np.uint64)
    assert lib.dsk_bbit_pack(ptr(sig), I64(n), k, 3, ptr(blocks), None) == 0               # b = 3 -> 4-bit slots
    unp = np.zeros_like(sig)
    assert lib.dsk_bbit_unpack(ptr(blocks), I64(n), k, 3, ptr(unp), None) == 0 and np.array_equal(unp, sig & 7)
    assert lib.dsk_bbit_pack(ptr(sig), I64(n), k, 40, ptr(blocks), None) != 0
    # Jaccard
    low = rs.randint(0, 3, size=(200, 64)).astype(np.uint32)
    ia, ib = rs.randint(0, 200, size=300).astype(np.int64), rs.randint(0, 200, size=300).astype(np.int64)
    cnt = np.zeros(300, dtype=np.int32)
    assert lib.dsk_jaccard_pairs(ptr(low), I64(200), 64, ptr(ia), ptr(ib), I64(300), ptr(cnt), None) == 0
    assert np.array_equal(cnt, (low[ia] == low[ib]).sum(axis=1))
    tc, ti = np.zeros((200, 4), dtype=np.int32), np.zeros((200, 4), dtype=np.int64)
    assert lib.dsk_jaccard_topk(ptr(low), I64(200), ptr(low), I64(200), 64, 4, I64(0), ptr(tc), ptr(ti), None) == 0
    c0 = (low == low[0][None, :]).sum(axis=1)
    order = np.lexsort((np.arange(200), -c0))
    order = order[order != 0][:4]
    assert np.array_equal(ti[0], order) and np.array_equal(tc[0], c0[order])
    assert lib.dsk_jaccard_topk(ptr(low), I64(200), ptr(low), I64(200), 64, 99, I64(0), ptr(tc), ptr(ti), None) != 0
    # LSH index handle
    ix = ctypes.c_void_p()
    assert lib.dsk_lsh_create(64, 8, 8, I64(200), 0, ctypes.byref(ix)) == 0
    assert lib.dsk_lsh_create(64, 9, 8, I64(200), 0, ctypes.byref(ctypes.c_void_p())) != 0     # b * r > num_perm
    assert lib.dsk_lsh_insert(ix, ptr(low), I64(200), None) == 0
    assert lib.dsk_lsh_insert(ix, ptr(low), I64(1), None) != 0 and b"capacity" in lib.dsk_last_error()
    nd, cap = ctypes.c_int64(), ctypes.c_int64()
    assert lib.dsk_lsh_size(ix, ctypes.byref(nd), ctypes.byref(cap)) == 0 and (nd.value, cap.value) == (200, 200)
    counts = np.zeros(201, dtype=np.int64)
    assert lib.dsk_lsh_query_count(ix, ptr(low), I64(200), ptr(counts), None) == 0
    pref, scratch = np.zeros(201, dtype=np.int64), np.zeros(8, dtype=np.int64)
    assert lib.dsk_exclusive_scan(ptr(counts), I64(200), ptr(pref), ptr(scratch), None) == 0
    idx = np.zeros(max(int(pref[-1]), 1), dtype=np.int32)
    assert lib.dsk_lsh_query_fill(ix, ptr(low), I64(200), ptr(pref), ptr(idx), None) == 0
    ref = o.DictLSH(64, 8, 8)
    for i, row in enumerate(low):
        ref.insert(i, row.astype(np.uint64))
    for j in (0, 57, 199):
        assert sorted(idx[pref[j]:pref[j + 1]].tolist()) == sorted(ref.query(low[j].astype(np.uint64)))
    lib.dsk_lsh_destroy(ix)
    # token hashes and Weighted MinHash
    g = golden("hashes")
    blob, off = np.concatenate([g["blob"], np.zeros(8, np.uint8)]), np.ascontiguousarray(g["off"])
    nt = len(off) - 1
    h32 = np.zeros(nt, dtype=np.uint32)
    assert lib.dsk_hash_tokens(ptr(blob), ptr(off), I64(nt), 1, 1, ptr(h32), None) == 0
    assert np.array_equal(h32, g["xxh32_seed1"])
    assert lib.dsk_sha1_tokens(ptr(blob), ptr(off), I64(nt), ptr(h32), 0, None) == 0
    assert h32[3] == o.sha1_hash32(bytes(blob[off[3]:off[4]]))
    w = golden("wmh")
    dim, ss, seed = (int(x) for x in w["tiny_cfg"])
    rs_, ln_cs, betas = (np.ascontiguousarray(x) for x in o.wmh_params(dim, ss, seed))
    gen = ctypes.c_void_p()
    assert lib.dsk_wmh_create(ptr(rs_), ptr(ln_cs), ptr(betas), ss, dim, 0, ctypes.byref(gen)) == 0
    V = np.ascontiguousarray(w["tiny_v"], dtype=np.float32)
    kt = np.zeros((len(V), ss, 2), dtype=np.int64)
    stv = np.zeros(len(V), dtype=np.int32)
    assert lib.dsk_wmh_minhash(gen, ptr(V), I64(len(V)), ptr(kt), ptr(stv), 0, None) == 0
    assert not stv.any() and np.array_equal(kt, w["tiny_out"])
    assert lib.dsk_wmh_minhash(gen, ptr(V), I64(len(V)), ptr(kt), ptr(stv), 5, None) != 0      # unknown flags
    lib.dsk_wmh_destroy(gen)


def test_release_host_pipeline_and_counter_leases(emu_lib, monkeypatch):
    """dsk_release_host_pipeline frees the per-device staging (the next host call re-creates it and still gives the
    oracle's rows); more launches than counter sets (64) through one permutation handle stay correct."""
    lib = emu_lib
    monkeypatch.setenv("DSK_SLICE_TOKENS", "4096")
    rs = np.random.RandomState(2)
    lens = rs.randint(0, 120, size=40)
    off = np.zeros(41, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    tok = rs.randint(0, 1 << 32, size=int(off[-1]), dtype=np.uint64).astype(np.uint32)
    perms = o.init_permutations(64, 1)
    a, b = np.ascontiguousarray(perms[0]), np.ascontiguousarray(perms[1])
    h = ctypes.c_void_p()
    assert lib.dsk_perm_create(a.ctypes.data, b.ctypes.data, 64, 0, ctypes.byref(h)) == 0
    want = oc.minhash_bulk_u32tok(tok, off, perms)
    out = np.zeros((40, 64), dtype=np.uint32)
    for round_ in range(3):
        out[:] = 0
        assert lib.dsk_minhash_bulk_host(h, tok.ctypes.data, 0, off.ctypes.data, 40, None, 0, 0, out.ctypes.data, 0, 0) == 0
        assert np.array_equal(out, want)
        assert lib.dsk_release_host_pipeline(0 if round_ else -1) == 0
    assert lib.dsk_release_host_pipeline(99) != 0
    tok16 = np.zeros(len(tok) + 4, dtype=np.uint32)
    tok16[:len(tok)] = tok
    for i in range(70):                                   # > 64 launches: counter sets are leased round-robin
        out[:] = 0
        assert lib.dsk_minhash_bulk(h, tok16.ctypes.data, 0, off.ctypes.data, 40, len(tok), None, 0, 0, out.ctypes.data, 0, 0, None) == 0
        if i % 23 == 0:
            assert np.array_equal(out, want)
    assert np.array_equal(out, want)
    lib.dsk_perm_destroy(h)


def test_minhash_bulk_ws_cuts_long_documents(emu_lib):
    """dsk_minhash_bulk_ws with the library-sized workspace: a 60 000-token and a 20 000-token document between short ones
    (the reference's GPU benchmark shape is one 50 000-token update_batch); workspace size / alignment validation."""
    lib = emu_lib
    rs = np.random.RandomState(9)
    lens = np.array([50, 60_000, 0, 300, 20_000, 4096, 9], dtype=np.int64)
    off = np.zeros(len(lens) + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    nt = int(off[-1])
    tok = np.zeros(nt + 4, dtype=np.uint32)
    tok[:nt] = rs.randint(0, 1 << 32, size=nt, dtype=np.uint64).astype(np.uint32)
    perms = o.init_permutations(32, 1)
    a, b = np.ascontiguousarray(perms[0]), np.ascontiguousarray(perms[1])
    h = ctypes.c_void_p()
    assert lib.dsk_perm_create(a.ctypes.data, b.ctypes.data, 32, 0, ctypes.byref(h)) == 0
    want = oc.minhash_bulk_u32tok(tok[:nt], off, perms)
    need = lib.dsk_minhash_bulk_workspace_size(len(lens), nt)
    assert need > 0 and lib.dsk_minhash_bulk_workspace_size(5, 4096) == 0
    ws = np.full(need // 8 + 4, 0xAB, dtype=np.uint64)
    out = np.zeros((len(lens), 32), dtype=np.uint32)
    args = (h, tok.ctypes.data, 0, off.ctypes.data, len(lens), nt, None, 0, 0, out.ctypes.data, 0, 0)
    assert lib.dsk_minhash_bulk_ws(*args, ws.ctypes.data, need, None) == 0
    assert np.array_equal(out, want)
    assert int(ws.view(np.uint32)[0]) == 59 + 20                     # 1024-token pieces of the two long documents
    out[:] = 0
    assert lib.dsk_minhash_bulk_ws(*args, None, 0, None) == 0        # no workspace: one warp per document, same rows
    assert np.array_equal(out, want)
    assert lib.dsk_minhash_bulk_ws(*args, ws.ctypes.data, need - 1, None) != 0 and b"workspace" in lib.dsk_last_error()
    assert lib.dsk_minhash_bulk_ws(*args, ws.ctypes.data + 4, need, None) != 0
    lib.dsk_perm_destroy(h)


def test_interior_offsets_are_validated_by_the_library(dsk_on_emu):
    """A decreasing interior offset is refused by dsk_minhash_bulk_host (DSK_ERR_INVALID -> ValueError) before any
    token is indexed with it -- including one hidden inside a slice whose end points look fine."""
    dsk = dsk_on_emu
    P = dsk.minhash._make_permutations(16, 1)
    tok = np.arange(10, dtype=np.uint32)
    for off in ([0, 6, 4, 10], [0, 2, 9, 3, 10], [0, 10, 0, 10]):
        with pytest.raises(ValueError):
            dsk.engine.bulk_signatures(tok, np.array(off, dtype=np.int64), P)
    assert dsk.engine.bulk_signatures(tok, np.array([0, 4, 4, 10], dtype=np.int64), P).shape == (3, 16)
