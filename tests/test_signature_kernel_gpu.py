"""The two-phase signature kernel (csrc/signature_kernel.cu) on the GPU against the C oracle: a randomised sweep over
the shapes that take its different paths -- in-place vs copied staging, ring wrap, sub-pieces, padding groups, repeated
tokens (flagged permutations, de-duplicating staging), tiny values next to the L'-7 wrap, long documents cut into pieces
on the device, K slices, running state, u64 output -- through the device-buffer entry (`dsk_minhash_bulk_ws`).
Reference: datasketch/minhash.py:294-297 (update_batch), :464-522 (bulk); oracle = oracle/oracle_c.c (pinned to the
reference's fixtures by tests/test_oracle_golden.py)."""
import numpy as np
import pytest

from oracle import oracle_clib as oc
from oracle import oracle_np as o

pytestmark = pytest.mark.gpu


def _case(rs, style):
    if style == 0:      # aligned, in place
        t = int(rs.choice([16, 64, 256, 512, 1024]))
        lens = np.full(int(rs.randint(50, 400)), t)
    elif style == 1:    # ragged short
        lens = rs.randint(0, 120, size=int(rs.randint(100, 2000)))
    elif style == 2:    # ragged medium, some empty
        lens = rs.randint(0, 700, size=int(rs.randint(50, 600)))
        lens[rs.randint(0, len(lens), 5)] = 0
    elif style == 3:    # a few long documents between short ones
        lens = rs.randint(0, 300, size=int(rs.randint(20, 200)))
        for i in rs.randint(0, len(lens), 4):
            lens[i] = int(rs.choice([4096, 4097, 5000, 20_000, 70_000]))
    else:               # one block / one group documents
        lens = rs.randint(1, 20, size=int(rs.randint(200, 3000)))
    off = np.zeros(len(lens) + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    T = int(off[-1])
    tok = rs.randint(0, 1 << 32, size=T + 1, dtype=np.uint64).astype(np.uint32)[:T]
    u = rs.uniform()
    if u < 0.4 and T > 2:         # repeated tokens inside documents
        share = rs.uniform(0.01, 0.7)
        doc = np.repeat(np.arange(len(lens)), lens)
        pos = np.arange(T) - off[doc]
        rep = (rs.uniform(size=T) < share) & (pos > 0)
        src = off[doc] + (rs.uniform(size=T) * pos).astype(np.int64)
        tok = np.where(rep, tok[src], tok)
    elif u < 0.55 and T > 0:      # tiny / huge values: products next to the wrap, many exact ties
        tok[: T // 2] = rs.randint(0, 30, size=T // 2)
        tok[T // 2:] = (np.uint64(1 << 32) - rs.randint(1, 30, size=T - T // 2).astype(np.uint64)).astype(np.uint32)
    return np.ascontiguousarray(tok), off


@pytest.mark.parametrize("seed", range(6))
def test_randomised_shapes_against_the_oracle(seed):
    import torch
    import datasketch_b200 as dsk
    rs = np.random.RandomState(1000 + seed)
    for it in range(10):
        style = int(rs.randint(0, 5))
        tok, off = _case(rs, style)
        k = int(rs.choice([16, 64, 100, 128, 128, 128, 200, 256, 300]))
        P = o.init_permutations(k, int(rs.randint(1, 40)))
        want = oc.minhash_bulk_u32tok(tok, off, P).astype(np.uint64)
        n = len(off) - 1
        d_tok = torch.from_numpy(np.concatenate([tok, np.zeros(4, np.uint32)]).view(np.int32)).cuda()[:len(tok)]
        d_off = torch.from_numpy(off).cuda()
        mode = int(rs.randint(0, 3))
        if mode == 0:
            got = dsk.engine.bulk_signatures_device(d_tok, d_off, len(tok), P).cpu().numpy().view(np.uint32).astype(np.uint64)
        else:
            init = rs.randint(0, 1 << 32, size=(n, k) if mode == 1 else (1, k), dtype=np.uint64)
            d_init = torch.from_numpy(init.view(np.int64)).cuda()
            d_out = torch.empty((n, k), dtype=torch.int64, device="cuda")
            dsk.engine.bulk_signatures_device(d_tok, d_off, len(tok), P, d_out=d_out, d_init=d_init,
                                              init_stride=k if mode == 1 else 0)
            got = d_out.cpu().numpy().view(np.uint64)
            want = np.minimum(want, init if mode == 1 else init[0][None, :])
        assert np.array_equal(got, want), dict(seed=seed, it=it, style=style, k=k, mode=mode, docs=n, tokens=len(tok))


def test_repeated_calls_reuse_counters_and_workspace():
    """More launches of one permutation handle than it has counter sets (64), on two streams, with long documents in every
    batch: the leases (event-guarded) and the per-call workspaces keep every result equal to the oracle's."""
    import torch
    import datasketch_b200 as dsk
    rs = np.random.RandomState(5)
    lens = np.array([300] * 30 + [9000, 40, 12_000], dtype=np.int64)
    off = np.zeros(len(lens) + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    tok = rs.randint(0, 1 << 32, size=int(off[-1]), dtype=np.uint64).astype(np.uint32)
    P = o.init_permutations(128, 3)
    want = oc.minhash_bulk_u32tok(tok, off, P)
    d_tok = torch.from_numpy(tok.view(np.int32)).cuda()
    d_off = torch.from_numpy(off).cuda()
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    outs = []
    for i in range(150):
        with torch.cuda.stream(streams[i & 1]):
            outs.append(dsk.engine.bulk_signatures_device(d_tok, d_off, len(tok), P))
    torch.cuda.synchronize()
    for i in (0, 1, 63, 64, 65, 128, 149):
        assert np.array_equal(outs[i].cpu().numpy().view(np.uint32), want), i


# ---- general variants: 64-bit tokens (GEN = 2) and permutations that reach the conditional subtract (GEN = 1) ---------
P61 = (1 << 61) - 1


def _unsafe_perms(k, rs):
    """Every third permutation is built so that one small token lands in the set where `% (2^61-1)` subtracts."""
    perms = o.init_permutations(k, 6).copy()
    for i in range(0, k, 3):
        j = int(rs.randint(0, 8))
        a = int(rs.choice([1, 3, 5, 1 << 20, (1 << 40) + 1]))
        h = int(rs.randint(0, 200))
        x = (j << 61) | (P61 - int(rs.randint(0, j + 1)))
        perms[0, i] = a
        perms[1, i] = (x - a * h) % (1 << 64)
    return perms


@pytest.mark.parametrize("seed", range(4))
def test_u64_tokens_randomised_shapes_against_the_oracle(seed):
    """The same sweep with 64-bit hash values (minhash.py:294 accepts anything below 2^64) through both entries."""
    import torch
    import datasketch_b200 as dsk
    rs = np.random.RandomState(3000 + seed)
    for it in range(6):
        style = int(rs.randint(0, 5))
        tok32, off = _case(rs, style)
        T = len(tok32)
        hi = rs.randint(0, 1 << 32, size=T, dtype=np.uint64)
        if it % 3 == 0:
            hi[:] = hi % np.uint64(3)          # few distinct high words: equal low words often are equal tokens
        tok = (hi << np.uint64(32)) | tok32.astype(np.uint64)
        k = int(rs.choice([64, 100, 128, 128, 256, 300]))
        P = o.init_permutations(k, int(rs.randint(1, 40)))
        want = oc.minhash_bulk_u64tok(tok, off, P)
        d_tok = torch.from_numpy(np.concatenate([tok, np.zeros(2, np.uint64)]).view(np.int64)).cuda()[:T]
        d_off = torch.from_numpy(off).cuda()
        got = dsk.engine.bulk_signatures_device(d_tok, d_off, T, P).cpu().numpy().view(np.uint32)
        assert np.array_equal(got, want), dict(seed=seed, it=it, style=style, k=k, docs=len(off) - 1, tokens=T)
        if it < 2:
            assert np.array_equal(dsk.engine.bulk_signatures(tok, off, P), want)
            assert np.array_equal(dsk.engine.bulk_signatures(tok, off, P, kernel="exact"), want)


def test_unsafe_permutations_randomised_against_the_oracle():
    import torch
    import datasketch_b200 as dsk
    from datasketch_b200 import _native as nv
    rs = np.random.RandomState(77)
    for it in range(6):
        tok, off = _case(rs, int(rs.randint(0, 5)))
        tok[::3] = rs.randint(0, 200, size=len(tok[::3]))      # small tokens hit the subtract set of the built permutations
        k = int(rs.choice([33, 128, 256]))
        P = _unsafe_perms(k, rs)
        assert nv.perm_handle(P).n_unsafe > 0
        want = oc.minhash_bulk_u32tok(tok, off, P)
        d_tok = torch.from_numpy(np.concatenate([tok, np.zeros(4, np.uint32)]).view(np.int32)).cuda()[:len(tok)]
        d_off = torch.from_numpy(off).cuda()
        got = dsk.engine.bulk_signatures_device(d_tok, d_off, len(tok), P).cpu().numpy().view(np.uint32)
        assert np.array_equal(got, want), it
        assert np.array_equal(dsk.engine.bulk_signatures(tok, off, P, kernel="two_phase"), want)
    # the general variant on ordinary input equals the default one (same documents, a handle that is flagged unsafe)
    tok, off = _case(rs, 2)
    P = o.init_permutations(128, 1)
    Pu = P.copy()
    Pu[0, 5], Pu[1, 5] = 1, P61 - 3                         # token 3 -> x = p: the subtract fires
    assert nv.perm_handle(Pu).n_unsafe == 1
    got = dsk.engine.bulk_signatures(tok, off, Pu)
    ref = dsk.engine.bulk_signatures(tok, off, P)
    keep = np.arange(128) != 5
    assert np.array_equal(got[:, keep], ref[:, keep]) and np.array_equal(got, oc.minhash_bulk_u32tok(tok, off, Pu))


def test_u64_tokens_full_size_properties():
    """250k documents x 256 64-bit tokens: sampled rows against the C oracle, plus size-independent properties --
    reversing every document, and appending a copy of each document's first half, leave the matrix unchanged."""
    import torch
    import datasketch_b200 as dsk
    rs = np.random.RandomState(9)
    n, t = 250_000, 256
    tok = rs.randint(0, 1 << 63, size=n * t, dtype=np.uint64) * np.uint64(2) + np.uint64(1)
    off = np.arange(n + 1, dtype=np.int64) * t
    P = o.init_permutations(128, 1)
    d_tok = torch.from_numpy(tok.view(np.int64)).cuda()
    d_off = torch.from_numpy(off).cuda()
    sig = dsk.engine.bulk_signatures_device(d_tok, d_off, tok.size, P)
    idx = np.arange(0, n, 997)
    sub = tok.reshape(n, t)[idx].reshape(-1)
    sub_off = np.arange(len(idx) + 1, dtype=np.int64) * t
    got = sig.cpu().numpy().view(np.uint32)
    assert np.array_equal(got[idx], oc.minhash_bulk_u64tok(sub, sub_off, P))
    d_rev = torch.flip(d_tok.view(n, t), dims=[1]).contiguous().view(-1)
    assert torch.equal(dsk.engine.bulk_signatures_device(d_rev, d_off, tok.size, P), sig)
    d_dup = torch.cat([d_tok.view(n, t), d_tok.view(n, t)[:, : t // 2]], dim=1).contiguous().view(-1)
    d_off2 = torch.from_numpy(np.arange(n + 1, dtype=np.int64) * (t + t // 2)).cuda()
    assert torch.equal(dsk.engine.bulk_signatures_device(d_dup, d_off2, d_dup.numel(), P), sig)


@pytest.mark.parametrize("u64", [False, True])
def test_general_variants_structured_tokens_wraps_and_extreme_parameters(u64):
    """Inputs random data never produces (the same case runs on the CPU emulation): tiny and near-2^32 low words (L' < 8:
    the wrap of L' - 8, many exact ties), high words 0 / 1 / 2^32-1, tokens 0 and 2^64-1, permutations at the ends of their
    ranges (a = 1, a = p - 1, b = 0, b = p - 1, a with a zero low word) next to ones built to hit the conditional subtract."""
    import datasketch_b200 as dsk
    rs = np.random.RandomState(2 if u64 else 1)
    lo = np.concatenate([np.arange(0, 1500, dtype=np.uint64), np.uint64(1 << 32) - np.arange(1, 1501, dtype=np.uint64),
                         np.zeros(250, dtype=np.uint64), np.full(250, 0xFFFFFFFF, dtype=np.uint64)])
    k = 64
    perms = _unsafe_perms(k, rs)
    perms[0, 1], perms[1, 1] = 1, 0
    perms[0, 2], perms[1, 2] = P61 - 1, P61 - 1
    perms[0, 4], perms[1, 4] = np.uint64(7 << 32), 5              # a_lo = 0: L' is the same for every token
    perms[0, 5], perms[1, 5] = 1, P61 - 1                          # x = h + p - 1: the subtract fires for h = 1 .. 8
    off = np.arange(0, len(lo) + 1, 125, dtype=np.int64)
    if u64:
        hi = rs.choice(np.array([0, 1, 0xFFFFFFFF, 0x80000000, 12345], dtype=np.uint64), size=len(lo))
        tok = (hi << np.uint64(32)) | lo
        tok[7], tok[130] = np.uint64(0), np.uint64(0xFFFFFFFFFFFFFFFF)
        want = oc.minhash_bulk_u64tok(tok, off, perms)
    else:
        tok = lo.astype(np.uint32)
        want = oc.minhash_bulk_u32tok(tok, off, perms)
    for kernel in ("auto", "exact"):
        assert np.array_equal(dsk.engine.bulk_signatures(tok, off, perms, kernel=kernel), want), kernel
