"""CPU model of the arithmetic shortcuts the CUDA kernels rely on, checked against the
oracle.  These are the exactness arguments themselves, independent of any GPU:

  1. fast formula   r = lo32(x) + (hi32(x) >> 29)  (mod 2^32) equals ((a*h+b) mod 2^64) % (2^61-1) & (2^32-1)
                    for every x except the 36 values where the `% p` subtract fires;
  2. window         L' = lo32(a_lo*h + b_lo + 7) satisfies  L' - r in [0, 7]  (mod 2^32);
  3. two-phase      tracking (min L', second-smallest block min, winning block) + re-evaluating only the
                    winning block, with the exact re-scan when another block is within the window or min L' < 7,
                    reproduces the signature -- including documents with repeated tokens and tokens whose
                    products sit next to the 2^32 wrap.
"""
import numpy as np
import pytest

from oracle import oracle_np as o

P61 = np.uint64((1 << 61) - 1)
M32 = np.uint64(0xFFFFFFFF)


def _x(a, b, h):
    with np.errstate(over="ignore"):
        return a * h + b                      # uint64, wraps mod 2^64 like the reference (minhash.py:294)


def _exact(a, b, h):
    return (_x(a, b, h) % P61) & M32


def _fast(a, b, h):
    x = _x(a, b, h)
    return ((x & M32) + (x >> np.uint64(61))) & M32      # lo32(x) + top3(x)  ==  lo32 + (hi32 >> 29)


def _lprime(a, b, h):
    return ((a & M32) * h + (b & M32) + np.uint64(7)) & M32


def test_fast_formula_and_window_identities():
    rs = np.random.RandomState(1)
    a = rs.randint(1, int(P61), size=400_000, dtype=np.uint64)
    b = rs.randint(0, int(P61), size=400_000, dtype=np.uint64)
    h = rs.randint(0, 1 << 32, size=400_000, dtype=np.uint64)
    # adversarial tokens: products around the 2^32 / 2^61 / 2^64 boundaries
    h[:1000] = np.uint64(0xFFFFFFFF) - np.arange(1000, dtype=np.uint64)
    h[1000:2000] = np.arange(1000, dtype=np.uint64)
    x = _x(a, b, h)
    s = (x & P61) + (x >> np.uint64(61))
    fires = s >= P61                                       # the `% p` conditional subtract
    assert not fires.any()                                 # 36 of 2^64 values: never hit at random
    assert np.array_equal(_fast(a, b, h), _exact(a, b, h))
    d = (_lprime(a, b, h) - _fast(a, b, h)) & M32
    assert d.max() <= 7 and np.array_equal(d, np.uint64(7) - (x >> np.uint64(61)))
    # the subtract case itself: x = k*p + (p + e) style values are the only ones where the formulas part
    for xv in [int(P61), 2 * int(P61), 7 * int(P61) + 7, (1 << 64) - 1]:
        xs = np.uint64(xv)
        s1 = int((xs & P61) + (xs >> np.uint64(61)))
        assert (s1 >= int(P61)) == (xv % int(P61) != s1)


def _two_phase(tok, a, b, blk=16):
    """Signature of one document by the kernel's two-phase scheme (numpy model, all permutations at once)."""
    k = a.shape[0]
    if tok.size == 0:
        return np.full(k, 0xFFFFFFFF, dtype=np.uint64)
    nb = -(-tok.size // blk)
    pad = np.concatenate([tok, np.repeat(tok[-1:], nb * blk - tok.size)])      # duplicates: min is idempotent
    L = _lprime(a[None, :], b[None, :], pad[:, None]).reshape(nb, blk, k)
    bm = L.min(axis=1)                                                           # [nb, k] block minima
    order = np.argsort(bm, axis=0, kind="stable")
    widx = order[0]
    m = bm[widx, np.arange(k)]
    m2 = bm[order[1], np.arange(k)] if nb > 1 else np.full(k, 0xFFFFFFFF, dtype=np.uint64)
    slow = (m < 7) | ((m2 - m) <= 7)
    out = np.empty(k, dtype=np.uint64)
    R = _fast(a[None, :], b[None, :], pad[:, None]).reshape(nb, blk, k)
    for j in range(k):
        if not slow[j]:
            out[j] = R[widx[j], :, j].min()                                      # phase 2: the winning block only
        else:                                                                    # exact re-scan
            cand = np.ones(nb, bool) if m[j] < 7 else (bm[:, j] - m[j]) <= 7
            out[j] = R[cand, :, j].min()
    return out, int(slow.sum())


@pytest.mark.parametrize("share", [0.0, 0.1, 0.9])
def test_two_phase_model_matches_oracle_with_repeated_tokens(share):
    rs = np.random.RandomState(int(share * 10) + 3)
    k = 64
    perms = o.init_permutations(k, 1)
    a, b = perms[0], perms[1]
    n_slow = 0
    for d in range(60):
        n = int(rs.randint(1, 300))
        tok = rs.randint(0, 1 << 32, size=n, dtype=np.uint64)
        if d % 7 == 0:
            tok[: n // 2] = rs.randint(0, 50, size=n // 2)                       # tiny tokens
        if d % 11 == 0:
            tok[: n // 3] = np.uint64(0xFFFFFFFF) - rs.randint(0, 50, size=n // 3).astype(np.uint64)
        rep = np.nonzero(rs.uniform(size=n) < share)[0]
        rep = rep[rep > 0]
        tok[rep] = tok[(rs.uniform(size=len(rep)) * rep).astype(np.int64)]        # repeats of earlier tokens
        got, ns = _two_phase(tok, a, b)
        if d % 7 and d % 11:                                                     # purely random documents only
            n_slow += ns
        want = o.update_batch(o.init_hashvalues(k), tok.tolist(), perms)
        assert np.array_equal(got, want)
    assert (n_slow > 0) == (share > 0)              # repeats are what sends a permutation to the re-scan


def test_small_minimum_needs_the_full_rescan():
    """Why `min L' < 7` must re-scan everything: a token whose lo32(x) is just below 2^32 has a tiny (wrapped) L'
    but a huge r, so the window around min L' can miss the true minimum."""
    k = 256
    perms = o.init_permutations(k, 2)
    a, b = perms[0], perms[1]
    rs = np.random.RandomState(9)
    hit = 0
    for j in range(k):
        inv = pow(int(a[j]) & 0xFFFFFFFF | 1, -1, 1 << 32)
        if int(a[j]) & 1 == 0:
            continue
        # token with lo32(a_lo*h + b_lo) = 2^32 - 3  ->  L' = 4 (wrapped), r = 2^32 - 3 + top3
        h = ((((1 << 32) - 3 - (int(b[j]) & 0xFFFFFFFF)) % (1 << 32)) * inv) % (1 << 32)
        tok = np.concatenate([[np.uint64(h)], rs.randint(0, 1 << 32, size=40, dtype=np.uint64)])
        got, _ = _two_phase(tok, a, b)
        want = o.update_batch(o.init_hashvalues(k), tok.tolist(), perms)
        assert np.array_equal(got, want)
        hit += int(_lprime(a[j], b[j], np.uint64(h)) < 7)
    assert hit > 50
