/* TEST INFRASTRUCTURE ONLY -- plain-C restatement (oracle) of the integer hot path.
 *
 * Not part of the product.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library (oracle/_build/).
 * Parity: PINNED -- tests/test_oracle_golden.py checks it against fixtures
 * generated from the real reference (oracle/gen_golden.py).
 *
 * Each function restates the numpy uint64 arithmetic the reference executes:
 *   datasketch/minhash.py:294-297   phv = ((hv*a + b) mod 2^64) % (2^61-1)) & (2^32-1); column min
 *   datasketch/minhash.py:324       jaccard = count(h1 == h2) / K
 *   datasketch/lean_minhash.py:174  record = 'q' seed | 'i' K | K x 'I' (low 32 bits)
 *   datasketch/lsh.py:537-538       band key = r uint64 values, byte-swapped to big endian
 */
#include <stdint.h>
#include <stddef.h>
#include <string.h>

#define ORACLE_API __attribute__((visibility("default")))

static const uint64_t P61 = ((uint64_t)1 << 61) - 1;

/* One (token, permutation) evaluation: C unsigned arithmetic wraps mod 2^64
 * exactly like numpy uint64 (minhash.py:295), then the true remainder. */
static inline uint32_t permute(uint64_t a, uint64_t b, uint64_t h) {
    uint64_t x = a * h + b;
    return (uint32_t)((x % P61) & 0xFFFFFFFFull);
}

/* CSR batch of documents -> [n_docs, k] uint64 signatures (MinHash.bulk,
 * minhash.py:464-522 over :294-297).  tokens are uint64 hash values. */
ORACLE_API void oracle_minhash_bulk_u64tok(const uint64_t *tokens, const int64_t *offsets,
                                           int64_t n_docs, const uint64_t *a, const uint64_t *b,
                                           int k, uint64_t *out) {
    for (int64_t d = 0; d < n_docs; ++d) {
        uint64_t *row = out + (size_t)d * k;
        for (int j = 0; j < k; ++j) row[j] = 0xFFFFFFFFull;      /* minhash.py:167-168 */
        for (int64_t t = offsets[d]; t < offsets[d + 1]; ++t) {
            uint64_t h = tokens[t];
            for (int j = 0; j < k; ++j) {
                uint32_t v = permute(a[j], b[j], h);
                if (v < row[j]) row[j] = v;
            }
        }
    }
}

/* Same with uint32 token storage (the documented 32-bit hashfunc contract,
 * minhash.py:69-70) and docs [d0, d1) only -- lets the caller thread it. */
ORACLE_API void oracle_minhash_bulk_u32tok(const uint32_t *tokens, const int64_t *offsets,
                                           int64_t d0, int64_t d1, const uint64_t *a,
                                           const uint64_t *b, int k, uint32_t *out) {
    for (int64_t d = d0; d < d1; ++d) {
        uint32_t *row = out + (size_t)d * k;
        for (int j = 0; j < k; ++j) row[j] = 0xFFFFFFFFu;
        for (int64_t t = offsets[d]; t < offsets[d + 1]; ++t) {
            uint64_t h = tokens[t];
            for (int j = 0; j < k; ++j) {
                uint32_t v = permute(a[j], b[j], h);
                if (v < row[j]) row[j] = v;
            }
        }
    }
}

/* count of equal positions for m pairs (minhash.py:324 numerator). */
ORACLE_API void oracle_jaccard_pairs_u32(const uint32_t *sig, const int64_t *ia, const int64_t *ib,
                                         int64_t m, int k, int32_t *out_count) {
    for (int64_t p = 0; p < m; ++p) {
        const uint32_t *x = sig + (size_t)ia[p] * k, *y = sig + (size_t)ib[p] * k;
        int32_t c = 0;
        for (int j = 0; j < k; ++j) c += (x[j] == y[j]);
        out_count[p] = c;
    }
}

/* LeanMinHash records, little-endian / native layout with no padding
 * (lean_minhash.py:174-175, byteorder '<' or '@' on x86-64: 8 + 4 + 4K bytes). */
ORACLE_API void oracle_lean_pack_le(const uint32_t *sig, int64_t n, int k, int64_t seed, uint8_t *out) {
    size_t rec = 12 + (size_t)4 * k;
    for (int64_t i = 0; i < n; ++i) {
        uint8_t *p = out + (size_t)i * rec;
        int32_t kk = k;
        memcpy(p, &seed, 8);
        memcpy(p + 8, &kk, 4);
        memcpy(p + 12, sig + (size_t)i * k, (size_t)4 * k);
    }
}

/* Big-endian band keys: for doc i, band j, r values each as 8 big-endian bytes
 * (lsh.py:344 + :537-538).  out is [n, b, 8*r] bytes. */
ORACLE_API void oracle_band_keys_be(const uint32_t *sig, int64_t n, int k, int b, int r, uint8_t *out) {
    for (int64_t i = 0; i < n; ++i)
        for (int j = 0; j < b; ++j)
            for (int q = 0; q < r; ++q) {
                uint64_t v = sig[(size_t)i * k + (size_t)j * r + q];
                uint8_t *p = out + (((size_t)i * b + j) * r + q) * 8;
                for (int s = 0; s < 8; ++s) p[s] = (uint8_t)(v >> (56 - 8 * s));
            }
}
