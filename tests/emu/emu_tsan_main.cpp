// emu_tsan_main.cpp -- TEST INFRASTRUCTURE ONLY: the emulated kernel under ThreadSanitizer.
//   g++ -std=c++17 -O1 -g -pthread -fsanitize=thread -DDSK_EMU -Itests/emu tests/emu/emu_minhash.cpp tests/emu/emu_tsan_main.cpp
// Every CUDA thread is a host thread and the warp collectives are pthread barriers, so a shared-memory access that the
// kernel does not order with __syncwarp / the mbarrier (e.g. refilling a ring slot that a lane is still reading) is a
// data race TSan reports.  Also checks the results against a scalar evaluation of the reference formula.
// Run with TSAN_OPTIONS=history_size=7: with the default history this TSan version cannot restore the earlier access's
// stack for these short-lived lane threads and silently drops the report (seen in a control experiment).
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

extern "C" int emu_lean_pack(const void *sig, int sig_is_u64, int64_t n, int k, int64_t seed, int big_endian, uint8_t *rec,
                             int sm_count);
extern "C" int emu_lean_unpack(const uint8_t *rec, int64_t n, int k, int64_t seed, int big_endian, void *sig, int sig_is_u64,
                               int *status, int sm_count);
extern "C" void *emu_lsh_create(int k, int b, int r, int64_t cap_docs);
extern "C" void emu_lsh_destroy(void *h);
extern "C" int emu_lsh_insert(void *h, const uint32_t *sig, int64_t n, int sm_count);
extern "C" int64_t emu_lsh_query_count(void *h, const uint32_t *q, int64_t nq, int64_t *ptr, int sm_count);
extern "C" int emu_lsh_query_fill(void *h, const uint32_t *q, int64_t nq, const int64_t *ptr, int32_t *idx, int sm_count);
extern "C" int emu_minhash_bulk(const void *tokens, int token_is_u64, const int64_t *offsets, int64_t n_docs,
                                const uint64_t *a, const uint64_t *b, int k, int mode, int v1, const void *init,
                                int64_t init_stride, int init_is_u64, void *out, int out_is_u64, int docs_per_unit,
                                int grid_x, int gen);

static uint64_t rng_state = 88172645463325252ull;
static uint64_t rnd() {
    rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17;
    return rng_state;
}

int main() {
    const uint64_t p61 = (1ull << 61) - 1;
    const int k = 128, n_docs = 24;
    std::vector<uint64_t> a(k), b(k);
    for (int i = 0; i < k; ++i) { a[i] = 1 + rnd() % (p61 - 1); b[i] = rnd() % p61; }
    std::vector<int64_t> off(n_docs + 1, 0);
    for (int d = 0; d < n_docs; ++d) {
        const int64_t len = (d % 6 == 0) ? 2500 + (int64_t)(rnd() % 1500) : (int64_t)(rnd() % 300);  // some longer than the ring
        off[d + 1] = off[d] + (d == 3 ? 0 : len);
    }
    std::vector<uint32_t> tok(off[n_docs] + 16);
    for (auto &t : tok) t = (uint32_t)rnd();
    for (int d = 0; d < n_docs; ++d)          // repeated tokens: exact ties -> the re-scan paths run too
        for (int64_t i = off[d] + 1; i < off[d + 1]; ++i)
            if (rnd() % 5 == 0) tok[i] = tok[off[d] + (int64_t)(rnd() % (uint64_t)(i - off[d]))];
    std::vector<uint32_t> want((size_t)n_docs * k, 0xFFFFFFFFu);
    for (int d = 0; d < n_docs; ++d)
        for (int64_t i = off[d]; i < off[d + 1]; ++i)
            for (int j = 0; j < k; ++j) {
                const uint64_t x = a[j] * (uint64_t)tok[i] + b[j];
                const uint32_t r = (uint32_t)(x % p61);
                if (r < want[(size_t)d * k + j]) want[(size_t)d * k + j] = r;
            }
    int bad = 0;
    for (int mode = 0; mode < 3; ++mode)
        for (int v1 = 0; v1 < (mode == 0 ? 2 : 1); ++v1) {
            std::vector<uint32_t> got((size_t)n_docs * k, 0);
            emu_minhash_bulk(tok.data(), 0, off.data(), n_docs, a.data(), b.data(), k, mode, v1, nullptr, 0, 0,
                             got.data(), 0, 3, 2, 0);
            const bool ok = got == want;
            printf("mode %d v1 %d: %s\n", mode, v1, ok ? "identical" : "MISMATCH");
            bad += !ok;
        }
    {   // general variants: u32 tokens with GEN = 1, and 64-bit tokens (GEN = 2: word ring, two staged planes, 64-bit
        // de-duplication table) -- repeats kept (equal u32 tokens -> equal u64 tokens), so the flagged path and the
        // de-duplicating stage run under the race detector too
        std::vector<uint32_t> got((size_t)n_docs * k, 0);
        emu_minhash_bulk(tok.data(), 0, off.data(), n_docs, a.data(), b.data(), k, 0, 0, nullptr, 0, 0, got.data(), 0, 3, 2, 1);
        bool ok = got == want;
        printf("general u32 variant: %s\n", ok ? "identical" : "MISMATCH");
        bad += !ok;
        std::vector<uint64_t> tok64(tok.size());
        for (size_t i = 0; i < tok.size(); ++i) tok64[i] = ((uint64_t)(tok[i] * 2654435761u) << 32) | tok[i];
        std::vector<uint32_t> want64((size_t)n_docs * k, 0xFFFFFFFFu);
        for (int d = 0; d < n_docs; ++d)
            for (int64_t i = off[d]; i < off[d + 1]; ++i)
                for (int j = 0; j < k; ++j) {
                    const uint64_t x = a[j] * tok64[i] + b[j];
                    const uint32_t r = (uint32_t)(x % p61);
                    if (r < want64[(size_t)d * k + j]) want64[(size_t)d * k + j] = r;
                }
        emu_minhash_bulk(tok64.data(), 1, off.data(), n_docs, a.data(), b.data(), k, 0, 0, nullptr, 0, 0, got.data(), 0, 3, 2, 0);
        ok = got == want64;
        printf("u64 tokens: %s\n", ok ? "identical" : "MISMATCH");
        bad += !ok;
    }
    {   // LeanMinHash codec: 4-stage bulk-copy tile pipeline, several tiles per CTA, ragged last tile
        const int lk = 128;
        const int64_t ln = 777;
        std::vector<uint32_t> sig((size_t)ln * lk), back((size_t)ln * lk, 0);
        for (auto &v : sig) v = (uint32_t)rnd();
        std::vector<uint8_t> rec((size_t)ln * (12 + 4 * lk) + 64);
        uint8_t *recp = rec.data() + (16 - (reinterpret_cast<uintptr_t>(rec.data()) & 15)) % 16;
        int status = 0;
        emu_lean_pack(sig.data(), 0, ln, lk, 42, 1, recp, 1);
        emu_lean_unpack(recp, ln, lk, 42, 1, back.data(), 0, &status, 1);
        const bool ok = back == sig && status == 0 && recp[0] == 0 && recp[7] == 42 && recp[11] == lk;
        printf("lean codec round trip: %s\n", ok ? "identical" : "MISMATCH");
        bad += !ok;
    }
    {   // LSH index: concurrent atomicCAS slot claims and atomicExch chain pushes from many threads, then queries
        const int k = 32, b = 8, r = 4;
        const int64_t n = 400;
        std::vector<uint32_t> sig((size_t)n * k);
        for (auto &v : sig) v = (uint32_t)(rnd() % 3);                     // low entropy: long chains, shared buckets
        void *h = emu_lsh_create(k, b, r, n);
        emu_lsh_insert(h, sig.data(), n, 2);
        std::vector<int64_t> ptr(n + 1, 0);
        const int64_t total = emu_lsh_query_count(h, sig.data(), n, ptr.data(), 2);
        std::vector<int32_t> idx((size_t)(total > 0 ? total : 1), -1);
        emu_lsh_query_fill(h, sig.data(), n, ptr.data(), idx.data(), 2);
        bool ok = total >= n;
        for (int64_t i = 0; i < n && ok; ++i) {                              // every document finds itself, exactly once
            int self = 0;
            for (int64_t p = ptr[i]; p < ptr[i + 1]; ++p) self += idx[p] == (int32_t)i;
            ok = self == 1;
        }
        emu_lsh_destroy(h);
        printf("lsh insert/query: %s\n", ok ? "identical" : "MISMATCH");
        bad += !ok;
    }
    return bad ? 1 : 0;
}
