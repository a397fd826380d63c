// emu_kernels.cpp -- TEST INFRASTRUCTURE ONLY: runs the kernels AND launchers of datasketch_b200/csrc/*_kernels.cu
// on host threads (tests/emu/cuda_emu.h) so the CPU test-suite can check their logic against the oracle.
//   g++ -std=c++17 -O1 -pthread -DDSK_EMU -Itests/emu -shared -fPIC tests/emu/emu_kernels.cpp
#include "cuda_emu.h"

thread_local uint3e threadIdx, blockIdx, gridDim, blockDim;
thread_local EmuWarp *emu_warp = nullptr;
thread_local int emu_lane = 0;
thread_local EmuCta *emu_cta = nullptr;

#include "../../datasketch_b200/csrc/minhash_kernels.cu"
#include "../../datasketch_b200/csrc/signature_kernel.cu"
#include "../../datasketch_b200/csrc/codec_kernels.cu"
#include "../../datasketch_b200/csrc/lsh_kernels.cu"
#include "../../datasketch_b200/csrc/jaccard_kernels.cu"
#include "../../datasketch_b200/csrc/sha1_kernels.cu"
#include "../../datasketch_b200/csrc/hash_kernels.cu"
#include "../../datasketch_b200/csrc/wmh_kernels.cu"
#include "../../datasketch_b200/csrc/bloom_kernels.cu"

#include <vector>

// a / b: the uint64 permutation parameters; init (optional): [n_docs or 1][k] u32/u64 running signatures.
// Goes through the product launcher (launch_minhash_bulk: kernel and P selection, slices, counters); docs_per_unit > 0
// forces the unit size, grid_x plays the role of sm_count (the launcher's grid is grid_x * OCC CTAs at most).
// v1 = 1 runs the round-1 two-phase kernel (minhash_bulk_kernel<MODE_TWO_PHASE>) instead of minhash_sig_kernel.
extern "C" int emu_minhash_bulk(const void *tokens, int token_is_u64, const int64_t *offsets, int64_t n_docs,
                                const uint64_t *a, const uint64_t *b, int k, int mode, int v1,
                                const void *init, int64_t init_stride, int init_is_u64, void *out, int out_is_u64,
                                int docs_per_unit, int grid_x, int gen) {
    const int P = k <= 32 ? 1 : k <= 64 ? 2 : k <= 128 ? 4 : 8;
    const int kpad = (k + 255) / 256 * 256;
    std::vector<uint32_t> tab((size_t)6 * kpad);
    for (int i = 0; i < kpad; ++i) {  // padding slots repeat real permutations, like dsk_perm_create
        const int s = i % k;
        tab[i] = (uint32_t)a[s]; tab[kpad + i] = (uint32_t)(a[s] >> 32);
        tab[2 * kpad + i] = (uint32_t)b[s]; tab[3 * kpad + i] = (uint32_t)(b[s] >> 32);
        tab[4 * kpad + i] = (uint32_t)b[s] + 7u; tab[5 * kpad + i] = (uint32_t)b[s] + 8u;
    }
    std::vector<unsigned> counters(64, 0u);
    dsk::BulkParams prm{};
    prm.tokens = tokens; prm.offsets = offsets; prm.n_docs = n_docs; prm.n_tokens = offsets[n_docs];
    prm.a_lo = tab.data(); prm.a_hi = tab.data() + kpad; prm.b_lo = tab.data() + 2 * kpad; prm.b_hi = tab.data() + 3 * kpad; prm.b_lo7 = tab.data() + 4 * kpad;
    prm.b_lo8 = tab.data() + 5 * kpad;
    prm.gen = token_is_u64 ? 2 : gen;   // gen = 1: the general u32 variant (the API picks it when a permutation is unsafe)
    prm.k = k; prm.init = init; prm.init_stride = init_stride; prm.init_is_u64 = init_is_u64;
    prm.out = out; prm.out_is_u64 = out_is_u64; prm.work_counter = counters.data();
    prm.docs_per_unit = docs_per_unit; prm.n_peers = 0; prm.peer_row_offset = 0;
    if (token_is_u64 && mode == dsk::MODE_DIRECT) return -1;
    if (v1 && mode == dsk::MODE_TWO_PHASE) {
        switch (P) {
            case 1: return dsk::launch_bulk<1, dsk::MODE_TWO_PHASE, uint32_t, 4>(prm, grid_x, nullptr);
            case 2: return dsk::launch_bulk<2, dsk::MODE_TWO_PHASE, uint32_t, 4>(prm, grid_x, nullptr);
            case 4: return dsk::launch_bulk<4, dsk::MODE_TWO_PHASE, uint32_t, 4>(prm, grid_x, nullptr);
            default: return dsk::launch_bulk<8, dsk::MODE_TWO_PHASE, uint32_t, 4>(prm, grid_x, nullptr);
        }
    }
    return dsk::launch_minhash_bulk(prm, mode, token_is_u64, grid_x, nullptr);
}

// the signature kernel with the long-document piece table (thresholds chosen by the test, e.g. 100 / 32 tokens)
extern "C" int emu_minhash_sig_long(const void *tokens, const int64_t *offsets, int64_t n_docs, const uint64_t *a,
                                    const uint64_t *b, int k, const void *init, int64_t init_stride, int init_is_u64,
                                    void *out, int out_is_u64, int docs_per_unit, int grid_x, int64_t long_doc_tokens,
                                    int piece_shift, long long *n_pieces_out, int gen) {
    const int kpad = (k + 255) / 256 * 256;
    std::vector<uint32_t> tab((size_t)6 * kpad);
    for (int i = 0; i < kpad; ++i) {
        const int s = i % k;
        tab[i] = (uint32_t)a[s]; tab[kpad + i] = (uint32_t)(a[s] >> 32);
        tab[2 * kpad + i] = (uint32_t)b[s]; tab[3 * kpad + i] = (uint32_t)(b[s] >> 32);
        tab[4 * kpad + i] = (uint32_t)b[s] + 7u; tab[5 * kpad + i] = (uint32_t)b[s] + 8u;
    }
    std::vector<unsigned> counters(64, 0u);
    const int64_t n_tokens = offsets[n_docs];
    const size_t cap = (size_t)((n_tokens >> piece_shift) + n_tokens / long_doc_tokens + 2);
    std::vector<unsigned char> ws(dsk::kPieceHdrBytes + cap * sizeof(dsk::PieceDesc) + 64, 0xCD);   // contents undefined on entry
    unsigned char *wsp = ws.data() + (16 - (reinterpret_cast<uintptr_t>(ws.data()) & 15)) % 16;
    dsk::BulkParams prm{};
    prm.tokens = tokens; prm.offsets = offsets; prm.n_docs = n_docs; prm.n_tokens = n_tokens;
    prm.a_lo = tab.data(); prm.a_hi = tab.data() + kpad; prm.b_lo = tab.data() + 2 * kpad; prm.b_hi = tab.data() + 3 * kpad;
    prm.b_lo7 = tab.data() + 4 * kpad; prm.b_lo8 = tab.data() + 5 * kpad; prm.gen = gen;
    prm.k = k; prm.init = init; prm.init_stride = init_stride; prm.init_is_u64 = init_is_u64;
    prm.out = out; prm.out_is_u64 = out_is_u64; prm.work_counter = counters.data();
    prm.docs_per_unit = docs_per_unit;
    prm.long_doc_tokens = long_doc_tokens; prm.piece_shift = piece_shift;
    prm.piece_hdr = reinterpret_cast<unsigned *>(wsp);
    prm.pieces = reinterpret_cast<dsk::PieceDesc *>(wsp + dsk::kPieceHdrBytes);
    const int rc = dsk::launch_minhash_sig(prm, grid_x, nullptr);
    if (n_pieces_out) *n_pieces_out = *reinterpret_cast<unsigned *>(wsp);
    return rc;
}

// path counters of minhash_sig_kernel (read-and-reset): in place / copied / deduplicated sub-pieces, tokens removed, flagged perms
extern "C" void emu_sig_stats(long long *out5) {
    for (int i = 0; i < dsk::STAT_COUNT; ++i) out5[i] = dsk::g_sig_stat[i].exchange(0);
}

extern "C" int emu_seg_min(const uint32_t *part, const int64_t *seg, int64_t n_docs, int k, const void *init,
                           int64_t init_stride, int init_is_u64, void *out, int out_is_u64) {
    return dsk::launch_seg_min(part, seg, n_docs, k, init, init_stride, init_is_u64, out, out_is_u64, 1, nullptr);
}
extern "C" int emu_sig_merge_min(const uint32_t *x, const uint32_t *y, int64_t n, uint32_t *out) {
    return dsk::launch_sig_merge_min(x, y, n, out, 1, nullptr);
}
extern "C" int emu_lean_pack(const void *sig, int sig_is_u64, int64_t n, int k, int64_t seed, int big_endian, uint8_t *rec,
                             int sm_count) {
    return dsk::launch_lean_pack(sig, sig_is_u64, n, k, seed, big_endian, rec, sm_count, nullptr);
}
extern "C" int emu_lean_unpack(const uint8_t *rec, int64_t n, int k, int64_t seed, int big_endian, void *sig, int sig_is_u64,
                               int *status, int sm_count) {
    return dsk::launch_lean_unpack(rec, n, k, seed, big_endian, sig, sig_is_u64, status, sm_count, nullptr);
}
extern "C" int emu_band_keys(const uint32_t *sig, int64_t n, int k, int b, int r, uint8_t *out) {
    return dsk::launch_band_keys_be(sig, n, k, b, r, out, 1, nullptr);
}
extern "C" int emu_band_fingerprints(const uint32_t *sig, int64_t n, int k, int b, int r, uint64_t *out) {
    return dsk::launch_band_fingerprints(sig, n, k, b, r, out, 1, nullptr);
}
extern "C" int emu_bbit_pack(const uint32_t *sig, int64_t n, int k, int b, int slot, uint64_t *out) {
    return dsk::launch_bbit_pack(sig, n, k, b, slot, out, 1, nullptr);
}
extern "C" int emu_bbit_unpack(const uint64_t *blocks, int64_t n, int k, int slot, uint32_t *sig) {
    return dsk::launch_bbit_unpack(blocks, n, k, slot, sig, 1, nullptr);
}

// ---- device-resident LSH index (lsh_kernels.cu), set up the way dsk_lsh_create does -----------------------------------
struct EmuLsh {
    dsk::LshDev dev;
    std::vector<uint32_t> sig;
    std::vector<uint64_t> slots;
    std::vector<int32_t> next;
    int64_t n_docs = 0;
};
extern "C" void *emu_lsh_create(int k, int b, int r, int64_t cap_docs) {
    EmuLsh *ix = new EmuLsh();
    dsk::LshDev &v = ix->dev;
    v.k = k; v.b = b; v.r = r; v.cap_docs = cap_docs;
    v.cap_slots = 1024;
    while (v.cap_slots < 2 * cap_docs) v.cap_slots <<= 1;
    ix->sig.assign((size_t)cap_docs * k, 0u);
    ix->slots.assign((size_t)b * v.cap_slots * 2, ~0ull);        // cudaMemset 0xFF: empty key, head = -1
    ix->next.assign((size_t)b * cap_docs, 0);
    v.sig = ix->sig.data(); v.slots = ix->slots.data(); v.next = ix->next.data();
    return ix;
}
extern "C" void emu_lsh_destroy(void *h) { delete static_cast<EmuLsh *>(h); }
extern "C" int emu_lsh_insert(void *h, const uint32_t *sig, int64_t n, int sm_count) {
    EmuLsh *ix = static_cast<EmuLsh *>(h);
    if (ix->n_docs + n > ix->dev.cap_docs) return -1;
    const int rc = dsk::launch_lsh_insert(ix->dev, sig, ix->n_docs, n, sm_count, nullptr);
    ix->n_docs += n;
    return rc;
}
// the fused route of dsk_lsh_insert_tokens: signature kernel with the insert epilogue, rows built in the index's storage
extern "C" int emu_lsh_insert_tokens(void *h, const uint32_t *tokens, const int64_t *offsets, int64_t n_docs, const uint64_t *a,
                                     const uint64_t *b, int k, int docs_per_unit, int grid_x) {
    EmuLsh *ix = static_cast<EmuLsh *>(h);
    if (ix->n_docs + n_docs > ix->dev.cap_docs || k != ix->dev.k) return -1;
    const int kpad = (k + 255) / 256 * 256;
    std::vector<uint32_t> tab((size_t)6 * kpad);
    for (int i = 0; i < kpad; ++i) {
        const int s = i % k;
        tab[i] = (uint32_t)a[s]; tab[kpad + i] = (uint32_t)(a[s] >> 32);
        tab[2 * kpad + i] = (uint32_t)b[s]; tab[3 * kpad + i] = (uint32_t)(b[s] >> 32);
        tab[4 * kpad + i] = (uint32_t)b[s] + 7u; tab[5 * kpad + i] = (uint32_t)b[s] + 8u;
    }
    std::vector<unsigned> counters(64, 0u);
    dsk::BulkParams prm{};
    prm.tokens = tokens; prm.offsets = offsets; prm.n_docs = n_docs; prm.n_tokens = offsets[n_docs];
    prm.a_lo = tab.data(); prm.a_hi = tab.data() + kpad; prm.b_lo = tab.data() + 2 * kpad; prm.b_hi = tab.data() + 3 * kpad;
    prm.b_lo7 = tab.data() + 4 * kpad; prm.b_lo8 = tab.data() + 5 * kpad; prm.gen = 0;
    prm.k = k; prm.out = ix->dev.sig + ix->n_docs * k; prm.out_is_u64 = 0; prm.work_counter = counters.data();
    prm.docs_per_unit = docs_per_unit;
    prm.lsh_slots = ix->dev.slots; prm.lsh_next = ix->dev.next; prm.lsh_cap_slots = ix->dev.cap_slots;
    prm.lsh_doc0 = ix->n_docs; prm.lsh_b = ix->dev.b; prm.lsh_r = ix->dev.r;
    const int rc = dsk::launch_minhash_sig(prm, grid_x, nullptr);
    ix->n_docs += n_docs;
    return rc;
}
extern "C" const uint32_t *emu_lsh_rows(void *h) { return static_cast<EmuLsh *>(h)->dev.sig; }
// count -> exclusive scan -> fill, the sequence GpuLSH.query runs; ptr has nq + 1 entries, idx must hold ptr[nq]
extern "C" int64_t emu_lsh_query_count(void *h, const uint32_t *q, int64_t nq, int64_t *ptr, int sm_count) {
    EmuLsh *ix = static_cast<EmuLsh *>(h);
    std::vector<int64_t> counts(nq + 1, 0), scratch(nq / 1024 + 2, 0);
    if (dsk::launch_lsh_query(ix->dev, q, nq, ix->n_docs, counts.data(), nullptr, nullptr, 0, sm_count, nullptr)) return -1;
    if (dsk::launch_exclusive_scan(counts.data(), nq, ptr, scratch.data(), nullptr)) return -1;
    return ptr[nq];
}
extern "C" int emu_lsh_query_fill(void *h, const uint32_t *q, int64_t nq, const int64_t *ptr, int32_t *idx, int sm_count) {
    EmuLsh *ix = static_cast<EmuLsh *>(h);
    return dsk::launch_lsh_query(ix->dev, q, nq, ix->n_docs, nullptr, ptr, idx, 1, sm_count, nullptr);
}
extern "C" int emu_exclusive_scan(const int64_t *in, int64_t n, int64_t *out) {
    std::vector<int64_t> scratch(n / 1024 + 2, 0);
    return dsk::launch_exclusive_scan(in, n, out, scratch.data(), nullptr);
}
extern "C" int emu_forest_query(const uint32_t *sig, const int32_t *order, int64_t n, int K, int l, int k,
                                const uint32_t *qsig, int64_t nq, int topk, int32_t *out, int sm_count) {
    return dsk::launch_forest_query(sig, order, n, K, l, k, qsig, nq, topk, out, sm_count, nullptr);
}
extern "C" int emu_jaccard_pairs(const uint32_t *sig, int64_t n_rows, int k, const int64_t *ia, const int64_t *ib, int64_t m,
                                 int32_t *out) {
    return dsk::launch_jaccard_pairs(sig, n_rows, k, ia, ib, m, out, 1, nullptr);
}
extern "C" int emu_jaccard_topk(const uint32_t *q, int64_t nq, const uint32_t *db, int64_t n, int k, int topk,
                                int64_t self_base, int32_t *out_cnt, int64_t *out_idx, int sm_count) {
    return dsk::launch_jaccard_topk(q, nq, db, n, k, topk, self_base, out_cnt, out_idx, sm_count, nullptr);
}
// the fingerprint-prefilter variant (its own workspace for the bit planes), any n
extern "C" int emu_jaccard_topk_pf(const uint32_t *q, int64_t nq, const uint32_t *db, int64_t n, int k, int topk,
                                   int64_t self_base, int32_t *out_cnt, int64_t *out_idx, int sm_count) {
    const int words = (k + 31) / 32;
    const size_t qt = (size_t)((nq + dsk::kTQ - 1) / dsk::kTQ), dt = (size_t)((n + dsk::kTD - 1) / dsk::kTD);
    std::vector<uint32_t> ws((qt * dsk::kTQ + dt * dsk::kTD) * (size_t)dsk::kPfB * words + 8, 0xCDCDCDCDu);
    uint32_t *wsp = ws.data() + ((16 - (reinterpret_cast<uintptr_t>(ws.data()) & 15)) % 16) / 4;
    return dsk::launch_jaccard_topk_pf(q, nq, db, n, k, topk, self_base, out_cnt, out_idx, wsp, sm_count, nullptr);
}
extern "C" int emu_sha1_tokens(const uint8_t *bytes, const int64_t *off, int64_t n_tok, void *out, int out_is_u64) {
    return dsk::launch_sha1_tokens(bytes, off, n_tok, out, out_is_u64, 1, nullptr);
}
extern "C" int emu_hash_tokens(const uint8_t *bytes, const int64_t *off, int64_t n_tok, int kind, uint32_t seed, uint32_t *out) {
    return dsk::launch_hash_tokens(bytes, off, n_tok, kind, seed, out, 1, nullptr);
}
// rs / ln_cs / betas: [ss][dim] as the generator holds them; transposed to [dim][ss_pad] like dsk_wmh_create
extern "C" int emu_wmh(const float *rs, const float *ln_cs, const float *betas, int ss, int dim, const float *v, int64_t n,
                       int64_t *out, int32_t *status, int flags) {
    const int many = flags & 1, input_log = (flags >> 1) & 1;   // DSK_WMH_MINHASH_MANY | DSK_WMH_INPUT_LOG
    const int ss_pad = (ss + 31) / 32 * 32;
    std::vector<float> par((size_t)3 * dim * ss_pad);
    const float *src[3] = {rs, ln_cs, betas};
    for (int i = 0; i < 3; ++i)
        if (dsk::launch_wmh_transpose(src[i], ss, dim, ss_pad, par.data() + (size_t)i * dim * ss_pad, nullptr)) return -1;
    const size_t plane = (size_t)dim * ss_pad;
    return dsk::launch_wmh(par.data(), par.data() + plane, par.data() + 2 * plane, ss, ss_pad, dim, v, n, out, status, many,
                           input_log, 2, nullptr);
}
