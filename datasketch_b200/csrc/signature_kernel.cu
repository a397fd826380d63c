// signature_kernel.cu -- the two-phase signature build (default kernel of dsk_minhash_bulk*), sm_100a.
//
// Replaces datasketch/minhash.py:294-297 (numpy) / :281-291 (CuPy) batched over documents as MinHash.bulk
// does (:464-522).  Arithmetic and exactness argument: DESIGN.md sections 2-3; in short
//   x = (a*h + b) mod 2^64,  r = lo32(x) + top3(x)      (exact when dsk_perm_create found no unsafe permutation)
//   L' = lo32(a_lo*h + b_lo + 7)  (one IMAD)  and  r = L' - 7 + top3(x)  lies in [L'-7, L'],
// so only tokens whose L' is within 7 of a document's smallest L' can carry the minimum.
//
// Structure per warp (lane <-> P consecutive permutations, their (a, b) in registers):
//   stream   units of <= 32 consecutive documents come from a global atomic counter; the unit's token range is
//            streamed through a per-warp ring of four 1 KB slots by 1-D TMA bulk copies (cp.async.bulk +
//            mbarrier).  The ring is circular in ABSOLUTE token position: token g lives at ring[g mod 1024].
//   stage    a document is handled in sub-pieces of <= 512 tokens.  A sub-piece that is 16-byte aligned, a
//            multiple of 16 tokens long and does not wrap the ring is used in place; any other one is copied
//            (and re-aligned, padded to a multiple of 16 with a duplicate of its last token -- min is idempotent)
//            into a per-warp line buffer by the lanes.  When the warp has seen repeated tokens it also DEDUPLICATES
//            during that copy (a 1024-slot shared-memory table, one atomicExch per token, no probing: the first
//            copy of a token always survives, most later copies are dropped): the signature of a multiset is the
//            signature of its support, and distinct tokens cannot tie.
//   phase 1  per 16-token block: 16 IMAD (L') + 8 VIMNMX3 per permutation, then four tracking ops on
//            key = (block min & ~31) | block index:  m = smallest key, m2 = second smallest.
//   phase 2  per permutation: recompute L' on the winning block (block index = m & 31), find the one 4-token group
//            inside the +7 window and evaluate those 4 tokens exactly (groups that hold padding only are ignored).
//   long     with a caller-provided piece table (dsk_minhash_bulk_ws) a document longer than 4096 tokens is not processed
//            where it is met: its row gets the initial value and ceil(len / 1024) piece descriptors are appended; a
//            second launch (template flag PIECES) spreads the pieces over all warps and min-merges with atomicMin.
//   general  (template GEN) the same kernel for inputs the formula above does not cover.  For ANY 64-bit x,
//            (x mod p) & (2^32-1) = lo32(x) + top3(x) + s  (mod 2^32),  s = 1 iff `% p` takes its conditional subtract
//            ((x & p) + top3(x) >= p; -p = +1 mod 2^32), so r lies in [L'-8, L'] with L' = lo32(a_lo*h_lo + b_lo + 8): phase 1 is
//            unchanged (one IMAD on the LOW words) and only the exact stage evaluates the 64-bit form.  GEN = 1: u32 tokens
//            with permutations that can reach the subtract (user-supplied ones); GEN = 2: u64 tokens (hash values up to
//            2^64-1, minhash.py:294) -- staged as a low-word plane (phase 1) and a high-word plane (exact stage only).
//   lsh      (template LSH) fused MinHashLSH insert: the warp that finishes a document puts the row into its line buffer,
//            lane j fingerprints band j and does the bucket update (atomicCAS claim + atomicExch chain head + link store,
//            as lsh_insert_kernel) -- the tables' DRAM latency hides behind the other warps' integer work, and the
//            signature matrix is never read back.  Replaces lsh.py:326-347 for a whole batch of token lists.
//   flagged  a permutation with another block inside the window (m2 - m <= 69), a second group inside the window,
//            or a minimum so small that L'-7 could wrap (m < 32) is resolved by the whole warp, two permutations
//            at a time: every lane filters 1/32 of the sub-piece's tokens with L' and evaluates r exactly for the
//            ones inside the window, then one redux.sync.min.  Exact for any input; ~55 issue slots per flagged
//            (document, permutation).  A sub-piece with >= 8 flagged permutations (repeated tokens) switches the
//            warp to de-duplicating staging, which removes the ties at their source.
#include "dsk_common.cuh"

namespace dsk {

constexpr int kSigWarps = 4;        // warps per CTA
constexpr int kRingTok = 1024;      // tokens in a warp's ring
constexpr int kChunkShift = 8;      // 256 tokens = 1 KB per TMA bulk copy / ring slot
constexpr int kChunkTok = 1 << kChunkShift;
constexpr int kRingSlots = kRingTok / kChunkTok;
constexpr int kSubTok = 512;        // tokens per sub-piece
constexpr int kTabSlots = 1024;     // de-duplication table: one slot per hashed token value, later tokens overwrite earlier ones
constexpr uint32_t kEmptySlot = 0xFFFFFFFFu;
constexpr uint32_t kKeyMask = 31u;  // low bits of a tracking key hold the block index (32 blocks of 16 tokens)
template <int GEN> constexpr uint32_t kWin = GEN ? 8u : 7u;          // r lies in [L' - kWin, L']
template <int GEN> constexpr uint32_t kNearWindow = kWin<GEN> + 2u * kKeyMask;  // m2 - m <= this: another block may be inside the window
constexpr int kHiOff = kSubTok / 2;  // GEN = 2: high words of a staged sub-piece (<= 256 tokens) live at buf[kHiOff + i]
constexpr int kDedupeOnFlags = 8;   // flagged permutations in one sub-piece (of 32*P) that switch de-duplication on

// Path counters for the CPU emulation tests (tests/emu): which staging path / how many flagged permutations.
// The product build compiles them away.
#ifdef DSK_EMU
enum { STAT_IN_PLACE = 0, STAT_COPIED, STAT_DEDUPED, STAT_REMOVED, STAT_FLAGGED, STAT_COUNT };
std::atomic<long long> g_sig_stat[STAT_COUNT];
#define DSK_SIG_STAT(i, n) do { if (lane == 0) g_sig_stat[i].fetch_add((n), std::memory_order_relaxed); } while (0)
#else
#define DSK_SIG_STAT(i, n) do { } while (0)
#endif

static_assert(kRingSlots == 4, "ring = 4 slots indexed by absolute chunk number & 3");
static_assert(kSubTok / 16 <= (int)kKeyMask + 1, "block index must fit the key's low bits");

// r = lo32(x) + top3(x)
__device__ __forceinline__ uint32_t sig_eval(uint32_t alo, uint32_t ahi, uint64_t b, uint32_t h) {
    uint64_t x = (uint64_t)alo * h + b;                     // IMAD.WIDE.U32
    uint32_t xh = (uint32_t)(x >> 32) + ahi * h;            // IMAD
    return (uint32_t)x + (xh >> 29);                        // LEA.HI
}

// any x: r = ((x mod 2^64) % (2^61-1)) & (2^32-1) with 64-bit tokens and the conditional subtract (general variants)
__device__ __forceinline__ uint32_t sig_eval_gen(uint32_t alo, uint32_t ahi, uint64_t b, uint32_t hlo, uint32_t hhi) {
    const uint64_t a = ((uint64_t)ahi << 32) | alo, h = ((uint64_t)hhi << 32) | hlo;
    const uint64_t x = a * h + b;
    const uint64_t p = (1ull << 61) - 1;
    uint64_t y = (x & p) + (x >> 61);
    if (y >= p) y -= p;
    return (uint32_t)y;
}

__device__ __forceinline__ void load_block16(const uint32_t *src, uint32_t (&t)[16]) {
    const uint4 *q = reinterpret_cast<const uint4 *>(src);   // four broadcast LDS.128
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        uint4 v = q[i];
        t[4 * i] = v.x; t[4 * i + 1] = v.y; t[4 * i + 2] = v.z; t[4 * i + 3] = v.w;
    }
}

// A long document's pieces (2^piece_shift tokens each) go to the table the second launch works through.
#ifdef DSK_EMU
static inline
#elif DSK_SIG_APPEND_NOINLINE
__device__ __noinline__
#else
__device__ __forceinline__
#endif
void append_pieces(unsigned *piece_hdr, PieceDesc *pieces, int piece_shift, int64_t row, int64_t start, int64_t end, int lane) {
    const int64_t step = (int64_t)1 << piece_shift;
    const int64_t np = (end - start + step - 1) >> piece_shift;
    unsigned base = 0;
    if (lane == 0) base = atomicAdd(piece_hdr, (unsigned)np);
    base = __shfl_sync(0xFFFFFFFFu, base, 0);
    for (int64_t q = lane; q < np; q += 32) {
        PieceDesc pd;
        pd.row = row; pd.start = start + (q << piece_shift);
        pd.end = min(end, pd.start + step); pd.reserved = 0;
        pieces[base + q] = pd;
    }
}

// Code-generation switches.  The kernel's speed moves by several per cent with changes that do not touch the block loop
// (ptxas schedules / allocates the whole kernel at once), so the combination in use was picked by timing builds of the same
// sources with different -D values on one box (tools/build_variants.sh + tools/ab_libs.py, profiles/r2r_kernel_codegen_grid.txt:
// 16 combinations span 2.94 .. 3.14 ms on the C2 shape; the defaults below are the fastest one).
#ifndef DSK_SIG_PEND
#define DSK_SIG_PEND 1          // fold a block's tracking ops in during the next block's IMADs (P <= 4 only)
#endif
#ifndef DSK_SIG_G
#define DSK_SIG_G 4             // permutations handled together in phase 2 (P <= 4)
#endif
#ifndef DSK_SIG_WSEL
#define DSK_SIG_WSEL 0          // 0: re-read the winning group through a pointer; 1: select it from registers
#endif
#ifndef DSK_SIG_APPEND_NOINLINE
#define DSK_SIG_APPEND_NOINLINE 1   // the long-document piece append as an out-of-line call
#endif
// K > 128 (8 permutations per lane): CTAs per SM, software pipelining of the block loop, folded tracking ops
#ifndef DSK_SIG_OCC8
#define DSK_SIG_OCC8 3
#endif
#ifndef DSK_SIG_SWP8
#define DSK_SIG_SWP8 1
#endif
#ifndef DSK_SIG_PEND8
#define DSK_SIG_PEND8 1
#endif
template <int P> constexpr bool kPend = DSK_SIG_PEND && (P <= 4 || DSK_SIG_PEND8);

template <int P, int OCC, bool PIECES, int GEN, bool LSH = false>
__global__ void __launch_bounds__(kSigWarps * 32, OCC) minhash_sig_kernel(const BulkParams prm) {
    static_assert(!(LSH && PIECES), "the fused insert needs whole rows: no piece mode");
    constexpr int TW = GEN == 2 ? 2 : 1;            // 32-bit words per token; the ring and its copies count WORDS
    constexpr int SUB = GEN == 2 ? kHiOff : kSubTok;  // tokens per sub-piece
    constexpr uint32_t W = kWin<GEN>;
    __shared__ __align__(128) uint32_t s_ring[kSigWarps][kRingTok];
    __shared__ __align__(128) uint32_t s_buf[kSigWarps][kSubTok];
    __shared__ __align__(16) uint32_t s_tab[kSigWarps][kTabSlots];
    __shared__ __align__(8) uint64_t s_bar[kSigWarps][kRingSlots];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t *__restrict__ tokens = static_cast<const uint32_t *>(prm.tokens);
    const int64_t *__restrict__ offsets = prm.offsets;
    const int K = prm.k;
    const int kl = blockIdx.y * (32 * P) + lane * P;  // first permutation owned by this lane
    // piece mode (second launch, long documents): the work items are the PieceDesc entries the first launch appended
    const int64_t n_docs = PIECES ? (int64_t)*reinterpret_cast<const volatile unsigned *>(prm.piece_hdr) : prm.n_docs;
    const int64_t n_tokens = prm.n_tokens;
    unsigned *const work_counter = PIECES ? prm.piece_hdr + 64 : prm.work_counter;
    const int docs_per_unit = PIECES ? 1 : prm.docs_per_unit;

    // c7 = b_lo + 7 is what the hot loop adds; b_lo = c7 - 7 is rebuilt where the exact evaluation needs it
    uint32_t alo[P], ahi[P], c7[P], bhi[P];
#pragma unroll
    for (int j = 0; j < P; ++j) {
        alo[j] = __ldg(prm.a_lo + kl + j); ahi[j] = __ldg(prm.a_hi + kl + j);
        // the table holds b_lo + 7 as its own plane: computed here, ptxas re-adds the 7 in every loop iteration
        c7[j] = __ldg((GEN ? prm.b_lo8 : prm.b_lo7) + kl + j); bhi[j] = __ldg(prm.b_hi + kl + j);   // c7 = b_lo + W
    }

    uint32_t *const ring = s_ring[warp];
    uint32_t *const buf = s_buf[warp];
    uint32_t *const tab = s_tab[warp];
    uint64_t *const bar = s_bar[warp];
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < kRingSlots; ++i) mbar_init(&bar[i], 1);
        fence_mbar_init();
    }
    __syncwarp();
    uint32_t par_mask = 0;  // bit s = phase parity the next wait on slot s must use

    // bulk copies move 16-byte granules: the last (n_tokens mod 4) tokens of the array are patched in by lanes
    const int64_t n_words = n_tokens * TW;
    const int64_t copy_end_tok = n_words & ~(int64_t)3;
    const int64_t tail_chunk = (copy_end_tok < n_words) ? (copy_end_tok >> kChunkShift) : -1;

    bool dedupe = false;     // warp-uniform: copy stage removes repeated tokens
    int clean_run = 0;       // consecutive deduplicated sub-pieces in which nothing was removed

    while (true) {
        int64_t unit = 0;
        if (lane == 0) unit = (int64_t)atomicAdd(work_counter + blockIdx.y, 1u);
        unit = __shfl_sync(0xFFFFFFFFu, unit, 0);
        const int64_t dlo = unit * docs_per_unit;
        if (dlo >= n_docs) break;
        const int64_t dhi = min(dlo + (int64_t)docs_per_unit, n_docs);
        int64_t tok_lo, tok_hi, piece_row = 0;
        if constexpr (PIECES) {
            const PieceDesc pd = prm.pieces[dlo];
            tok_lo = pd.start; tok_hi = pd.end; piece_row = pd.row;
        } else {
            tok_lo = __ldg(offsets + dlo); tok_hi = __ldg(offsets + dhi);
        }
        // ring positions below are WORD positions (= token positions unless tokens are 64-bit)
        int64_t c_next = (tok_lo * TW) >> kChunkShift;          // next chunk to issue
        int64_t c_wait = c_next;                                // next chunk to wait for
        const int64_t c_last = (tok_hi * TW - 1) >> kChunkShift;  // last chunk this unit touches (if it has tokens)

        // consume the completions of copies that were issued for chunks < upto but never needed (skipped tokens): every
        // issued copy must be waited for exactly once, or the slot's phase parity goes out of step
        auto drain = [&](int64_t upto) {
            while (c_wait < upto && c_wait < c_next) {
                const int slot = (int)(c_wait & (kRingSlots - 1));
                mbar_wait(&bar[slot], (par_mask >> slot) & 1u);
                par_mask ^= 1u << slot;
                ++c_wait;
            }
        };
        // make tokens [s, e) resident in the ring: chunk c lives in slot c & 3.  Chunks before s's chunk are dead (every
        // earlier sub-piece is finished), so up to three chunks beyond it can be in flight.
        auto ensure = [&](int64_t s, int64_t e) {
            const int64_t cf = s >> kChunkShift, cl = (e - 1) >> kChunkShift;
            if (c_wait < cf) {   // tokens were skipped (a deferred long document): drain what is in flight, jump ahead
                drain(cf);
                c_wait = cf;
                if (c_next < cf) c_next = cf;
            }
            const int64_t lim = min(cf + (kRingSlots - 1), c_last);
            if (c_next <= lim) {
                __syncwarp();   // every lane is done reading the chunks these copies overwrite (in-place sub-pieces)
                if (lane == 0) {
                    for (int64_t c = c_next; c <= lim; ++c) {
                        const int slot = (int)(c & (kRingSlots - 1));
                        const int64_t t0 = c << kChunkShift;
                        const int64_t t1 = min(t0 + kChunkTok, copy_end_tok);
                        if (t1 > t0) {
                            const uint32_t bytes = (uint32_t)(t1 - t0) * 4u;
                            mbar_arrive_expect_tx(&bar[slot], bytes);
                            bulk_g2s(ring + slot * kChunkTok, tokens + t0, bytes, &bar[slot]);
                        } else {
                            mbar_arrive(&bar[slot]);
                        }
                    }
                }
                c_next = lim + 1;
            }
            while (c_wait <= cl) {
                const int slot = (int)(c_wait & (kRingSlots - 1));
                mbar_wait(&bar[slot], (par_mask >> slot) & 1u);
                par_mask ^= 1u << slot;
                if (c_wait == tail_chunk) {     // the < 16-byte tail of the whole array cannot travel by bulk copy
                    const int64_t g = copy_end_tok + lane;
                    if (g < n_words) ring[g & (kRingTok - 1)] = __ldg(tokens + g);
                    fence_proxy_async();        // generic-proxy write, later bulk copies reuse the slot
                    __syncwarp();
                }
                ++c_wait;
            }
        };

        int64_t start = tok_lo;
        int64_t end_pref = PIECES ? tok_hi : __ldg(offsets + dlo + 1);  // offsets are read one document ahead
        for (int64_t d = dlo; d < dhi; ++d) {
            const int64_t end = end_pref;
            if (!PIECES && d + 1 < dhi) end_pref = __ldg(offsets + d + 2);

            uint32_t acc[P];
#pragma unroll
            for (int j = 0; j < P; ++j) acc[j] = 0xFFFFFFFFu;

            // a long document is cut into pieces for the second launch; here only its initial row gets stored below
            bool defer = false;
            if constexpr (!PIECES) {
                defer = prm.long_doc_tokens > 0 && end - start > prm.long_doc_tokens;
                if (defer && blockIdx.y == 0) append_pieces(prm.piece_hdr, prm.pieces, prm.piece_shift, d, start, end, lane);
            }

            for (int64_t s = start; s < (defer ? start : end); s += SUB) {
                const int len = (int)min((int64_t)SUB, end - s);
                ensure(s * TW, (s + len) * TW);
                const uint32_t rpos = (uint32_t)(s * TW) & (kRingTok - 1);

                // ---- stage: in place, or copy / re-align / pad / deduplicate into the line buffer ----------------
                const uint32_t *src;
                int n_eff = len;
                const bool in_place = TW == 1 && !dedupe && (len & 15) == 0 && (rpos & 3u) == 0 && rpos + (uint32_t)len <= (uint32_t)kRingTok;
                if (in_place) {
                    src = ring + rpos;
                    DSK_SIG_STAT(STAT_IN_PLACE, 1);
                } else {
                    __syncwarp();   // every lane is done with the previous contents of buf / tab
                    if (!dedupe) {
                        if constexpr (TW == 2) {   // de-interleave: low words -> buf[i] (phase 1), high words -> buf[kHiOff + i]
                            const uint2 *r2 = reinterpret_cast<const uint2 *>(ring);
                            for (int i = lane; i < len; i += 32) {
                                const uint2 t = r2[((rpos >> 1) + (uint32_t)i) & (kRingTok / 2 - 1)];
                                buf[i] = t.x; buf[kHiOff + i] = t.y;
                            }
                        } else {
                            for (int i = lane; i < len; i += 32) buf[i] = ring[(rpos + (uint32_t)i) & (kRingTok - 1)];
                        }
                        DSK_SIG_STAT(STAT_COPIED, 1);
                    } else {
                        uint4 *t4 = reinterpret_cast<uint4 *>(tab);
#pragma unroll
                        for (int q = 0; q < kTabSlots / 128; ++q)
                            t4[q * 32 + lane] = make_uint4(kEmptySlot, kEmptySlot, kEmptySlot, kEmptySlot);
                        __syncwarp();
                        // One atomicExch per token, no probing: a token is dropped iff the exchange returns the token
                        // itself, i.e. an earlier copy of it already went through this slot -- so the first copy always
                        // survives (exact), while a copy whose slot was meanwhile taken by another token survives too
                        // (a missed repeat only costs a flagged permutation later).  The marker value itself is kept.
                        int base = 0;
                        for (int i0 = 0; i0 < len; i0 += 32) {
                            const int i = i0 + lane;
                            bool keep = i < len;
                            uint32_t t = kEmptySlot, thi = kEmptySlot;
                            if constexpr (TW == 2) {
                                // 64-bit tokens: the slot holds the POSITION of the last token that went through it (a 32-bit
                                // exchange -- 64-bit shared-memory atomics are far slower), identity = the whole token, read
                                // back from the ring: dropped iff the token at the returned position equals this one
                                if (keep) {
                                    const uint2 *r2 = reinterpret_cast<const uint2 *>(ring);
                                    const uint2 t2 = r2[((rpos >> 1) + (uint32_t)i) & (kRingTok / 2 - 1)];
                                    t = t2.x; thi = t2.y;
                                    const uint32_t prev = atomicExch(&tab[((t ^ (thi * 0x85EBCA6Bu)) * 0x9E3779B1u) >> 22], (uint32_t)i);
                                    if (prev != kEmptySlot) {
                                        const uint2 o2 = r2[((rpos >> 1) + prev) & (kRingTok / 2 - 1)];
                                        keep = o2.x != t || o2.y != thi;
                                    }
                                }
                            } else {
                                if (keep) t = ring[(rpos + (uint32_t)i) & (kRingTok - 1)];
                                if (t != kEmptySlot) keep = atomicExch(&tab[(t * 0x9E3779B1u) >> 22], t) != t;
                            }
                            const unsigned bal = __ballot_sync(0xFFFFFFFFu, keep);
                            const int at = base + __popc(bal & ((1u << lane) - 1u));
                            if (keep) buf[at] = t;
                            if constexpr (TW == 2) { if (keep) buf[kHiOff + at] = thi; }
                            base += __popc(bal);
                        }
                        n_eff = base;
                        DSK_SIG_STAT(STAT_DEDUPED, 1);
                        DSK_SIG_STAT(STAT_REMOVED, len - n_eff);
                        clean_run = (n_eff == len) ? clean_run + 1 : 0;
                        if (clean_run >= 16) { dedupe = false; clean_run = 0; }
                        __syncwarp();   // buf[n_eff - 1] below was written by another lane
                    }
                    const int pad = (16 - (n_eff & 15)) & 15;
                    if constexpr (TW == 2) {
                        __syncwarp();   // buf[n_eff - 1] was written by another lane
                        if (lane < pad) { buf[n_eff + lane] = buf[n_eff - 1]; buf[kHiOff + n_eff + lane] = buf[kHiOff + n_eff - 1]; }
                    } else {
                        if (lane < pad) buf[n_eff + lane] = dedupe || n_eff != len ? buf[n_eff - 1]
                                                                                   : ring[(rpos + (uint32_t)len - 1u) & (kRingTok - 1)];
                    }
                    __syncwarp();
                    src = buf;
                }
                const int nblk = (n_eff + 15) >> 4;
                const int ngrp = (n_eff + 3) >> 2;   // 4-token groups holding at least one real token

                // ---- phase 1: m = smallest key, m2 = second smallest, key = (block min L' & ~31) | block -------
                // kPend: the tracking ops of a block are folded in while the NEXT block's IMADs issue (pend = the previous
                // block's key; 2^32-1 is a no-op), so that no run of 16 non-multiply instructions ends each block.
                uint32_t m[P], m2[P], pend[kPend<P> ? P : 1];
#pragma unroll
                for (int j = 0; j < P; ++j) { m[j] = 0xFFFFFFFFu; m2[j] = 0xFFFFFFFFu; }
                if constexpr (kPend<P>) {
#pragma unroll
                    for (int j = 0; j < P; ++j) pend[j] = 0xFFFFFFFFu;
                }
                auto fold = [&](uint32_t key, int j) {
                    const uint32_t om = m[j];
                    m2[j] = min(m2[j], max(key, om));
                    m[j] = min(om, key);
                };
                auto compute = [&](const uint32_t (&t)[16], uint32_t lb) {
#pragma unroll
                    for (int j = 0; j < P; ++j) {
                        const uint32_t c = c7[j];
                        uint32_t bm = umin3(alo[j] * t[0] + c, alo[j] * t[1] + c, alo[j] * t[2] + c);
                        if constexpr (kPend<P>) fold(pend[j], j);
#pragma unroll
                        for (int i = 3; i < 15; i += 2) bm = umin3(bm, alo[j] * t[i] + c, alo[j] * t[i + 1] + c);
                        bm = min(bm, alo[j] * t[15] + c);
                        const uint32_t key = (bm & ~kKeyMask) | lb;
                        if constexpr (kPend<P>) pend[j] = key;
                        else fold(key, j);
                    }
                };
                {
                    const uint32_t *q = src;
                    const uint32_t *const qe = src + nblk * 16;
                    uint32_t lb = 0;
                    if constexpr (P > 4 && !DSK_SIG_SWP8) {
#pragma unroll 1
                        for (; q < qe; q += 16, ++lb) {
                            uint32_t t[16];
                            load_block16(q, t);
                            compute(t, lb);
                        }
                    } else {
                        // software pipeline over two register buffers: the next block's LDS.128 are in flight while
                        // the current block's IMADs issue
                        uint32_t ta[16], tb[16];
                        load_block16(q, ta);
#pragma unroll 1
                        while (true) {
                            const uint32_t *q1 = q + 16;
                            if (q1 < qe) load_block16(q1, tb);
                            compute(ta, lb);
                            if (q1 >= qe) break;
                            const uint32_t *q2 = q1 + 16;
                            if (q2 < qe) load_block16(q2, ta);
                            compute(tb, lb + 1);
                            if (q2 >= qe) break;
                            q = q2;
                            lb += 2;
                        }
                    }
                }

                if constexpr (kPend<P>) {
#pragma unroll
                    for (int j = 0; j < P; ++j) fold(pend[j], j);   // the last block's key
                }

                // ---- phase 2: exact evaluation inside each permutation's winning block ------------------------
                uint32_t res[P], win[P];   // win: upper end of the L' window the flagged path must evaluate
                unsigned need_slow = 0;
                constexpr int G = P > 4 ? 4 : (P < DSK_SIG_G ? P : DSK_SIG_G);   // permutations handled together
#pragma unroll
                for (int j0 = 0; j0 < P; j0 += G) {
                    uint4 v[G][4];
#pragma unroll
                    for (int g = 0; g < G; ++g) {
                        const uint4 *wb = reinterpret_cast<const uint4 *>(src + (m[j0 + g] & kKeyMask) * 16u);
#pragma unroll
                        for (int i = 0; i < 4; ++i) v[g][i] = wb[i];
                    }
                    const uint4 *wsrc[G];   // the one group inside the window (DSK_SIG_WSEL 0: re-read below; 1: register selects)
                    uint4 w[G];
#pragma unroll
                    for (int g = 0; g < G; ++g) {
                        const int j = j0 + g;
                        const uint32_t c = c7[j];
                        uint32_t gm[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            gm[i] = min(umin3(alo[j] * v[g][i].x + c, alo[j] * v[g][i].y + c, alo[j] * v[g][i].z + c),
                                        alo[j] * v[g][i].w + c);
                        const uint32_t thr = min(umin3(gm[0], gm[1], gm[2]), gm[3]) + W;   // a wrap implies m2 - m <= window
                        // groups made of padding only (last block) repeat a real token: they never count as a second group
                        const int gv = ngrp - (int)(m[j] & kKeyMask) * 4;
                        const bool in0 = gm[0] <= thr, in1 = gm[1] <= thr && gv > 1, in2 = gm[2] <= thr && gv > 2,
                                   in3 = gm[3] <= thr && gv > 3;
                        const int nin = (int)in0 + (int)in1 + (int)in2 + (int)in3;
                        // another block or another group inside the window, or L'-7 may wrap: the warp resolves it below
                        if (nin != 1 || m[j] <= kKeyMask || (m2[j] - m[j]) <= kNearWindow<GEN>) need_slow |= 1u << j;
                        win[j] = (m[j] <= kKeyMask || thr < W) ? 0xFFFFFFFFu : thr;
                        if constexpr (DSK_SIG_WSEL && GEN != 2) w[g] = in0 ? v[g][0] : in1 ? v[g][1] : in2 ? v[g][2] : v[g][3];
                        else wsrc[g] = reinterpret_cast<const uint4 *>(src + (m[j] & kKeyMask) * 16u) + (in0 ? 0 : in1 ? 1 : in2 ? 2 : 3);
                    }
                    if constexpr (!DSK_SIG_WSEL || GEN == 2) {
#pragma unroll
                        for (int g = 0; g < G; ++g) w[g] = *wsrc[g];
                    }
#pragma unroll
                    for (int g = 0; g < G; ++g) {
                        const int j = j0 + g;
                        const uint64_t b64 = ((uint64_t)bhi[j] << 32) | (c7[j] - W);
                        if constexpr (GEN) {   // full `% p` (conditional subtract); GEN = 2: the group's high words from the second plane
                            uint4 wh = make_uint4(0u, 0u, 0u, 0u);
                            if constexpr (GEN == 2) wh = wsrc[g][kHiOff / 4];
                            res[j] = min(umin3(sig_eval_gen(alo[j], ahi[j], b64, w[g].x, wh.x), sig_eval_gen(alo[j], ahi[j], b64, w[g].y, wh.y),
                                               sig_eval_gen(alo[j], ahi[j], b64, w[g].z, wh.z)), sig_eval_gen(alo[j], ahi[j], b64, w[g].w, wh.w));
                        } else {
                            res[j] = min(umin3(sig_eval(alo[j], ahi[j], b64, w[g].x), sig_eval(alo[j], ahi[j], b64, w[g].y),
                                               sig_eval(alo[j], ahi[j], b64, w[g].z)), sig_eval(alo[j], ahi[j], b64, w[g].w));
                        }
                    }
                }

                // ---- flagged permutations: resolved by the whole warp, two at a time ---------------------------------
                // Every lane filters 1/32 of the sub-piece's tokens with the one-IMAD value L' and evaluates exactly the
                // ones inside the window [.., win]: a token with L' > (winning block's min L') + 7 cannot hold the
                // minimum (r >= L' - 7).  win = 2^32-1 (evaluate everything) when L' - 7 may wrap.
                if (__any_sync(0xFFFFFFFFu, need_slow != 0)) {
                    int nflag = 0;
                    const int n_pad = nblk * 16;
#pragma unroll
                    for (int j = 0; j < P; ++j) {
                        unsigned bal = __ballot_sync(0xFFFFFFFFu, (need_slow >> j) & 1u);
                        nflag += __popc(bal);
                        while (bal) {
                            const int o0 = __ffs(bal) - 1;
                            bal &= bal - 1;
                            const int o1 = bal ? __ffs(bal) - 1 : o0;    // second owner (or the first one again)
                            bal &= bal - 1;
                            const uint32_t a0 = __shfl_sync(0xFFFFFFFFu, alo[j], o0), a1 = __shfl_sync(0xFFFFFFFFu, alo[j], o1);
                            const uint32_t h0 = __shfl_sync(0xFFFFFFFFu, ahi[j], o0), h1 = __shfl_sync(0xFFFFFFFFu, ahi[j], o1);
                            const uint32_t c0 = __shfl_sync(0xFFFFFFFFu, c7[j], o0), c1 = __shfl_sync(0xFFFFFFFFu, c7[j], o1);
                            const uint32_t g0 = __shfl_sync(0xFFFFFFFFu, bhi[j], o0), g1 = __shfl_sync(0xFFFFFFFFu, bhi[j], o1);
                            const uint32_t w0 = __shfl_sync(0xFFFFFFFFu, win[j], o0), w1 = __shfl_sync(0xFFFFFFFFu, win[j], o1);
                            const uint64_t b0 = ((uint64_t)g0 << 32) | (c0 - W), b1 = ((uint64_t)g1 << 32) | (c1 - W);
                            uint32_t r0 = 0xFFFFFFFFu, r1 = 0xFFFFFFFFu;
#pragma unroll 4
                            for (int i = lane; i < n_pad; i += 32) {
                                const uint32_t t = src[i];
                                const bool k0 = a0 * t + c0 <= w0, k1 = a1 * t + c1 <= w1;
                                if (k0 || k1) {
                                    if constexpr (GEN) {
                                        const uint32_t thi = GEN == 2 ? src[kHiOff + i] : 0u;
                                        if (k0) r0 = min(r0, sig_eval_gen(a0, h0, b0, t, thi));
                                        if (k1) r1 = min(r1, sig_eval_gen(a1, h1, b1, t, thi));
                                    } else {
                                        if (k0) r0 = min(r0, sig_eval(a0, h0, b0, t));
                                        if (k1) r1 = min(r1, sig_eval(a1, h1, b1, t));
                                    }
                                }
                            }
                            r0 = __reduce_min_sync(0xFFFFFFFFu, r0);
                            r1 = __reduce_min_sync(0xFFFFFFFFu, r1);
                            if (lane == o1) res[j] = r1;
                            if (lane == o0) res[j] = r0;
                        }
                    }
                    DSK_SIG_STAT(STAT_FLAGGED, nflag);
                    // many ties = many repeated tokens: from now on this warp deduplicates while it stages
                    if (nflag >= kDedupeOnFlags && !dedupe) { dedupe = true; clean_run = 0; }
                }
#pragma unroll
                for (int j = 0; j < P; ++j) acc[j] = min(acc[j], res[j]);
            }

            // ---- merge with the running state (minhash.py:297) and store ---------------------------
            if (!PIECES && prm.init != nullptr) {
#pragma unroll
                for (int j = 0; j < P; ++j) {
                    if (kl + j < K) {
                        const int64_t e = d * prm.init_stride + kl + j;
                        if (prm.init_is_u64) {
                            const uint64_t v = __ldg(static_cast<const uint64_t *>(prm.init) + e);
                            acc[j] = (uint32_t)min((uint64_t)acc[j], v);  // acc <= 2^32-1 so the min fits
                        } else {
                            acc[j] = min(acc[j], __ldg(static_cast<const uint32_t *>(prm.init) + e));
                        }
                    }
                }
            }
            if constexpr (PIECES) {   // min-merge this piece into the row the first launch initialised
                if (prm.out_is_u64) {
                    unsigned long long *row = static_cast<unsigned long long *>(prm.out) + piece_row * (int64_t)K + kl;
#pragma unroll
                    for (int j = 0; j < P; ++j)
                        if (kl + j < K) atomicMin(row + j, (unsigned long long)acc[j]);
                } else {
                    uint32_t *row = static_cast<uint32_t *>(prm.out) + piece_row * (int64_t)K + kl;
#pragma unroll
                    for (int j = 0; j < P; ++j)
                        if (kl + j < K) atomicMin(row + j, acc[j]);
                }
                start = end;
                continue;
            }
            // One store per destination: the caller's matrix, or -- fused all-gather -- the same row of the full
            // [N_total, K] matrix on EVERY rank (peer pointers mapped over NVLink; plain st.global to a peer address).
            const int n_dst = prm.n_peers > 0 ? prm.n_peers : 1;
            for (int pd = 0; pd < n_dst; ++pd) {
                void *obase = prm.n_peers > 0 ? prm.peer_out[pd] : prm.out;
                const int64_t orow = d + (prm.n_peers > 0 ? prm.peer_row_offset : 0);
                if (prm.out_is_u64) {
                    uint64_t *row = static_cast<uint64_t *>(obase) + orow * (int64_t)K + kl;
#pragma unroll
                    for (int j = 0; j < P; ++j)
                        if (kl + j < K) row[j] = acc[j];
                } else {
                    uint32_t *row = static_cast<uint32_t *>(obase) + orow * (int64_t)K + kl;
                    if (P == 4 && (K & 3) == 0) {
                        if (kl < K) *reinterpret_cast<uint4 *>(row) = make_uint4(acc[0], acc[1 % P], acc[2 % P], acc[3 % P]);
                    } else if (P == 8 && (K & 7) == 0) {
                        if (kl < K) {
                            reinterpret_cast<uint4 *>(row)[0] = make_uint4(acc[0], acc[1 % P], acc[2 % P], acc[3 % P]);
                            reinterpret_cast<uint4 *>(row)[1] = make_uint4(acc[4 % P], acc[5 % P], acc[6 % P], acc[7 % P]);
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < P; ++j)
                            if (kl + j < K) row[j] = acc[j];
                    }
                }
            }
            // ---- fused LSH insert: row -> line buffer, lane <-> band (lsh_insert_kernel's update, one document) ----------
            if constexpr (LSH) {
                __syncwarp();   // every lane is done with the staged tokens of this document's last sub-piece
#pragma unroll
                for (int j = 0; j < P; ++j)
                    if (kl + j < K) buf[kl + j] = acc[j];
                __syncwarp();
                const uint64_t mask = (uint64_t)prm.lsh_cap_slots - 1;
                const int64_t doc = prm.lsh_doc0 + d;
                for (int band = lane; band < prm.lsh_b; band += 32) {
                    const uint64_t fp = band_fp(buf + band * prm.lsh_r, prm.lsh_r, band);
                    uint64_t *tabp = prm.lsh_slots + (int64_t)band * prm.lsh_cap_slots * 2;
                    uint64_t slot = lsh_mix64(fp) & mask;
                    while (true) {
                        const unsigned long long prev = atomicCAS(reinterpret_cast<unsigned long long *>(tabp + 2 * slot),
                                                                  (unsigned long long)kEmptyKey, (unsigned long long)fp);
                        if (prev == kEmptyKey || prev == fp) break;
                        slot = (slot + 1) & mask;
                    }
                    const int32_t old = atomicExch(reinterpret_cast<int32_t *>(tabp + 2 * slot + 1), (int32_t)doc);
                    prm.lsh_next[doc * prm.lsh_b + band] = old;
                }
                __syncwarp();   // the next document stages into the line buffer
            }
            start = end;
        }
        drain(c_next);   // (only a unit that ends in a deferred document has copies left in flight)
    }  // units
}

constexpr int64_t kLongDocTokens = kLongDocTokensApi;
constexpr int kPieceTokens = 1 << kPieceShiftApi;

// bytes of the piece table a launch over n_tokens tokens can need (0: no document can be long enough)
size_t minhash_sig_workspace_bytes(int64_t n_tokens) {
    if (n_tokens <= kLongDocTokens) return 0;
    const int64_t cap = n_tokens / kPieceTokens + n_tokens / kLongDocTokens + 2;
    return (size_t)kPieceHdrBytes + (size_t)cap * sizeof(PieceDesc);
}

template <int P, int OCC, int GEN>
static cudaError_t launch_sig(const BulkParams &prm_in, int sm_count, cudaStream_t s) {
    BulkParams prm = prm_in;
    const int slices = (prm.k + 32 * P - 1) / (32 * P);
    int64_t gx = (prm.n_docs + kSigWarps - 1) / kSigWarps;
    int64_t gmax = (int64_t)sm_count * OCC / slices;  // persistent: every CTA of every slice resident, one wave
    if (gmax < 1) gmax = 1;
    if (gx > gmax) gx = gmax;
    if (gx < 1) gx = 1;
    // unit size: ~8 units per warp for balance, at most 32 documents so a unit's ring restart is amortised
    if (prm.docs_per_unit <= 0) {   // 0 = choose here; > 0 = forced by the caller (tests)
        int64_t dpu = prm.n_docs / (gx * kSigWarps * 8);
        prm.docs_per_unit = (int)(dpu < 1 ? 1 : (dpu > 32 ? 32 : dpu));
    }
    // long documents: only with a caller-provided piece table, plain (non-gather) output
    const bool split = prm.piece_hdr != nullptr && prm.n_peers == 0 && prm.n_tokens > prm.long_doc_tokens && prm.long_doc_tokens > 0;
    if (!split) prm.long_doc_tokens = 0;
    cudaError_t e = cudaMemsetAsync(prm.work_counter, 0, sizeof(unsigned) * (size_t)slices, s);
    if (e != cudaSuccess) return e;
    if (split) {
        e = cudaMemsetAsync(prm.piece_hdr, 0, kPieceHdrBytes, s);
        if (e != cudaSuccess) return e;
    }
    dim3 grid((unsigned)gx, (unsigned)slices);
    DSK_LAUNCH((minhash_sig_kernel<P, OCC, false, GEN>), grid, kSigWarps * 32, 0, s, prm);
    e = cudaGetLastError();
    if (e != cudaSuccess || !split) return e;
    // second launch: the pieces (their number is only known on the device; an empty table costs one idle launch)
    dim3 pgrid((unsigned)gmax, (unsigned)slices);
    DSK_LAUNCH((minhash_sig_kernel<P, OCC, true, GEN>), pgrid, kSigWarps * 32, 0, s, prm);
    return cudaGetLastError();
}

// fused LSH insert: default variant only (u32 tokens, safe permutations), one K slice (the whole row in one warp), no pieces
template <int P, int OCC>
static cudaError_t launch_sig_lsh(const BulkParams &prm_in, int sm_count, cudaStream_t s) {
    BulkParams prm = prm_in;
    int64_t gx = (prm.n_docs + kSigWarps - 1) / kSigWarps;
    const int64_t gmax = (int64_t)sm_count * OCC;
    if (gx > gmax) gx = gmax;
    if (gx < 1) gx = 1;
    if (prm.docs_per_unit <= 0) {
        int64_t dpu = prm.n_docs / (gx * kSigWarps * 8);
        prm.docs_per_unit = (int)(dpu < 1 ? 1 : (dpu > 32 ? 32 : dpu));
    }
    prm.long_doc_tokens = 0;
    cudaError_t e = cudaMemsetAsync(prm.work_counter, 0, sizeof(unsigned), s);
    if (e != cudaSuccess) return e;
    DSK_LAUNCH((minhash_sig_kernel<P, OCC, false, 0, true>), dim3((unsigned)gx, 1u), kSigWarps * 32, 0, s, prm);
    return cudaGetLastError();
}

template <int GEN>
static cudaError_t launch_sig_k(const BulkParams &prm, int sm_count, cudaStream_t s) {
    if (prm.k <= 32) return launch_sig<1, 4, GEN>(prm, sm_count, s);
    if (prm.k <= 64) return launch_sig<2, 4, GEN>(prm, sm_count, s);
    if (prm.k <= 128) return launch_sig<4, 4, GEN>(prm, sm_count, s);
    // K > 128: 8 permutations per lane, or (DSK_SIG_WIDE=4, an A/B switch) K-slices of 128 on blockIdx.y with 4 per lane
    static const bool wide4 = [] { const char *e = getenv("DSK_SIG_WIDE"); return e && atoi(e) == 4; }();
    if (wide4) return launch_sig<4, 4, GEN>(prm, sm_count, s);
    return launch_sig<8, DSK_SIG_OCC8, GEN>(prm, sm_count, s);
}

// 4 CTAs (16 warps) per SM: 5 and 6 were measured slower (register cap, profiles/r2i_kernel_variants_ab.txt)
cudaError_t launch_minhash_sig(const BulkParams &prm, int sm_count, cudaStream_t s) {
    if (prm.lsh_slots != nullptr) {
        if (prm.gen != 0 || prm.k > 256 || prm.n_peers != 0 || prm.out_is_u64 || prm.lsh_b * prm.lsh_r > prm.k) return cudaErrorInvalidValue;
        if (prm.k <= 32) return launch_sig_lsh<1, 4>(prm, sm_count, s);
        if (prm.k <= 64) return launch_sig_lsh<2, 4>(prm, sm_count, s);
        if (prm.k <= 128) return launch_sig_lsh<4, 4>(prm, sm_count, s);
        return launch_sig_lsh<8, DSK_SIG_OCC8>(prm, sm_count, s);
    }
    switch (prm.gen) {
        case 0: return launch_sig_k<0>(prm, sm_count, s);
        case 1: return launch_sig_k<1>(prm, sm_count, s);
        case 2: return launch_sig_k<2>(prm, sm_count, s);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace dsk
