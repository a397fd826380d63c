// minhash_kernels.cu -- the permutation-hash-and-min path as hand-written sm_100a CUDA.
//
// Replaces datasketch/minhash.py:294-297 (numpy) / :281-291 (CuPy) batched over
// documents as MinHash.bulk does (:464-522).  See DESIGN.md "Kernel 1".
//
// Work decomposition
//   warp  <-> units of up to 32 consecutive documents pulled from a global atomic counter
//             (dynamic distribution), streamed through a per-warp 3-slot shared-memory ring
//             (previous | current | next) that is filled by 1-D TMA bulk copies
//             (cp.async.bulk + mbarrier), 2 KB per chunk.
//   lane  <-> P consecutive permutations (a_k, b_k live in registers); a 16-token block is
//             read back from shared memory with four broadcast LDS.128.
//
// Arithmetic (bit-exact with numpy's uint64 wrap-then-mod):
//   x = (a*h + b) mod 2^64;  r = (x mod (2^61-1)) & 0xFFFFFFFF
//   With p = 2^61-1:  x mod p = (x & p) + (x >> 61), minus p iff that sum >= p.  The
//   subtract fires for only 36 of the 2^64 values of x; dsk_perm_create proves on the host
//   (modular inverses) that no 32-bit token can reach them for a given permutation set, and
//   then   r = lo32(x) + (hi32(x) >> 29)   exactly  (IMAD.WIDE + IMAD + LEA.HI).
//
// TWO_PHASE kernel (default): per evaluation only  L' = lo32(a_lo*h + b_lo + 7)  is computed
//   (one IMAD) and min-reduced with VIMNMX3; since r = L' - (7 - top3(x)) lies in [L'-7, L'],
//   the true minimum can only come from 16-token blocks whose block-min L' is within 7 of the
//   document min.  Each lane tracks (min, argmin block, second-smallest block min) and then
//   re-evaluates just the winning block exactly; if another block is within the window (or
//   the min is < 7, where L'-7 would wrap) a slow exact pass over candidate blocks runs.
// DIRECT kernel: the full 3.5-instruction evaluation for every (token, perm).
// EXACT kernel: true 64-bit `% p` for every evaluation (any permutation, u32 or u64 tokens).
#include <stdlib.h>

#include "dsk_common.cuh"

namespace dsk {

constexpr int kWarps = 4;           // warps per CTA
constexpr int kChunkBytes = 2048;   // bytes per TMA bulk copy / ring slot
constexpr int kNBuf = 3;            // ring depth per warp: previous | current | next chunk
constexpr int kBlkTok = 16;         // tokens per block (phase-1 tracking granularity)
constexpr uint64_t kP61 = (1ull << 61) - 1;

// r = lo32(x) + top3(x), valid when the `% p` subtract cannot fire (see header).
__device__ __forceinline__ uint32_t eval_fast(uint32_t alo, uint32_t ahi, uint64_t b, uint32_t h) {
    uint64_t x = (uint64_t)alo * h + b;                     // IMAD.WIDE.U32
    uint32_t xh = (uint32_t)(x >> 32) + ahi * h;            // IMAD
    return (uint32_t)x + (xh >> 29);                        // LEA.HI
}

// true ((a*h+b) mod 2^64) % (2^61-1), low 32 bits
__device__ __forceinline__ uint32_t eval_exact(uint64_t a, uint64_t b, uint64_t h) {
    uint64_t x = a * h + b;
    uint64_t s = (x & kP61) + (x >> 61);
    if (s >= kP61) s -= kP61;
    return (uint32_t)s;
}

template <typename TokT> struct TokLoad;
template <> struct TokLoad<uint32_t> {
    // 16 tokens = 4 x LDS.128 (all lanes read the same address: broadcast, 1 wavefront each)
    static __device__ __forceinline__ void block(const uint32_t *src, uint32_t (&t)[kBlkTok]) {
        const uint4 *q = reinterpret_cast<const uint4 *>(src);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            uint4 v = q[i];
            t[4 * i] = v.x; t[4 * i + 1] = v.y; t[4 * i + 2] = v.z; t[4 * i + 3] = v.w;
        }
    }
};
template <> struct TokLoad<uint64_t> {
    static __device__ __forceinline__ void block(const uint64_t *src, uint64_t (&t)[kBlkTok]) {
        const ulonglong2 *q = reinterpret_cast<const ulonglong2 *>(src);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            ulonglong2 v = q[i];
            t[2 * i] = v.x; t[2 * i + 1] = v.y;
        }
    }
};

template <int P, int MODE, typename TokT, int OCC>
__global__ void __launch_bounds__(kWarps * 32, OCC) minhash_bulk_kernel(const BulkParams prm) {
    static_assert(MODE == MODE_EXACT || sizeof(TokT) == 4, "fast paths need 32-bit token hashes");
    constexpr int kBlkBytes = kBlkTok * (int)sizeof(TokT);
    constexpr int kBlkPerChunk = kChunkBytes / kBlkBytes;
    constexpr int kLogBlkPerChunk = (kBlkPerChunk == 32) ? 5 : (kBlkPerChunk == 16) ? 4 : -1;
    static_assert(kLogBlkPerChunk > 0 && (1 << kLogBlkPerChunk) == kBlkPerChunk, "chunk/block geometry");
    static_assert(kNBuf == 3, "ring = previous | current | next chunk");

    __shared__ __align__(128) unsigned char s_buf[kWarps][kNBuf][kChunkBytes];
    __shared__ __align__(16) TokT s_scratch[kWarps][kBlkTok];
    __shared__ __align__(8) uint64_t s_bar[kWarps][kNBuf];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const TokT *__restrict__ tokens = static_cast<const TokT *>(prm.tokens);
    const int64_t *__restrict__ offsets = prm.offsets;
    const int K = prm.k;
    const int kl = blockIdx.y * (32 * P) + lane * P;  // first permutation owned by this lane

    const int64_t n_docs = prm.n_docs, n_tokens = prm.n_tokens;

    // ---- permutation parameters -> registers ------------------------------------------
    uint32_t alo[P], ahi[P], blo[P], bhi[P];
#pragma unroll
    for (int j = 0; j < P; ++j) {
        alo[j] = __ldg(prm.a_lo + kl + j); ahi[j] = __ldg(prm.a_hi + kl + j);
        blo[j] = __ldg(prm.b_lo + kl + j); bhi[j] = __ldg(prm.b_hi + kl + j);
    }

    // ---- per-warp TMA ring: previous | current | next chunk ---------------------------------
    const int64_t copy_end_bytes = (n_tokens * (int64_t)sizeof(TokT)) & ~(int64_t)15;  // bulk copies need 16 B granules
    const int64_t tail_tok = copy_end_bytes / (int64_t)sizeof(TokT);                  // tokens >= this come by plain loads

    uint64_t *bar = s_bar[warp];
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < kNBuf; ++i) mbar_init(&bar[i], 1);
        fence_mbar_init();
    }
    __syncwarp();
    uint32_t par_mask = 0;  // bit s = phase parity the next wait on slot s must use

    // ---- dynamic work distribution: warps pull units of `docs_per_unit` consecutive documents from a
    // global counter, so a warp the SM's arbiter favours simply takes more units and all warps finish
    // together (a static split left ~20 % of the warp slots empty in the tail).
    while (true) {
    int64_t unit = 0;
    if (lane == 0) unit = (int64_t)atomicAdd(prm.work_counter + blockIdx.y, 1u);
    unit = __shfl_sync(0xFFFFFFFFu, unit, 0);
    const int64_t dlo = unit * prm.docs_per_unit;
    if (dlo >= n_docs) break;
    const int64_t dhi = min(dlo + (int64_t)prm.docs_per_unit, n_docs);

    const int64_t tok_lo = __ldg(offsets + dlo), tok_hi = __ldg(offsets + dhi);
    const int64_t blk_begin = tok_lo / kBlkTok;
    const int64_t blk_end = (tok_hi + kBlkTok - 1) / kBlkTok;
    const int64_t nchunks = (tok_hi > tok_lo) ? (blk_end - blk_begin + kBlkPerChunk - 1) / kBlkPerChunk : 0;

    auto issue = [&](int64_t c) {  // lane 0 only; chunk c of this unit lives in slot c % 3
        const int slot = (int)(c % kNBuf);
        const int64_t b0 = blk_begin + c * kBlkPerChunk;
        const int64_t b1 = min(b0 + (int64_t)kBlkPerChunk, blk_end);
        const int64_t byte_lo = b0 * kBlkBytes;
        const int64_t byte_hi = min(b1 * kBlkBytes, copy_end_bytes);
        if (byte_hi > byte_lo) {
            const uint32_t bytes = (uint32_t)(byte_hi - byte_lo);
            mbar_arrive_expect_tx(&bar[slot], bytes);
            bulk_g2s(s_buf[warp][slot], reinterpret_cast<const unsigned char *>(tokens) + byte_lo, bytes, &bar[slot]);
        } else {
            mbar_arrive(&bar[slot]);
        }
    };
    __syncwarp();  // every lane is done with the previous unit's ring contents
    if (lane == 0) {
        for (int64_t c = 0; c < min(nchunks, (int64_t)2); ++c) issue(c);
    }

    int64_t cur = 0;           // chunk currently mapped
    int cur_slot = 0;          // == cur % 3
    int prev_slot = kNBuf - 1; // == (cur - 1) % 3 (only meaningful when cur >= 1)
    bool ready = false;        // cur has been waited on
    auto map_chunk = [&](int64_t c) {
        while (cur < c) {
            __syncwarp();  // every lane is done reading chunk cur-1, whose slot chunk cur+2 takes
            if (lane == 0 && cur + 2 < nchunks) issue(cur + 2);
            ++cur;
            prev_slot = cur_slot;
            if (++cur_slot == kNBuf) cur_slot = 0;
            ready = false;
        }
        if (!ready) {
            mbar_wait(&bar[cur_slot], (par_mask >> cur_slot) & 1u);
            par_mask ^= 1u << cur_slot;
            // the last (< 16 B) tokens of the whole array cannot travel by bulk copy
            const int64_t c0 = (blk_begin + cur * kBlkPerChunk) * kBlkTok;
            const int64_t i = tail_tok + lane;
            if (i < n_tokens && i >= c0 && i < c0 + (int64_t)kBlkPerChunk * kBlkTok)
                reinterpret_cast<TokT *>(s_buf[warp][cur_slot])[i - c0] = tokens[i];
            __syncwarp();
            ready = true;
        }
    };

    // ---- documents ---------------------------------------------------------------------
    int64_t start = tok_lo;
    int64_t end_pref = __ldg(offsets + dlo + 1);  // offsets are read one document ahead
    for (int64_t d = dlo; d < dhi; ++d) {
        const int64_t end = end_pref;
        if (d + 1 < dhi) end_pref = __ldg(offsets + d + 2);

        // TWO_PHASE: m = min L', m2 = 2nd-smallest block min, widx = winning block (doc-local).
        // DIRECT / EXACT: m = running signature.
        uint32_t m[P], m2[P], widx[P];
#pragma unroll
        for (int j = 0; j < P; ++j) { m[j] = 0xFFFFFFFFu; m2[j] = 0xFFFFFFFFu; widx[j] = 0; }

        // one 16-token block (the hot code: 64 IMAD + 32 VIMNMX3 + 20 tracking ops for P=4)
        auto compute = [&](const TokT (&t)[kBlkTok], uint32_t lb) {
            if constexpr (MODE == MODE_TWO_PHASE) {
#pragma unroll
                for (int j = 0; j < P; ++j) {
                    const uint32_t c7 = blo[j] + 7u;
                    uint32_t bm = umin3(alo[j] * (uint32_t)t[0] + c7, alo[j] * (uint32_t)t[1] + c7,
                                        alo[j] * (uint32_t)t[2] + c7);
#pragma unroll
                    for (int i = 3; i < kBlkTok - 1; i += 2)
                        bm = umin3(bm, alo[j] * (uint32_t)t[i] + c7, alo[j] * (uint32_t)t[i + 1] + c7);
                    bm = min(bm, alo[j] * (uint32_t)t[kBlkTok - 1] + c7);
                    const uint32_t om = m[j];
                    m2[j] = min(m2[j], max(bm, om));
                    const bool lt = bm < om;
                    m[j] = lt ? bm : om;
                    widx[j] = lt ? lb : widx[j];
                }
            } else if constexpr (MODE == MODE_DIRECT) {
#pragma unroll
                for (int j = 0; j < P; ++j) {
                    const uint64_t b64 = ((uint64_t)bhi[j] << 32) | blo[j];
#pragma unroll
                    for (int i = 0; i < kBlkTok; i += 2)
                        m[j] = umin3(m[j], eval_fast(alo[j], ahi[j], b64, (uint32_t)t[i]),
                                     eval_fast(alo[j], ahi[j], b64, (uint32_t)t[i + 1]));
                }
            } else {
#pragma unroll
                for (int j = 0; j < P; ++j) {
                    const uint64_t a64 = ((uint64_t)ahi[j] << 32) | alo[j];
                    const uint64_t b64 = ((uint64_t)bhi[j] << 32) | blo[j];
#pragma unroll 4
                    for (int i = 0; i < kBlkTok; ++i) m[j] = min(m[j], eval_exact(a64, b64, (uint64_t)t[i]));
                }
            }
        };
        auto process = [&](const TokT *src, uint32_t lb) {
            TokT t[kBlkTok];
            TokLoad<TokT>::block(src, t);
            compute(t, lb);
        };

        const int64_t dblk0 = start / kBlkTok;
        if (end > start) {
            const int64_t blast = (end - 1) / kBlkTok;
            const bool head_cut = (start % kBlkTok) != 0, tail_cut = (end % kBlkTok) != 0;
            int64_t blk = dblk0;
            while (blk <= blast) {
                map_chunk((blk - blk_begin) >> kLogBlkPerChunk);
                const int64_t cblk0 = blk_begin + cur * kBlkPerChunk;
                const TokT *cbuf = reinterpret_cast<const TokT *>(s_buf[warp][cur_slot]);
                // this document's run of blocks inside the mapped chunk, in 32-bit chunk-local terms
                const int nb = (int)(min(blast + 1, cblk0 + kBlkPerChunk) - blk);
                const TokT *p = cbuf + (int)(blk - cblk0) * kBlkTok;
                const uint32_t lb0 = (uint32_t)(blk - dblk0);
                const bool hp = head_cut && blk == dblk0;           // first block starts mid-block
                const bool tp = tail_cut && blk + nb - 1 == blast;  // last block ends mid-block

                // boundary block: out-of-document slots are replaced by a duplicate of an in-document
                // token (min is idempotent) in a per-warp scratch line, then handled like any other
                auto patched = [&](int i) {
                    if (lane < kBlkTok) {
                        int64_t q = (blk + i) * kBlkTok + lane;
                        q = max(start, min(q, end - 1));
                        s_scratch[warp][lane] = cbuf[q - cblk0 * kBlkTok];
                    }
                    __syncwarp();
                    process(s_scratch[warp], lb0 + (uint32_t)i);
                    __syncwarp();
                };
                int i = 0;
                if (hp || (tp && nb == 1)) { patched(0); i = 1; }
                const int nfull = (tp && nb > 1) ? nb - 1 : nb;
                {   // tight loop over whole blocks: pointer bump + 32-bit counter only
                    const TokT *q = p + i * kBlkTok;
                    const TokT *const qe = p + nfull * kBlkTok;
                    uint32_t lb = lb0 + (uint32_t)i;
                    if constexpr (P > 4) {   // 8 permutations per lane: registers are better spent on the evaluations
#pragma unroll 1
                        for (; q < qe; q += kBlkTok, ++lb) process(q, lb);
                    } else if (q < qe) {
                        // software pipeline (two register buffers): the next block's four LDS.128 are in
                        // flight while the current block's 64 IMADs issue, so the loop never waits on smem
                        TokT ta[kBlkTok], tb[kBlkTok];
                        TokLoad<TokT>::block(q, ta);
#pragma unroll 1
                        while (true) {
                            const TokT *q1 = q + kBlkTok;
                            if (q1 < qe) TokLoad<TokT>::block(q1, tb);
                            compute(ta, lb);
                            if (q1 >= qe) break;
                            const TokT *q2 = q1 + kBlkTok;
                            if (q2 < qe) TokLoad<TokT>::block(q2, ta);
                            compute(tb, lb + 1);
                            if (q2 >= qe) break;
                            q = q2;
                            lb += 2;
                        }
                    }
                }
                if (tp && nb > 1) patched(nb - 1);
                blk += nb;
            }
        }

        // ---- finalise the document ---------------------------------------------------------
        uint32_t res[P];
        if constexpr (MODE == MODE_TWO_PHASE) {
            unsigned need_slow = 0;
#pragma unroll
            for (int j = 0; j < P; ++j) res[j] = 0xFFFFFFFFu;
            if (end > start) {
                // phase 2: exact evaluation of each permutation's winning block.  Fast path: every lane's
                // winners are whole blocks still resident in the ring (previous | current chunk) -> LDS.128.
                const uint32_t *src[P];
                bool fast = true;
                unsigned res_mask = 0;
                // winner position relative to the first block of the mapped chunk (32-bit after one 64-bit add)
                const int64_t rel0 = dblk0 - (blk_begin + cur * kBlkPerChunk);
                const uint32_t nblk = (uint32_t)((end - 1) / kBlkTok - dblk0) + 1u;
                const bool head_cut = (start % kBlkTok) != 0, tail_cut = (end % kBlkTok) != 0;
                const uint32_t *ring_cur = reinterpret_cast<const uint32_t *>(s_buf[warp][cur_slot]);
                const uint32_t *ring_prev = reinterpret_cast<const uint32_t *>(s_buf[warp][prev_slot]);
#pragma unroll
                for (int j = 0; j < P; ++j) {
                    const int64_t rb = rel0 + (int64_t)widx[j];
                    const bool whole = (widx[j] != 0u || !head_cut) && (widx[j] + 1u != nblk || !tail_cut);
                    const bool resident = rb >= -(int64_t)kBlkPerChunk && rb < (int64_t)kBlkPerChunk;
                    src[j] = (rb >= 0 ? ring_cur : ring_prev) + (((int)rb) & (kBlkPerChunk - 1)) * kBlkTok;
                    fast = fast && whole && resident;
                    if (resident) res_mask |= 1u << j;
                    // another block within the +7 window, or L'-7 could wrap: resolve exactly below
                    if (m[j] < 7u || (m2[j] - m[j]) <= 7u) need_slow |= 1u << j;
                }
                if (__all_sync(0xFFFFFFFFu, fast)) {
                    // Refine inside the winning block with the cheap L' first: split its 16 tokens into four
                    // groups of 4; only a group whose min L' is within the +7 window can hold the minimum.
                    // Normally exactly one group qualifies and only its 4 tokens need the full evaluation
                    // (IMAD.WIDE is the expensive instruction); otherwise the slow exact path decides.
                    uint4 v[P][4];
#pragma unroll
                    for (int j = 0; j < P; ++j)
#pragma unroll
                        for (int i = 0; i < 4; ++i) v[j][i] = reinterpret_cast<const uint4 *>(src[j])[i];
                    int pick[P];
#pragma unroll
                    for (int j = 0; j < P; ++j) {
                        const uint32_t c7 = blo[j] + 7u, thr = m[j] + 7u;  // wrap of m+7 implies need_slow (m2-m<=7)
                        bool in[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const uint32_t g = min(umin3(alo[j] * v[j][i].x + c7, alo[j] * v[j][i].y + c7,
                                                         alo[j] * v[j][i].z + c7), alo[j] * v[j][i].w + c7);
                            in[i] = g <= thr;
                        }
                        pick[j] = in[0] ? 0 : in[1] ? 1 : in[2] ? 2 : 3;
                        const int nin = (int)in[0] + (int)in[1] + (int)in[2] + (int)in[3];
                        if (nin != 1) need_slow |= 1u << j;
                    }
                    uint4 w[P];
#pragma unroll
                    for (int j = 0; j < P; ++j) w[j] = reinterpret_cast<const uint4 *>(src[j])[pick[j]];
#pragma unroll
                    for (int j = 0; j < P; ++j) {
                        const uint64_t b64 = ((uint64_t)bhi[j] << 32) | blo[j];
                        res[j] = min(umin3(eval_fast(alo[j], ahi[j], b64, w[j].x), eval_fast(alo[j], ahi[j], b64, w[j].y),
                                           eval_fast(alo[j], ahi[j], b64, w[j].z)), eval_fast(alo[j], ahi[j], b64, w[j].w));
                    }
                } else {
                    // some winner is a boundary block or has left the ring: clamped per-token reads
                    // (duplicates of in-document tokens are harmless), from the ring when resident
#pragma unroll
                    for (int j = 0; j < P; ++j) {
                        const uint64_t b64 = ((uint64_t)bhi[j] << 32) | blo[j];
                        const int64_t base = (dblk0 + (int64_t)widx[j]) * kBlkTok;
                        const bool in_ring = (res_mask >> j) & 1u;
                        uint32_t r = 0xFFFFFFFFu;
#pragma unroll 4
                        for (int i = 0; i < kBlkTok; ++i) {
                            const int64_t q = max(start, min(base + i, end - 1));
                            const uint32_t h = in_ring ? src[j][(int)(q - base)] : (uint32_t)__ldg(tokens + q);
                            r = min(r, eval_fast(alo[j], ahi[j], b64, h));
                        }
                        res[j] = r;
                    }
                }
            }
            if (__any_sync(0xFFFFFFFFu, need_slow != 0)) {
                uint32_t rs[P];
#pragma unroll
                for (int j = 0; j < P; ++j) rs[j] = 0xFFFFFFFFu;
                for (int64_t blk = start / kBlkTok; blk <= (end - 1) / kBlkTok; ++blk) {
                    uint32_t t[kBlkTok];
#pragma unroll
                    for (int i = 0; i < kBlkTok; ++i) {
                        const int64_t q = max(start, min(blk * kBlkTok + i, end - 1));
                        t[i] = (uint32_t)__ldg(tokens + q);
                    }
#pragma unroll
                    for (int j = 0; j < P; ++j) {
                        if (!((need_slow >> j) & 1u)) continue;
                        const uint32_t c7 = blo[j] + 7u;
                        uint32_t bm = 0xFFFFFFFFu;
#pragma unroll
                        for (int i = 0; i < kBlkTok; ++i) bm = min(bm, alo[j] * t[i] + c7);
                        if (m[j] < 7u || (uint64_t)bm <= (uint64_t)m[j] + 7u) {
                            const uint64_t b64 = ((uint64_t)bhi[j] << 32) | blo[j];
#pragma unroll 4
                            for (int i = 0; i < kBlkTok; ++i) rs[j] = min(rs[j], eval_fast(alo[j], ahi[j], b64, t[i]));
                        }
                    }
                }
#pragma unroll
                for (int j = 0; j < P; ++j)
                    if ((need_slow >> j) & 1u) res[j] = rs[j];
            }
        } else {
#pragma unroll
            for (int j = 0; j < P; ++j) res[j] = m[j];
        }

        // ---- merge with the running state (minhash.py:297) and store ---------------------------
        if (prm.init != nullptr) {
#pragma unroll
            for (int j = 0; j < P; ++j) {
                if (kl + j < K) {
                    const int64_t e = d * prm.init_stride + kl + j;
                    if (prm.init_is_u64) {
                        const uint64_t v = __ldg(static_cast<const uint64_t *>(prm.init) + e);
                        res[j] = (uint32_t)min((uint64_t)res[j], v);  // res <= 2^32-1 so the min fits
                    } else {
                        res[j] = min(res[j], __ldg(static_cast<const uint32_t *>(prm.init) + e));
                    }
                }
            }
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
                    if (kl + j < K) row[j] = res[j];
            } else {
                uint32_t *row = static_cast<uint32_t *>(obase) + orow * (int64_t)K + kl;
                if (P == 4 && (K & 3) == 0) {
                    if (kl < K) *reinterpret_cast<uint4 *>(row) = make_uint4(res[0], res[1 % P], res[2 % P], res[3 % P]);
                } else if (P == 8 && (K & 7) == 0) {
                    if (kl < K) {
                        reinterpret_cast<uint4 *>(row)[0] = make_uint4(res[0], res[1 % P], res[2 % P], res[3 % P]);
                        reinterpret_cast<uint4 *>(row)[1] = make_uint4(res[4 % P], res[5 % P], res[6 % P], res[7 % P]);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < P; ++j)
                        if (kl + j < K) row[j] = res[j];
                }
            }
        }
        start = end;
    }
    }  // units
}

// ---- element-wise min of signature matrices (MinHash.merge / union) --------------------
__global__ void sig_merge_min_kernel(const uint32_t *__restrict__ x, const uint32_t *__restrict__ y,
                                     int64_t n, uint32_t *__restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = min(x[i], y[i]);
}

// ---- segmented min: partial signatures of one split document -> its signature --------------------
// A document longer than the split threshold is cut into pieces by the host pipeline so that many
// warps work on it; signatures of pieces combine by element-wise min (minhash.py:337-359).
__global__ void __launch_bounds__(256) seg_min_kernel(const uint32_t *__restrict__ part, const int64_t *__restrict__ seg,
                                                      int64_t n_docs, int k, const void *__restrict__ init,
                                                      int64_t init_stride, int init_is_u64, void *__restrict__ out,
                                                      int out_is_u64) {
    const int64_t total = n_docs * k, stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        const int64_t d = e / k;
        const int c = (int)(e - d * k);
        uint32_t v = 0xFFFFFFFFu;
        for (int64_t r = seg[d]; r < seg[d + 1]; ++r) v = min(v, part[r * k + c]);
        if (init != nullptr) {
            const int64_t ie = d * init_stride + c;
            if (init_is_u64) v = (uint32_t)min((uint64_t)v, static_cast<const uint64_t *>(init)[ie]);
            else v = min(v, static_cast<const uint32_t *>(init)[ie]);
        }
        if (out_is_u64) static_cast<uint64_t *>(out)[e] = v;
        else static_cast<uint32_t *>(out)[e] = v;
    }
}

cudaError_t launch_seg_min(const uint32_t *part, const int64_t *seg, int64_t n_docs, int k, const void *init,
                           int64_t init_stride, int init_is_u64, void *out, int out_is_u64, int sm_count,
                           cudaStream_t s) {
    if (n_docs <= 0) return cudaSuccess;
    int64_t grid = (n_docs * k + 255) / 256;
    if (grid > (int64_t)sm_count * 8) grid = (int64_t)sm_count * 8;
    DSK_LAUNCH(seg_min_kernel, (unsigned)grid, 256, 0, s, part, seg, n_docs, k, init, init_stride, init_is_u64, out, out_is_u64);
    return cudaGetLastError();
}

// ---- launchers --------------------------------------------------------------------------------
template <int P, int MODE, typename TokT, int OCC>
static cudaError_t launch_bulk(const BulkParams &prm_in, int sm_count, cudaStream_t s) {
    BulkParams prm = prm_in;
    const int slices = (prm.k + 32 * P - 1) / (32 * P);
    int64_t gx = (prm.n_docs + kWarps - 1) / kWarps;
    int64_t gmax = (int64_t)sm_count * OCC / slices;  // persistent: every CTA of every slice resident, one wave
    if (gmax < 1) gmax = 1;
    if (gx > gmax) gx = gmax;
    if (gx < 1) gx = 1;
    // unit size: ~8 units per warp for balance, at most 32 documents so a unit's ring restart is amortised
    if (prm.docs_per_unit <= 0) {   // 0 = choose here; > 0 = forced by the caller (tests)
        int64_t dpu = prm.n_docs / (gx * kWarps * 8);
        prm.docs_per_unit = (int)(dpu < 1 ? 1 : (dpu > 32 ? 32 : dpu));
    }
    cudaError_t e = cudaMemsetAsync(prm.work_counter, 0, sizeof(unsigned) * (size_t)slices, s);
    if (e != cudaSuccess) return e;
    dim3 grid((unsigned)gx, (unsigned)slices);
    DSK_LAUNCH((minhash_bulk_kernel<P, MODE, TokT, OCC>), grid, kWarps * 32, 0, s, prm);
    return cudaGetLastError();
}

// The round-1 two-phase kernel (MODE_TWO_PHASE of minhash_bulk_kernel) stays selectable for A/B measurements with
// DSK_TWO_PHASE_V1=1; the default two-phase kernel is minhash_sig_kernel (signature_kernel.cu).
static bool two_phase_v1() {
    static const bool on = [] { const char *e = getenv("DSK_TWO_PHASE_V1"); return e && atoi(e) == 1; }();
    return on;
}

template <int MODE, typename TokT>
static cudaError_t launch_bulk_p(const BulkParams &prm, int sm_count, cudaStream_t s) {
    if (prm.k <= 32) return launch_bulk<1, MODE, TokT, 4>(prm, sm_count, s);
    if (prm.k <= 64) return launch_bulk<2, MODE, TokT, 4>(prm, sm_count, s);
    if (prm.k <= 128) return launch_bulk<4, MODE, TokT, 4>(prm, sm_count, s);
    if (MODE == MODE_TWO_PHASE) return launch_bulk<8, MODE, TokT, 4>(prm, sm_count, s);
    return launch_bulk<8, MODE, TokT, 3>(prm, sm_count, s);
}

cudaError_t launch_minhash_bulk(const BulkParams &prm, int mode, int token_is_u64, int sm_count, cudaStream_t s) {
    if (mode == MODE_TWO_PHASE) {   // signature_kernel.cu: prm.gen names the variant (u32 safe / u32 any permutation / u64 tokens)
        if (prm.gen == 0 && !token_is_u64 && two_phase_v1()) return launch_bulk_p<MODE_TWO_PHASE, uint32_t>(prm, sm_count, s);
        if ((prm.gen == 2) != (token_is_u64 != 0)) return cudaErrorInvalidValue;
        return launch_minhash_sig(prm, sm_count, s);
    }
    if (token_is_u64) return launch_bulk_p<MODE_EXACT, uint64_t>(prm, sm_count, s);
    if (mode == MODE_DIRECT) return launch_bulk_p<MODE_DIRECT, uint32_t>(prm, sm_count, s);
    return launch_bulk_p<MODE_EXACT, uint32_t>(prm, sm_count, s);
}

cudaError_t launch_sig_merge_min(const uint32_t *x, const uint32_t *y, int64_t n, uint32_t *out, int sm_count,
                                 cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    int64_t blocks = (n + 255) / 256;
    if (blocks > (int64_t)sm_count * 8) blocks = (int64_t)sm_count * 8;
    DSK_LAUNCH(sig_merge_min_kernel, (unsigned)blocks, 256, 0, s, x, y, n, out);
    return cudaGetLastError();
}


}  // namespace dsk
