// lsh_kernels.cu -- device-resident MinHashLSH index (sm_100a).
//
// Replaces the per-document Python work of MinHashLSH._insert / query
// (datasketch/lsh.py:326-347, :370-432 over dict storage, storage.py:209-259):
//   insert: for each band j, bucket[j][ _H(sig[j*r:(j+1)*r]) ].add(doc)
//   query : union over bands of bucket[j][ _H(qsig[j*r:(j+1)*r]) ]
// The reference's bucket key is the r-tuple itself (`_H` = byteswap, lsh.py:537-538), so two
// documents share a bucket iff their r-tuples are EQUAL.  Here a bucket is found through a 64-bit
// fingerprint of the tuple in a per-band open-addressing table (atomicCAS claim), members are
// chained through next[band][doc] (atomicExch on the bucket head), and every candidate is verified
// by comparing the r-tuple itself, so the candidate sets are exactly the reference's.
// A candidate that matches in several bands is emitted once (by the first matching band), which is
// the set-union of lsh.py:426-429.
#include "dsk_common.cuh"

namespace dsk {

// kEmptyKey, lsh_mix64, band_fp: dsk_common.cuh (shared with the signature kernel's fused insert epilogue)

__device__ __forceinline__ bool tuple_eq(const uint32_t *a, const uint32_t *b, int r) {
    bool eq = true;
    for (int q = 0; q < r; ++q) eq = eq && (a[q] == b[q]);
    return eq;
}

// CTA <-> tile of D = 256 / b new documents.  The tile's rows are contiguous: they are read ONCE with fully coalesced
// loads, each value going to the index's own copy of the signatures (candidates are verified on the r-tuples
// themselves) and to shared memory; then thread <-> (document, band) takes its r-tuple from shared memory for the
// fingerprint and does the table update -- every thread has one atomicCAS chain in flight, which is what hides the DRAM
// latency of the (much larger than L2) tables.  next[] is laid out [doc][band], so the tile's links are one coalesced
// store.  The separate device-to-device copy and the b uncoalesced re-reads of each row of the first version are gone.
__global__ void __launch_bounds__(256) lsh_insert_kernel(const LshDev ix, const uint32_t *__restrict__ new_sig,
                                                         int64_t doc0, int64_t n_new, int docs_per_tile) {
    DSK_DYNAMIC_SMEM_T(uint32_t, s_rows, 16);
    const uint64_t mask = (uint64_t)ix.cap_slots - 1;
    const int64_t n_tiles = (n_new + docs_per_tile - 1) / docs_per_tile;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t i0 = tile * docs_per_tile;
        const int nd = (int)min((int64_t)docs_per_tile, n_new - i0);
        const uint32_t *src = new_sig + i0 * ix.k;
        uint32_t *keep = ix.sig + (doc0 + i0) * ix.k;
        __syncthreads();
        for (int e = threadIdx.x; e < nd * ix.k; e += blockDim.x) {
            const uint32_t v = __ldg(src + e);
            s_rows[e] = v;
            if (keep != src) keep[e] = v;   // (dsk_lsh_insert_tokens' two-kernel route built the rows in place)
        }
        __syncthreads();
        for (int pair = threadIdx.x; pair < nd * ix.b; pair += blockDim.x) {
            const int i = pair / ix.b, band = pair - i * ix.b;
            const int64_t doc = doc0 + i0 + i;
            const uint64_t fp = band_fp(s_rows + i * ix.k + band * ix.r, ix.r, band);
            uint64_t *tab = ix.slots + (int64_t)band * ix.cap_slots * 2;
            uint64_t slot = lsh_mix64(fp) & mask;
            while (true) {
                const unsigned long long prev = atomicCAS(reinterpret_cast<unsigned long long *>(tab + 2 * slot),
                                                          (unsigned long long)kEmptyKey, (unsigned long long)fp);
                if (prev == kEmptyKey || prev == fp) break;
                slot = (slot + 1) & mask;
            }
            const int32_t old = atomicExch(reinterpret_cast<int32_t *>(tab + 2 * slot + 1), (int32_t)doc);
            ix.next[doc * ix.b + band] = old;
        }
    }
}

// warp <-> query, lane <-> band (bands beyond 32 are handled in further rounds).
// FILL = false: counts[q] = number of distinct candidates.  FILL = true: writes them at out[ptr[q]...].
template <bool FILL>
__global__ void __launch_bounds__(256) lsh_query_kernel(const LshDev ix, const uint32_t *__restrict__ qsig, int64_t nq,
                                                        int64_t n_docs, int64_t *__restrict__ counts,
                                                        const int64_t *__restrict__ ptr, int32_t *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    const uint64_t mask = (uint64_t)ix.cap_slots - 1;
    for (int64_t q = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); q < nq; q += warps) {
        const uint32_t *qrow = qsig + q * ix.k;
        int64_t base = FILL ? ptr[q] : 0;
        int64_t total = 0;
        for (int band0 = 0; band0 < ix.b; band0 += 32) {
            const int band = band0 + lane;
            // pass 0 counts this lane's distinct candidates, pass 1 (FILL) writes them after a warp scan
            int mine = 0;
            int64_t wpos = 0;
            for (int pass = 0; pass < (FILL ? 2 : 1); ++pass) {
                int cnt = 0;
                if (band < ix.b) {
                    const uint32_t *qt = qrow + (int64_t)band * ix.r;
                    const uint64_t fp = band_fp(qt, ix.r, band);
                    const uint64_t *tab = ix.slots + (int64_t)band * ix.cap_slots * 2;
                    uint64_t slot = lsh_mix64(fp) & mask;
                    int32_t d = -1;
                    while (true) {
                        const uint64_t kk = tab[2 * slot];
                        if (kk == fp) { d = *reinterpret_cast<const int32_t *>(tab + 2 * slot + 1); break; }
                        if (kk == kEmptyKey) break;
                        slot = (slot + 1) & mask;
                    }
                    while (d >= 0) {
                        if (d < n_docs) {
                            const uint32_t *drow = ix.sig + (int64_t)d * ix.k;
                            if (tuple_eq(drow + (int64_t)band * ix.r, qt, ix.r)) {
                                bool dup = false;  // already produced by an earlier band?
                                for (int j = 0; j < band && !dup; ++j)
                                    dup = tuple_eq(drow + (int64_t)j * ix.r, qrow + (int64_t)j * ix.r, ix.r);
                                if (!dup) {
                                    if (FILL && pass == 1) out[wpos + cnt] = d;
                                    ++cnt;
                                }
                            }
                        }
                        d = ix.next[(int64_t)d * ix.b + band];
                    }
                }
                if (pass == 0) {
                    mine = cnt;
                    int incl = mine;  // inclusive warp scan of per-lane counts
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const int t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                        if (lane >= o) incl += t;
                    }
                    wpos = base + total + (incl - mine);
                    total += __shfl_sync(0xFFFFFFFFu, incl, 31);
                }
            }
        }
        if (!FILL && lane == 0) counts[q] = total;
    }
}

// ---- exclusive scan of int64 counts (three small kernels) -------------------------------------------
constexpr int kScanBlock = 1024;

__global__ void __launch_bounds__(kScanBlock) scan_block_kernel(const int64_t *in, int64_t n, int64_t *out,
                                                                int64_t *block_sums) {
    __shared__ int64_t s[kScanBlock];
    const int64_t i = (int64_t)blockIdx.x * kScanBlock + threadIdx.x;
    const int64_t v = i < n ? in[i] : 0;
    s[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < kScanBlock; o <<= 1) {
        const int64_t t = threadIdx.x >= o ? s[threadIdx.x - o] : 0;
        __syncthreads();
        s[threadIdx.x] += t;
        __syncthreads();
    }
    if (i < n) out[i] = s[threadIdx.x] - v;  // exclusive
    if (threadIdx.x == kScanBlock - 1) block_sums[blockIdx.x] = s[threadIdx.x];
}

__global__ void scan_sums_kernel(int64_t *block_sums, int64_t nb) {  // single thread block, serial over blocks
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        int64_t run = 0;
        for (int64_t i = 0; i < nb; ++i) { const int64_t t = block_sums[i]; block_sums[i] = run; run += t; }
        block_sums[nb] = run;
    }
}

__global__ void __launch_bounds__(kScanBlock) scan_add_kernel(int64_t *out, int64_t n, const int64_t *block_sums,
                                                              int64_t nb) {
    const int64_t i = (int64_t)blockIdx.x * kScanBlock + threadIdx.x;
    if (i < n) out[i] += block_sums[blockIdx.x];
    if (i == 0) out[n] = block_sums[nb];  // total
}

// ---- launchers ---------------------------------------------------------------------------------------
cudaError_t launch_lsh_insert(const LshDev &ix, const uint32_t *new_sig, int64_t doc0, int64_t n_new, int sm_count,
                              cudaStream_t s) {
    if (n_new <= 0) return cudaSuccess;
    int docs_per_tile = 256 / ix.b;                       // one (document, band) pair per thread
    if (docs_per_tile < 1) docs_per_tile = 1;
    while (docs_per_tile > 1 && (size_t)docs_per_tile * ix.k * 4 > 48 * 1024) --docs_per_tile;
    const size_t smem = (size_t)docs_per_tile * ix.k * 4;
    cudaError_t e = cudaFuncSetAttribute(lsh_insert_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int64_t grid = (n_new + docs_per_tile - 1) / docs_per_tile;
    if (grid > (int64_t)sm_count * 8) grid = (int64_t)sm_count * 8;
    DSK_LAUNCH(lsh_insert_kernel, (unsigned)grid, 256, smem, s, ix, new_sig, doc0, n_new, docs_per_tile);
    return cudaGetLastError();
}

cudaError_t launch_lsh_query(const LshDev &ix, const uint32_t *qsig, int64_t nq, int64_t n_docs, int64_t *counts,
                             const int64_t *ptr, int32_t *out, int fill, int sm_count, cudaStream_t s) {
    if (nq <= 0) return cudaSuccess;
    int64_t grid = (nq + 7) / 8;
    if (grid > (int64_t)sm_count * 8) grid = (int64_t)sm_count * 8;
    if (fill) DSK_LAUNCH((lsh_query_kernel<true>), (unsigned)grid, 256, 0, s, ix, qsig, nq, n_docs, counts, ptr, out);
    else DSK_LAUNCH((lsh_query_kernel<false>), (unsigned)grid, 256, 0, s, ix, qsig, nq, n_docs, counts, ptr, out);
    return cudaGetLastError();
}

// out has n + 1 entries; scratch needs (n / 1024 + 2) int64
cudaError_t launch_exclusive_scan(const int64_t *in, int64_t n, int64_t *out, int64_t *scratch, cudaStream_t s) {
    const int64_t nb = (n + kScanBlock - 1) / kScanBlock;
    if (n <= 0) return cudaMemsetAsync(out, 0, sizeof(int64_t), s);
    DSK_LAUNCH(scan_block_kernel, (unsigned)nb, kScanBlock, 0, s, in, n, out, scratch);
    DSK_LAUNCH(scan_sums_kernel, 1, 32, 0, s, scratch, nb);
    DSK_LAUNCH(scan_add_kernel, (unsigned)nb, kScanBlock, 0, s, out, n, scratch, nb);
    return cudaGetLastError();
}

}  // namespace dsk

// ---- LSH Forest query (datasketch/lshforest.py:74-128) --------------------------------------------------
// order[tree][pos] = document numbers sorted by (the tree's k-tuple, document number); the sorted list of
// distinct keys + per-key insertion-ordered buckets of the reference flatten to exactly this sequence.
// warp <-> query.  For each prefix length r = k..1: lanes binary-search the l trees in parallel (tuple compare
// on the signatures themselves), then the matches are emitted tree by tree, in sequence order, into the
// query's result list (first `topk` DISTINCT documents), stopping exactly where the reference returns.
namespace dsk {

__device__ __forceinline__ int lex_cmp(const uint32_t *a, const uint32_t *b, int r) {
    for (int q = 0; q < r; ++q) {
        const uint32_t x = a[q], y = b[q];
        if (x != y) return x < y ? -1 : 1;
    }
    return 0;
}

__global__ void __launch_bounds__(128) forest_query_kernel(const uint32_t *__restrict__ sig,
                                                           const int32_t *__restrict__ order, int64_t n, int K, int l,
                                                           int k, const uint32_t *__restrict__ qsig, int64_t nq,
                                                           int topk, int32_t *__restrict__ out) {
    DSK_DYNAMIC_SMEM_T(int32_t, s_res, 4);  // [warps][topk]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int32_t *res = s_res + warp * topk;
    const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t q = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp; q < nq; q += warps) {
        const uint32_t *qrow = qsig + q * K;
        int cnt = 0;
        bool done = false;
        for (int r = k; r >= 1 && !done; --r) {
            for (int t0 = 0; t0 < l && !done; t0 += 32) {
                const int tree = t0 + lane;
                int64_t lo = 0;
                if (tree < l) {  // lower bound of the r-prefix in this lane's tree
                    const uint32_t *qt = qrow + tree * k;
                    const int32_t *ord = order + (int64_t)tree * n;
                    int64_t hi = n;
                    while (lo < hi) {
                        const int64_t mid = lo + ((hi - lo) >> 1);
                        if (lex_cmp(sig + (int64_t)ord[mid] * K + tree * k, qt, r) < 0) lo = mid + 1;
                        else hi = mid;
                    }
                }
                const int ntree = min(32, l - t0);
                for (int j = 0; j < ntree && !done; ++j) {  // emit in tree order; lane 0 owns the result list
                    int64_t pos = __shfl_sync(0xFFFFFFFFu, lo, j);
                    if (lane == 0) {
                        const int tj = t0 + j;
                        const uint32_t *qt = qrow + tj * k;
                        const int32_t *ord = order + (int64_t)tj * n;
                        while (pos < n) {
                            const int32_t d = ord[pos];
                            if (lex_cmp(sig + (int64_t)d * K + tj * k, qt, r) != 0) break;
                            bool seen = false;
                            for (int e = 0; e < cnt; ++e) seen = seen || (res[e] == d);
                            if (!seen) {
                                res[cnt++] = d;
                                if (cnt >= topk) break;
                            }
                            ++pos;
                        }
                    }
                    cnt = __shfl_sync(0xFFFFFFFFu, cnt, 0);
                    done = cnt >= topk;
                }
            }
        }
        __syncwarp();
        for (int e = lane; e < topk; e += 32) out[q * topk + e] = e < cnt ? res[e] : -1;
        __syncwarp();
    }
}

cudaError_t launch_forest_query(const uint32_t *sig, const int32_t *order, int64_t n, int K, int l, int k,
                                const uint32_t *qsig, int64_t nq, int topk, int32_t *out, int sm_count,
                                cudaStream_t s) {
    if (nq <= 0) return cudaSuccess;
    int64_t grid = (nq + 3) / 4;
    if (grid > (int64_t)sm_count * 8) grid = (int64_t)sm_count * 8;
    const size_t smem = (size_t)4 * topk * sizeof(int32_t);
    DSK_LAUNCH(forest_query_kernel, (unsigned)grid, 128, smem, s, sig, order, n, K, l, k, qsig, nq, topk, out);
    return cudaGetLastError();
}

}  // namespace dsk
