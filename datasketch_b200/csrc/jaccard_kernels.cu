// jaccard_kernels.cu -- Jaccard estimate = count of equal signature positions / K
// (datasketch/minhash.py:324, inherited by LeanMinHash) on sm_100a.
//
//  jaccard_pairs : m explicit (i, j) pairs over a [N, K] u32 signature matrix.  HBM-bound gather:
//                  2 * 4K bytes in, 4 bytes out per pair.  warp <-> pair, lanes read the two rows
//                  with 128-bit loads, compare, popc-free add, warp-shuffle reduce.
//  jaccard_topk  : for each query row the `topk` database rows with the most equal positions
//                  (ties -> lower index, optional self-exclusion): tiled all-pairs compare with
//                  register blocking; see the kernel comment.
#include "dsk_common.cuh"

namespace dsk {

__device__ __forceinline__ int eq4(const uint4 a, const uint4 b) {
    return (a.x == b.x) + (a.y == b.y) + (a.z == b.z) + (a.w == b.w);
}

__global__ void __launch_bounds__(256) jaccard_pairs_kernel(const uint32_t *__restrict__ sig, int64_t n_rows, int k,
                                                            const int64_t *__restrict__ ia,
                                                            const int64_t *__restrict__ ib, int64_t m,
                                                            int32_t *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    const bool vec = (k & 3) == 0;
    for (int64_t p = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); p < m; p += warps) {
        const int64_t i = __ldg(ia + p), j = __ldg(ib + p);
        int cnt = 0;
        if (i >= 0 && i < n_rows && j >= 0 && j < n_rows) {
            const uint32_t *x = sig + i * k, *y = sig + j * k;
            if (vec) {
                const uint4 *x4 = reinterpret_cast<const uint4 *>(x), *y4 = reinterpret_cast<const uint4 *>(y);
                for (int c = lane; c < (k >> 2); c += 32) cnt += eq4(__ldg(x4 + c), __ldg(y4 + c));
            } else {
                for (int c = lane; c < k; c += 32) cnt += (__ldg(x + c) == __ldg(y + c));
            }
        } else {
            cnt = (lane == 0) ? -1 : 0;  // out-of-range pair index -> -1
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, o);
        if (lane == 0) out[p] = cnt;
    }
}

// ---- all-pairs top-k -----------------------------------------------------------------------------
// CTA tile: kTQ = 64 query rows x kTD = 128 database rows per step; 128 threads, thread <-> 8 queries x
// 8 database rows (64 counters in registers).  Both tiles are staged TRANSPOSED in shared memory
// ([position][row]) so a thread reads its 8 + 8 values of one signature position with four LDS.128.
// After a database tile, counts above the query's current k-th best are inserted into the query's
// sorted list in shared memory (ties keep the lower database index because tiles arrive in index
// order and insertion requires a strictly larger count than the k-th entry ... or an unfilled slot).
constexpr int kTQ = 64, kTD = 128, kKc = 32, kTopThreads = 128, kMaxTopK = 32;

struct TopkParams {
    const uint32_t *q;    // [nq, k]
    const uint32_t *db;   // [n, k]
    int64_t nq, n;
    int k, topk;
    int64_t self_base;    // query i is database row self_base + i (excluded), or < 0
    int32_t *out_cnt;     // [nq, topk]
    int64_t *out_idx;     // [nq, topk]  (-1 = no such row)
};

__global__ void __launch_bounds__(kTopThreads) jaccard_topk_kernel(const TopkParams p) {
    DSK_DYNAMIC_SMEM_T(uint32_t, sm, 16);
    const int K = p.k;
    const int kpad = (K + kKc - 1) / kKc * kKc;
    uint32_t *sQ = sm;                              // [kpad][kTQ]
    uint32_t *sD = sQ + (size_t)kpad * kTQ;         // [kKc][kTD]
    int32_t *top_cnt = reinterpret_cast<int32_t *>(sD + kKc * kTD);   // [kTQ][topk]
    int64_t *top_idx = reinterpret_cast<int64_t *>(top_cnt + kTQ * kMaxTopK);  // [kTQ][topk]
    int *lock = reinterpret_cast<int *>(top_idx + kTQ * kMaxTopK);    // [kTQ]
    volatile int32_t *thr = reinterpret_cast<volatile int32_t *>(lock + kTQ);  // [kTQ] k-th best count, -1 while unfilled
    volatile int64_t *kidx = reinterpret_cast<volatile int64_t *>(const_cast<int32_t *>(thr) + kTQ);  // [kTQ] its index

    const int tid = threadIdx.x;
    const int qg = tid >> 4, dg = tid & 15;         // 8 query groups x 16 database groups
    const int topk = p.topk;

    for (int64_t qt = blockIdx.x; qt * kTQ < p.nq; qt += gridDim.x) {
        const int64_t q0 = qt * kTQ;
        __syncthreads();
        for (int i = tid; i < kTQ * kMaxTopK; i += kTopThreads) { top_cnt[i] = -1; top_idx[i] = -1; }
        for (int i = tid; i < kTQ; i += kTopThreads) { lock[i] = 0; thr[i] = -1; kidx[i] = -1; }
        // query tile, transposed; rows / positions beyond the data are filled so they never match
        for (int e = tid; e < kTQ * kpad; e += kTopThreads) {
            const int r = e / kpad, c = e - r * kpad;
            uint32_t v = 0x51EDu + (uint32_t)c;           // padded positions never match: the db side uses ~filler
            if (q0 + r < p.nq && c < K) v = __ldg(p.q + (q0 + r) * (int64_t)K + c);
            sQ[(size_t)c * kTQ + r] = v;
        }
        __syncthreads();

        for (int64_t d0 = 0; d0 < p.n; d0 += kTD) {
            int cnt[8][8];
#pragma unroll
            for (int a = 0; a < 8; ++a)
#pragma unroll
                for (int b = 0; b < 8; ++b) cnt[a][b] = 0;
            for (int c0 = 0; c0 < kpad; c0 += kKc) {
                __syncthreads();
                // stage database chunk [kTD rows][kKc positions] transposed: each task = (row, 8 positions)
                for (int task = tid; task < kTD * (kKc / 8); task += kTopThreads) {
                    const int r = task % kTD, seg = task / kTD;
                    const int64_t row = d0 + r;
                    const int cb = c0 + seg * 8;
                    uint32_t v[8];
                    if (row < p.n && (K & 3) == 0 && cb + 8 <= K) {   // one full 32-byte sector per task
                        const uint4 x = __ldg(reinterpret_cast<const uint4 *>(p.db + row * (int64_t)K + cb));
                        const uint4 y = __ldg(reinterpret_cast<const uint4 *>(p.db + row * (int64_t)K + cb + 4));
                        v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w; v[4] = y.x; v[5] = y.y; v[6] = y.z; v[7] = y.w;
                    } else {
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            v[u] = ~(0x51EDu + (uint32_t)(cb + u));  // never equals the query filler at this position
                            if (row < p.n && cb + u < K) v[u] = __ldg(p.db + row * (int64_t)K + cb + u);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u) sD[(seg * 8 + u) * kTD + r] = v[u];
                }
                __syncthreads();
#pragma unroll 4
                for (int cc = 0; cc < kKc; ++cc) {
                    const uint4 qa = *reinterpret_cast<const uint4 *>(sQ + (size_t)(c0 + cc) * kTQ + qg * 8);
                    const uint4 qb = *reinterpret_cast<const uint4 *>(sQ + (size_t)(c0 + cc) * kTQ + qg * 8 + 4);
                    const uint4 da = *reinterpret_cast<const uint4 *>(sD + cc * kTD + dg * 8);
                    const uint4 db = *reinterpret_cast<const uint4 *>(sD + cc * kTD + dg * 8 + 4);
                    const uint32_t qv[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
                    const uint32_t dv[8] = {da.x, da.y, da.z, da.w, db.x, db.y, db.z, db.w};
#pragma unroll
                    for (int a = 0; a < 8; ++a)
#pragma unroll
                        for (int b = 0; b < 8; ++b) cnt[a][b] += (qv[a] == dv[b]);
                }
            }
            // fold this tile's counts into the per-query sorted lists
#pragma unroll
            for (int a = 0; a < 8; ++a) {
                const int ql = qg * 8 + a;
                const int64_t qrow = q0 + ql;
                if (qrow >= p.nq) continue;
#pragma unroll
                for (int b = 0; b < 8; ++b) {
                    const int64_t drow = d0 + dg * 8 + b;
                    if (drow >= p.n) continue;
                    if (p.self_base >= 0 && drow == p.self_base + qrow) continue;
                    const int c = cnt[a][b];
                    {   // cheap filter (re-checked under the lock): worse than the current k-th entry?
                        const int32_t t = thr[ql];
                        if (c < t || (c == t && drow > kidx[ql])) continue;
                    }
                    bool done = false;
                    while (!done) {
                        if (atomicCAS(&lock[ql], 0, 1) == 0) {
                            int32_t *tc = top_cnt + ql * kMaxTopK;
                            int64_t *ti = top_idx + ql * kMaxTopK;
                            // position: after every entry with count > c, or count == c and lower index
                            int pos = topk;
                            for (int s = 0; s < topk; ++s) {
                                const int32_t sc = tc[s];
                                if (sc < c || (sc == c && (ti[s] < 0 || ti[s] > drow))) { pos = s; break; }
                            }
                            if (pos < topk) {
                                for (int s = topk - 1; s > pos; --s) { tc[s] = tc[s - 1]; ti[s] = ti[s - 1]; }
                                tc[pos] = c; ti[pos] = drow;
                                kidx[ql] = ti[topk - 1];
                                thr[ql] = (ti[topk - 1] < 0) ? -1 : tc[topk - 1];
                            }
                            __threadfence_block();
                            atomicExch(&lock[ql], 0);
                            done = true;
                        }
                    }
                }
            }
        }
        __syncthreads();
        for (int e = tid; e < kTQ * topk; e += kTopThreads) {
            const int r = e / topk, s = e - r * topk;
            if (q0 + r < p.nq) {
                p.out_cnt[(q0 + r) * topk + s] = top_cnt[r * kMaxTopK + s];
                p.out_idx[(q0 + r) * topk + s] = top_idx[r * kMaxTopK + s];
            }
        }
    }
}

size_t jaccard_topk_smem(int k) {
    const int kpad = (k + kKc - 1) / kKc * kKc;
    return (size_t)kpad * kTQ * 4 + (size_t)kKc * kTD * 4 + (size_t)kTQ * kMaxTopK * (4 + 8) + (size_t)kTQ * 16;
}

cudaError_t launch_jaccard_topk(const uint32_t *q, int64_t nq, const uint32_t *db, int64_t n, int k, int topk,
                                int64_t self_base, int32_t *out_cnt, int64_t *out_idx, int sm_count, cudaStream_t s) {
    if (nq <= 0) return cudaSuccess;
    TopkParams p;
    p.q = q; p.db = db; p.nq = nq; p.n = n; p.k = k; p.topk = topk; p.self_base = self_base;
    p.out_cnt = out_cnt; p.out_idx = out_idx;
    const size_t smem = jaccard_topk_smem(k);
    cudaError_t e = cudaFuncSetAttribute(jaccard_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int64_t grid = (nq + kTQ - 1) / kTQ;
    if (grid > (int64_t)sm_count * 3) grid = (int64_t)sm_count * 3;
    DSK_LAUNCH(jaccard_topk_kernel, (unsigned)grid, kTopThreads, smem, s, p);
    return cudaGetLastError();
}

cudaError_t launch_jaccard_pairs(const uint32_t *sig, int64_t n_rows, int k, const int64_t *ia, const int64_t *ib,
                                 int64_t m, int32_t *out, int sm_count, cudaStream_t s) {
    if (m <= 0) return cudaSuccess;
    int64_t grid = (m + 7) / 8;
    if (grid > (int64_t)sm_count * 8) grid = (int64_t)sm_count * 8;
    DSK_LAUNCH(jaccard_pairs_kernel, (unsigned)grid, 256, 0, s, sig, n_rows, k, ia, ib, m, out);
    return cudaGetLastError();
}

}  // namespace dsk
