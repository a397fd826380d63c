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

cudaError_t launch_jaccard_pairs(const uint32_t *sig, int64_t n_rows, int k, const int64_t *ia, const int64_t *ib,
                                 int64_t m, int32_t *out, int sm_count, cudaStream_t s) {
    if (m <= 0) return cudaSuccess;
    int64_t grid = (m + 7) / 8;
    if (grid > (int64_t)sm_count * 8) grid = (int64_t)sm_count * 8;
    jaccard_pairs_kernel<<<(unsigned)grid, 256, 0, s>>>(sig, n_rows, k, ia, ib, m, out);
    return cudaGetLastError();
}

}  // namespace dsk
