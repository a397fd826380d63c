// loopbench.cu -- the signature kernel's block loop in isolation: 4 permutations per lane, 16-token blocks from shared
// memory, 64 IMAD + 32 VIMNMX3 + tracking per block.  Variants differ in SOURCE ORDER only; prints IMAD pipe utilisation.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
constexpr int P = 4, NBLK = 32, ITERS = 200;

__device__ __forceinline__ uint32_t umin3(uint32_t a, uint32_t b, uint32_t c) { return min(min(a, b), c); }

template <int V>
__global__ void __launch_bounds__(128, 4) k(const uint32_t *g_a, const uint32_t *g_c, const uint32_t *g_tok, uint32_t *out, unsigned long long *cyc) {
    __shared__ __align__(16) uint32_t s_tok[4][NBLK * 16];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t *tok = s_tok[warp];
    for (int i = lane; i < NBLK * 16; i += 32) tok[i] = g_tok[(blockIdx.x * 4 + warp) * 7 + i];
    uint32_t a[P], c[P], m[P], m2[P];
#pragma unroll
    for (int j = 0; j < P; ++j) { a[j] = g_a[threadIdx.x * P + j]; c[j] = g_c[threadIdx.x * P + j]; m[j] = m2[j] = 0xFFFFFFFFu; }
    __syncthreads();
    unsigned long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll 1
        for (int b = 0; b < NBLK; ++b) {
            const uint4 *q = reinterpret_cast<const uint4 *>(tok + b * 16);
            uint32_t t[16];
#pragma unroll
            for (int i = 0; i < 4; ++i) { uint4 v = q[i]; t[4*i] = v.x; t[4*i+1] = v.y; t[4*i+2] = v.z; t[4*i+3] = v.w; }
            if (V == 0) {          // compiler order, permutation-major source (the kernel's)
#pragma unroll
                for (int j = 0; j < P; ++j) {
                    uint32_t bm = umin3(a[j]*t[0]+c[j], a[j]*t[1]+c[j], a[j]*t[2]+c[j]);
#pragma unroll
                    for (int i = 3; i < 15; i += 2) bm = umin3(bm, a[j]*t[i]+c[j], a[j]*t[i+1]+c[j]);
                    bm = min(bm, a[j]*t[15]+c[j]);
                    const uint32_t key = (bm & ~31u) | (uint32_t)b;
                    const uint32_t om = m[j]; m2[j] = min(m2[j], max(key, om)); m[j] = min(om, key);
                }
            } else if (V == 1) {   // one asm block per permutation: 16 IMAD back to back, then the min tree
#pragma unroll
                for (int j = 0; j < P; ++j) {
                    uint32_t bm;
                    asm volatile("{\n\t.reg .u32 d<16>;\n\t"
                        "mad.lo.u32 d0, %1, %3, %2;\n\tmad.lo.u32 d1, %1, %4, %2;\n\tmad.lo.u32 d2, %1, %5, %2;\n\tmad.lo.u32 d3, %1, %6, %2;\n\t"
                        "mad.lo.u32 d4, %1, %7, %2;\n\tmad.lo.u32 d5, %1, %8, %2;\n\tmad.lo.u32 d6, %1, %9, %2;\n\tmad.lo.u32 d7, %1, %10, %2;\n\t"
                        "mad.lo.u32 d8, %1, %11, %2;\n\tmad.lo.u32 d9, %1, %12, %2;\n\tmad.lo.u32 d10, %1, %13, %2;\n\tmad.lo.u32 d11, %1, %14, %2;\n\t"
                        "mad.lo.u32 d12, %1, %15, %2;\n\tmad.lo.u32 d13, %1, %16, %2;\n\tmad.lo.u32 d14, %1, %17, %2;\n\tmad.lo.u32 d15, %1, %18, %2;\n\t"
                        "min.u32 d0, d0, d1;\n\tmin.u32 d0, d0, d2;\n\tmin.u32 d3, d3, d4;\n\tmin.u32 d3, d3, d5;\n\t"
                        "min.u32 d6, d6, d7;\n\tmin.u32 d6, d6, d8;\n\tmin.u32 d9, d9, d10;\n\tmin.u32 d9, d9, d11;\n\t"
                        "min.u32 d12, d12, d13;\n\tmin.u32 d12, d12, d14;\n\tmin.u32 d0, d0, d3;\n\tmin.u32 d0, d0, d6;\n\t"
                        "min.u32 d9, d9, d12;\n\tmin.u32 d9, d9, d15;\n\tmin.u32 %0, d0, d9;\n\t}"
                        : "=r"(bm) : "r"(a[j]), "r"(c[j]), "r"(t[0]), "r"(t[1]), "r"(t[2]), "r"(t[3]), "r"(t[4]), "r"(t[5]), "r"(t[6]), "r"(t[7]),
                          "r"(t[8]), "r"(t[9]), "r"(t[10]), "r"(t[11]), "r"(t[12]), "r"(t[13]), "r"(t[14]), "r"(t[15]));
                    const uint32_t key = (bm & ~31u) | (uint32_t)b;
                    const uint32_t om = m[j]; m2[j] = min(m2[j], max(key, om)); m[j] = min(om, key);
                }
            } else if (V == 2) {   // token-major: two tokens at a time through all permutations
                uint32_t bm[P];
#pragma unroll
                for (int j = 0; j < P; ++j) bm[j] = 0xFFFFFFFFu;
#pragma unroll
                for (int i = 0; i < 16; i += 2) {
#pragma unroll
                    for (int j = 0; j < P; ++j) {
                        uint32_t x, y;
                        asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(x) : "r"(a[j]), "r"(t[i]), "r"(c[j]));
                        asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(y) : "r"(a[j]), "r"(t[i+1]), "r"(c[j]));
                        asm volatile("{.reg .u32 z; min.u32 z, %0, %1; min.u32 %0, z, %2;}" : "+r"(bm[j]) : "r"(x), "r"(y));
                    }
                }
#pragma unroll
                for (int j = 0; j < P; ++j) {
                    const uint32_t key = (bm[j] & ~31u) | (uint32_t)b;
                    const uint32_t om = m[j]; m2[j] = min(m2[j], max(key, om)); m[j] = min(om, key);
                }
            } else if (V == 3) {   // IMADs only (no mins beyond one per permutation): the multiplier's own ceiling in this loop shape
#pragma unroll
                for (int j = 0; j < P; ++j) {
                    uint32_t s = 0;
#pragma unroll
                    for (int i = 0; i < 16; ++i) s ^= a[j]*t[i]+c[j];
                    m[j] = min(m[j], s);
                }
            } else if (V == 4) {   // permutation-major volatile triples: IMAD IMAD MIN3, same a and c throughout a permutation
#pragma unroll
                for (int j = 0; j < P; ++j) {
                    uint32_t bm = 0xFFFFFFFFu;
#pragma unroll
                    for (int i = 0; i < 16; i += 2) {
                        uint32_t x, y;
                        asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(x) : "r"(a[j]), "r"(t[i]), "r"(c[j]));
                        asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(y) : "r"(a[j]), "r"(t[i+1]), "r"(c[j]));
                        asm volatile("{.reg .u32 z; min.u32 z, %0, %1; min.u32 %0, z, %2;}" : "+r"(bm) : "r"(x), "r"(y));
                    }
                    const uint32_t key = (bm & ~31u) | (uint32_t)b;
                    const uint32_t om = m[j]; m2[j] = min(m2[j], max(key, om)); m[j] = min(om, key);
                }
            }
        }
    }
    unsigned long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < P; ++j) s += m[j] ^ m2[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (lane == 0) {
        unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        atomicMin(&cyc[2 * smid], t0); atomicMax(&cyc[2 * smid + 1], t1);
    }
}

template <int V> void run(const char *name, int ctas) {
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int blocks = sms * ctas;
    uint32_t *a, *c, *tok, *out; unsigned long long *cyc;
    cudaMalloc(&a, 128 * P * 4); cudaMalloc(&c, 128 * P * 4); cudaMalloc(&tok, (blocks * 4 * 7 + NBLK * 16) * 4); cudaMalloc(&out, blocks * 128 * 4); cudaMalloc(&cyc, sms * 16);
    uint32_t *h = new uint32_t[blocks * 4 * 7 + NBLK * 16];
    for (int i = 0; i < blocks * 4 * 7 + NBLK * 16; ++i) h[i] = 2654435761u * (i + 1);
    cudaMemcpy(tok, h, (blocks * 4 * 7 + NBLK * 16) * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(a, h, 128 * P * 4, cudaMemcpyHostToDevice); cudaMemcpy(c, h + 99, 128 * P * 4, cudaMemcpyHostToDevice);
    unsigned long long *hc = new unsigned long long[2 * sms];
    for (int rep = 0; rep < 2; ++rep) {
        for (int i = 0; i < sms; ++i) { hc[2*i] = ~0ull; hc[2*i+1] = 0; }
        cudaMemcpy(cyc, hc, sms * 16, cudaMemcpyHostToDevice);
        k<V><<<blocks, 128>>>(a, c, tok, out, cyc);
        cudaDeviceSynchronize();
    }
    cudaMemcpy(hc, cyc, sms * 16, cudaMemcpyDeviceToHost);
    double avg = 0; int used = 0;
    for (int i = 0; i < sms; ++i) if (hc[2*i+1]) { avg += (double)(hc[2*i+1] - hc[2*i]); ++used; }
    avg /= used;
    // IMAD warp-instructions per SM sub-partition: 4 warps x ITERS x NBLK x 64, two cycles each
    const double imad_cycles = (double)ctas * ITERS * NBLK * 64 * 2;
    printf("%-34s warps/SM=%2d %9.0f cycles  multiplier busy %.3f\n", name, ctas * 4, avg, imad_cycles / avg);
    cudaFree(a); cudaFree(c); cudaFree(tok); cudaFree(out); cudaFree(cyc); delete[] h; delete[] hc;
}
int main() {
    for (int ctas : {4, 8}) {
        run<0>("compiler order (kernel's source)", ctas);
        run<1>("perm-major asm block 16+tree", ctas);
        run<2>("token-major pairs", ctas);
        run<3>("IMAD only (xor-fold)", ctas);
        run<4>("perm-major volatile triples", ctas);
    }
    return 0;
}
