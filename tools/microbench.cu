// microbench.cu -- integer-pipe issue rates on B200 for the instruction mix of the signature kernel.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu
// Prints warp-instructions / clk / SM (max 4 = one per SMSP per clock).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int ITERS = 4096;
constexpr int CH = 8;  // independent chains per thread

enum Op { OP_IMAD, OP_IMADWIDE, OP_LEAHI, OP_VIMNMX, OP_VIMNMX3, OP_PHASE1, OP_DIRECT, OP_IADD3, OP_LOP3 };

template <int OP>
__global__ void __launch_bounds__(512) k(uint32_t *out, uint32_t seed, unsigned long long *cyc) {
    uint32_t a[CH], b[CH], c[CH];
    uint64_t w[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) { a[i] = seed * (i + 3) + threadIdx.x; b[i] = seed ^ (i * 77u); c[i] = seed + i; w[i] = ((uint64_t)a[i] << 32) | b[i]; }
    unsigned long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            if (OP == OP_IMAD) { a[i] = a[i] * b[i] + c[i]; }
            else if (OP == OP_IMADWIDE) { w[i] = (uint64_t)(uint32_t)w[i] * b[i] + w[i]; }
            else if (OP == OP_LEAHI) { a[i] = b[i] + (a[i] >> 29); b[i] ^= a[i]; }        // LEA.HI + LOP3
            else if (OP == OP_VIMNMX) { a[i] = min(a[i], b[i] + it); }                     // IADD + VIMNMX
            else if (OP == OP_VIMNMX3) { a[i] = min(min(a[i], b[i]), c[i]); c[i] += a[i]; } // VIMNMX3 + IADD
            else if (OP == OP_IADD3) { a[i] = a[i] + b[i] + c[i]; }
            else if (OP == OP_LOP3) { a[i] = (a[i] & b[i]) ^ c[i]; }
            else if (OP == OP_PHASE1) {  // 2 IMAD + 1 VIMNMX3 (two tokens per min3)
                uint32_t v0 = b[i] * (uint32_t)(it * 2654435761u) + c[i];
                uint32_t v1 = b[i] * (uint32_t)(it * 40503u + 1) + c[i];
                a[i] = min(min(a[i], v0), v1);
            } else if (OP == OP_DIRECT) {  // 2x(IMAD.WIDE + IMAD + LEA.HI) + VIMNMX3
                uint32_t h0 = it * 2654435761u, h1 = it * 40503u + 1;
                uint64_t x0 = (uint64_t)b[i] * h0 + w[i];
                uint32_t r0 = (uint32_t)x0 + (((uint32_t)(x0 >> 32) + c[i] * h0) >> 29);
                uint64_t x1 = (uint64_t)b[i] * h1 + w[i];
                uint32_t r1 = (uint32_t)x1 + (((uint32_t)(x1 >> 32) + c[i] * h1) >> 29);
                a[i] = min(min(a[i], r0), r1);
            }
        }
    }
    unsigned long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += a[i] + b[i] + c[i] + (uint32_t)w[i] + (uint32_t)(w[i] >> 32);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char *name, double instr_per_iter_chain, int warps_per_sm) {
    int dev = 0, sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int threads = 512, blocks_per_sm = warps_per_sm * 32 / threads;
    if (blocks_per_sm < 1) { blocks_per_sm = 1; threads = warps_per_sm * 32; }
    int blocks = sms * blocks_per_sm;
    uint32_t *out; unsigned long long *cyc;
    cudaMalloc(&out, (size_t)blocks * threads * 4);
    cudaMalloc(&cyc, blocks * 8);
    k<OP><<<blocks, threads>>>(out, 12345u, cyc);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<OP><<<blocks, threads>>>(out, 12345u, cyc);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    unsigned long long *h = new unsigned long long[blocks];
    cudaMemcpy(h, cyc, blocks * 8, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < blocks; ++i) avg += h[i]; avg /= blocks;
    double warp_instr_per_sm = (double)ITERS * CH * instr_per_iter_chain * (threads / 32) * blocks_per_sm;
    printf("%-10s warps/SM=%2d  %.3f warp-instr/clk/SM  (%.0f cycles, %.3f ms, eff clock %.0f MHz)\n", name,
           warps_per_sm, warp_instr_per_sm / avg, avg, ms, avg / (ms * 1e3));
    delete[] h; cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int w : {16, 32}) {
        run<OP_IMAD>("IMAD", 1, w);
        run<OP_IMADWIDE>("IMAD.WIDE", 1, w);
        run<OP_IADD3>("IADD3", 1, w);
        run<OP_LOP3>("LOP3", 1, w);
        run<OP_LEAHI>("LEA.HI+LOP", 2, w);
        run<OP_VIMNMX>("IADD+MNMX", 2, w);
        run<OP_VIMNMX3>("MNMX3+IADD", 2, w);
        run<OP_PHASE1>("phase1(3)", 3, w);   // per 2 evals (excludes the h0/h1 generation ~3 instr/iter shared)
        run<OP_DIRECT>("direct(7)", 7, w);
    }
    return 0;
}
